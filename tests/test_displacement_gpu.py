"""K3 (+K2) parity: phase -> displacement solves against the reference-made fixtures and the
oracle, and the reference's end-to-end accuracy test."""
import numpy as np
import pytest

import oracle
from conftest import load_golden
from parity import DISP_TOL
from pygpa_b200 import cuGPA, synth
from pygpa_b200 import geometric_phase_analysis as GPA

pytestmark = pytest.mark.gpu


def test_reconstruct_u_inv_branches_golden():
    g = load_golden("fixed_64x64.npz")
    ks, unw = g["in_ks"], g["out_unwrapped"]
    amps = np.abs(g["out_lockin"])
    assert np.allclose(GPA.reconstruct_u_inv(ks, unw), g["out_u_unweighted"], rtol=1e-9, atol=1e-10)
    assert np.allclose(GPA.reconstruct_u_inv(ks, unw, amps), g["out_u_weighted"], rtol=1e-8, atol=1e-9)
    assert np.allclose(GPA.reconstruct_u_inv(ks, unw, use_only_ks=[0, 2]), g["out_u_two_ks"], rtol=1e-9, atol=1e-10)
    # rank-deficient pixels: minimum-norm solution (all-zero weights -> 0), as the oracle
    got = GPA.reconstruct_u_inv(ks, unw, g["in_weights_rankdef"])
    ref = oracle.reconstruct_u_inv(ks, unw, g["in_weights_rankdef"])
    assert np.allclose(got, ref, rtol=1e-8, atol=1e-9)
    assert not got[:, :3].any()
    with pytest.raises(ValueError):
        GPA.reconstruct_u_inv(ks[:2], unw[:2])


def test_weighted_lstsq_random_against_oracle():
    rng = np.random.default_rng(8)
    d, n, m = 3, 37, 53
    ks = synth.primary_ks(0.08, 11.0, d)
    K = 2 * np.pi * ks
    b = rng.normal(size=(d, n, m))
    w = rng.uniform(1e-6, 1, size=(d, n + 2, m + 1)) ** 3      # larger than b, strongly varying
    ref = oracle.weighted_lstsq(b, K, w)
    got = GPA.myweighed_lstsq(b, K, w)
    # Least-squares conditioning: with a non-zero residual two backward-stable solvers (LAPACK SVD
    # in the oracle, QR here) may differ by ~ eps * cond^2.  Well-conditioned pixels are tight.
    a = w[:, :n, :m].reshape(d, -1).T[:, :, None] * K[None]
    sv = np.linalg.svd(a, compute_uv=False)
    cond = (sv[:, 0] / sv[:, 1]).reshape(n, m)
    err = np.abs(got - ref).max(axis=0) / np.abs(ref).max(axis=0)
    assert (cond < 100).mean() > 0.8
    assert err[cond < 100].max() < 1e-10
    assert np.all(err <= 1e-10 + 100 * np.finfo(float).eps * cond ** 2)


def test_from_phases_golden_and_oracle():
    g = load_golden("tail_64x48.npz")
    ks, ph, w = g["in_ks"], g["in_phases"], g["in_weights"]
    assert np.allclose(GPA.reconstruct_u_inv_from_phases(ks, ph, w), g["out_u"], rtol=1e-6, atol=1e-7)
    assert np.allclose(GPA.reconstruct_u_inv_from_phases(ks, ph, w, weighted_unwrap=False),
                       g["out_u_unweighted_unwrap"], rtol=1e-6, atol=1e-7)
    assert np.allclose(GPA.reconstruct_u_inv_from_phases(ks, g["in_grads"], w, pre_diff=True),
                       g["out_u_prediff"], rtol=1e-6, atol=1e-7)


def test_extract_displacement_field_golden():
    """Whole adaptive pipeline on the GPU vs the reference's result: displacement within 1e-3 px."""
    g = load_golden("tail_64x48.npz")
    img, ks, sigma = g["in_image"], g["in_ks"], int(g["in_sigma"])
    u_dev = GPA.extract_displacement_field(img, ks, sigma=sigma)                     # device-resident chain
    u_host, gs = GPA.extract_displacement_field(img, ks, sigma=sigma, return_gs=True)  # reference-style glue
    assert len(gs) == 3 and set(gs[0]) == {"lockin", "w"}
    for u in (u_dev, u_host):
        assert u.shape == (2,) + img.shape
        assert np.abs(u - g["out_u_edf"]).max() < DISP_TOL
    u_cu = GPA.extract_displacement_field(img, ks, sigma=sigma, wfr_func=cuGPA.wfr2_grad_opt)
    assert np.abs(u_cu - g["out_u_edf"]).max() < DISP_TOL
    # deconvolve=True (geometric_phase_analysis.py:928-929): the Wiener step applied to the reference's u
    u_dec = GPA.extract_displacement_field(img, ks, sigma=sigma, deconvolve=True)
    ref_dec = oracle.gaussian_deconvolve(g["out_u_edf"], sigma, 2 * sigma)
    assert np.abs(u_dec - ref_dec).max() < DISP_TOL


def test_displacement_field_accuracy_like_reference_test():
    """tests/test_geometric_phase_analysis.py:61-66 of the reference with a seeded noise field:
    u = -extract_displacement_field(deformed + noise, ks) recovers the Gaussian bump to < 0.9 px
    on the interior; and the GPU result tracks the oracle's within 1e-3 px."""
    import scipy.ndimage as ndi
    s = 256
    shape = (s, s)
    ks = synth.primary_ks(0.1, 7.0, 3)
    bump = 0.5 * synth.gaussian_bump(shape)      # this generator's lattice is weaker than latticegen's order-2 one:
    deformed = synth.lattice_image(shape, ks, bump, second_order=0.3)       # halve the bump, noise 2 instead of 5
    noise = ndi.gaussian_filter(2 * np.random.default_rng(0).normal(size=shape), sigma=0.5)
    img = deformed + noise
    u, gs = GPA.extract_displacement_field(img, ks, return_gs=True)
    u = -u
    assert u.shape == bump.shape
    assert np.all(np.abs(u - bump)[:, 20:-20, 20:-20] < 0.9)
    u_dev = -GPA.extract_displacement_field(img, ks)            # device-resident chain: same numbers
    assert np.abs(u_dev - u).max() < 1e-9

    def sweep(im, s_, kx, ky, kw, kstep):
        return oracle.wfr_sweep(im, s_, kx, ky, kw, kstep, want_grad=False, return_diag=True)
    u_ref, gs_ref = oracle.extract_displacement_field(img, ks, return_gs=True, sweep=sweep)
    # A pixel where the sweep legitimately picks another candidate (near-tie, gap < 1e-5) carries a
    # different lock-in phase, which shows up in u AT that pixel.  Everything else is within 1e-3 px.
    flips = np.zeros(shape, dtype=bool)
    for g_, r_ in zip(gs, gs_ref):
        differs = ~np.all(g_['w'] == r_['w'], axis=0)
        gap = (r_['amp1'] - r_['amp2']) / r_['amp1']
        assert np.all(gap[differs] < 1e-5)
        flips |= differs
    assert flips.mean() < 1e-3
    near = ndi.binary_dilation(flips, iterations=2)
    assert np.abs(-u_ref - u).max(axis=0)[~near].max() < DISP_TOL
