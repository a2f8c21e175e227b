"""The oracle (numpy restatement) against fixtures produced by the unmodified reference
(oracle/gen_golden.py).  CPU only."""
import numpy as np

import oracle
from conftest import load_golden

TIGHT = dict(rtol=1e-10, atol=1e-11)


def test_sweep_matches_reference_fixture():
    g = load_golden("sweep_64x48.npz")
    img, ks = g["in_image"], g["in_ks"]
    sigma, kw, kstep = int(g["in_sigma"]), float(g["in_kw"]), float(g["in_kstep"])
    for i, k in enumerate(ks):
        o = oracle.wfr_sweep(img, sigma, k[0], k[1], kw, kstep)
        assert np.allclose(o["lockin"], g["out_lockin"][i], **TIGHT)
        assert np.array_equal(o["w"], g["out_w"][i])
        assert np.allclose(o["grad"], g["out_grad"][i], **TIGHT)
        # kidx is the oracle's addition; it must reproduce w by table lookup
        nx, ny = len(o["wxs"]), len(o["wys"])
        assert o["kidx"].min() >= 0 and o["kidx"].max() < nx * ny
        assert np.array_equal(o["wxs"][o["kidx"] // ny], g["out_w"][i][0])
        assert np.array_equal(o["wys"][o["kidx"] % ny], g["out_w"][i][1])
    o2 = oracle.wfr_sweep(img, sigma, ks[0][0], ks[0][1], kw, kstep, want_grad=False)
    assert np.allclose(o2["lockin"], g["out_optwfr2_lockin"], **TIGHT)
    assert np.array_equal(o2["w"], g["out_optwfr2_w"])


def test_wfr3_matches_reference_fixture():
    g = load_golden("wfr3_64x48.npz")
    o = oracle.wfr_sweep_klist(g["in_image"], int(g["in_sigma"]), g["in_klist"], g["in_kref"],
                               want_grad=False)
    assert np.allclose(o["lockin"], g["out_lockin"], **TIGHT)
    assert np.array_equal(o["w"], g["out_w"])


def test_wfr4_matches_reference_fixture():
    g = load_golden("wfr4_64x48.npz")
    args = (g["in_image"], int(g["in_sigma"]))
    for tag, kl in (("", g["in_klist"]), ("rev_", g["in_klist"][::-1].copy())):
        o = oracle.wfr4(*args, kl, g["in_kref"], float(g["in_dk"]), return_diag=True)
        assert np.allclose(o["lockin"], g["out_" + tag + "lockin"], **TIGHT)
        assert np.array_equal(o["w"], g["out_" + tag + "w"])
        held = o["kidx"] >= 0
        assert np.array_equal(kl[o["kidx"][held]].T, o["w"][:, held])
    # the pairwise table is the reference's per-pixel test
    a = oracle.wfr4_allowed(g["in_klist"], float(g["in_dk"]))
    assert a.shape == (49, 49) and a.diagonal().all() and np.array_equal(a, a.T)


def test_iterate_gpa_and_plane_fit_match_reference_fixture():
    g = load_golden("iterate_96x80.npz")
    assert np.allclose(oracle.fit_plane(g["in_plane"]), g["out_plane_fit"], rtol=1e-9, atol=1e-10)
    prs, w, corr = oracle.iterate_GPA(g["in_image"], g["in_ks"], int(g["in_sigma"]), edge=4, iters=2, kmax_iter=15, kmax=60)
    assert np.allclose(prs, g["out_prs"], rtol=1e-8, atol=1e-8)
    assert np.allclose(w, g["out_w"], **TIGHT) and np.allclose(corr, g["out_corr"], rtol=1e-8, atol=1e-11)


def test_props_oracle_matches_reference_fixture():
    g = load_golden("props_64x48.npz")
    ks, grads, w, nm = g["in_ks_ani"], g["in_grads"], g["in_weights"], float(g["in_nmperpixel"])
    tol = dict(rtol=1e-9, atol=1e-12)
    assert np.allclose(oracle.phasegradient2J(ks, grads, w, nm), g["out_J_iso"], **tol)
    assert np.allclose(oracle.phasegradient2J(ks, grads, w, nm, iso_ref=False), g["out_J_plain"], **tol)
    assert np.allclose(oracle.phasegradient2J(ks, grads, w, nm, sort=-1), g["out_J_sorted_neg"], **tol)
    assert np.allclose(oracle.phasegradient2J(ks, grads, g["in_weights_rankdef"], nm), g["out_J_rankdef"], **tol)
    assert np.allclose(oracle.phasegradient2Jac(ks, grads, w, nm), g["out_Jac"], **tol)
    assert np.array_equal(oracle.props_from_Jac(g["out_Jac"]), g["out_props"])
    assert np.array_equal(oracle.props_from_J(g["out_J_iso"], refangle=-1.5, refscale=0.7), g["out_props_from_J"])


def test_unit_cell_oracle_matches_reference_fixture():
    g = load_golden("ucell_96x80.npz")
    ks, u, shape = g["in_ks"], g["in_u"], g["in_image"].shape
    for z in (2, 3):
        c = oracle.unit_cell_average(g["in_image"], ks, z=z)
        assert np.array_equal(np.isnan(c), np.isnan(g[f"out_cell_z{z}"]))
        assert np.allclose(c, g[f"out_cell_z{z}"], rtol=1e-12, atol=1e-13, equal_nan=True)
        assert np.allclose(oracle.expand_unitcell(g[f"out_cell_z{z}"], ks, shape, z=z), g[f"out_expand_z{z}"], **TIGHT)
        c = oracle.unit_cell_average(g["in_image_def"], ks, u=u, z=z)
        assert np.allclose(c, g[f"out_cell_def_z{z}"], rtol=1e-12, atol=1e-13, equal_nan=True)
        assert np.allclose(oracle.expand_unitcell(g[f"out_cell_def_z{z}"], ks, shape, z=z, u=u),
                           g[f"out_expand_def_z{z}"], **TIGHT)
    assert np.allclose(oracle.unit_cell_average(g["in_image_nan"], ks, z=2), g["out_cell_nan"], rtol=1e-12, atol=1e-13,
                       equal_nan=True)
    assert np.allclose(oracle.expand_unitcell(g["out_cell_z2"], ks, (120, 100), z=2, z2=1.5), g["out_expand_z2_zoom"], **TIGHT)


def test_tail_matches_reference_fixture():
    g = load_golden("tail_64x48.npz")
    ks, ph, w = g["in_ks"], g["in_phases"], g["in_weights"]
    assert np.allclose(oracle.reconstruct_u_inv_from_phases(ks, ph, w), g["out_u"], rtol=1e-9, atol=1e-10)
    assert np.allclose(oracle.reconstruct_u_inv_from_phases(ks, ph, w, weighted_unwrap=False),
                       g["out_u_unweighted_unwrap"], rtol=1e-9, atol=1e-10)
    assert np.allclose(oracle.reconstruct_u_inv_from_phases(ks, g["in_grads"], w, pre_diff=True),
                       g["out_u_prediff"], rtol=1e-9, atol=1e-10)
    assert np.allclose(oracle.extract_displacement_field(g["in_image"], ks, sigma=int(g["in_sigma"])),
                       g["out_u_edf"], rtol=1e-9, atol=1e-10)


def test_fixed_reference_path_matches_fixture():
    g = load_golden("fixed_64x64.npz")
    img, ks, sigma = g["in_image"], g["in_ks"], int(g["in_sigma"])
    rs = np.stack([oracle.lockin_fixed(img, k, sigma) for k in ks])
    assert np.allclose(rs, g["out_lockin"], **TIGHT)
    amps = np.abs(rs)
    unw = np.stack([oracle.phase_unwrap(np.angle(r), np.sqrt(a / a.max()), kmax=25)
                    for r, a in zip(rs, amps)])
    assert np.allclose(unw, g["out_unwrapped"], rtol=1e-9, atol=1e-9)
    unw = g["out_unwrapped"]
    assert np.allclose(oracle.reconstruct_u_inv(ks, unw), g["out_u_unweighted"], **TIGHT)
    assert np.allclose(oracle.reconstruct_u_inv(ks, unw, amps), g["out_u_weighted"], rtol=1e-9, atol=1e-10)
    assert np.allclose(oracle.reconstruct_u_inv(ks, unw, g["in_weights_rankdef"]), g["out_u_rankdef"],
                       rtol=1e-9, atol=1e-10)
    assert np.allclose(oracle.reconstruct_u_inv(ks, unw, use_only_ks=[0, 2]), g["out_u_two_ks"], **TIGHT)


def test_unwrap_matches_fixture_and_known_answer():
    g = load_golden("unwrap.npz")
    psi, psi0 = g["in_psi"], g["in_psi0"]
    r = oracle.phase_unwrap(psi, np.ones_like(psi), kmax=1)
    assert np.allclose(r, g["out_ramp_k1"], **TIGHT)
    # the reference's own known-answer check (tests/test_phase_unwrap.py:22-23)
    assert np.allclose(r - r.mean(), psi0 - psi0.mean())
    assert np.allclose(oracle.phase_unwrap(psi, None, kmax=30), g["out_ramp_unweighted"], **TIGHT)
    assert np.allclose(oracle.phase_unwrap(psi, g["in_gauss"]), g["out_ramp_gauss"], rtol=1e-8, atol=1e-9)
    pr, wr = g["in_psi_r"], g["in_w_r"]
    assert np.allclose(oracle.phase_unwrap(pr, wr, kmax=5), g["out_r_k5"], **TIGHT)
    assert np.allclose(oracle.phase_unwrap(pr, wr, kmax=100), g["out_r_k100"], rtol=1e-8, atol=1e-9)
    assert np.allclose(oracle.phase_unwrap(pr, None), g["out_r_unweighted"], **TIGHT)
    dx, dy = np.diff(pr, axis=1), np.diff(pr, axis=0)
    assert np.allclose(oracle.phase_unwrap_prediff(dx, dy, wr, kmax=7), g["out_r_prediff_k7"], **TIGHT)
    assert np.allclose(oracle.phase_unwrap_prediff(dx, dy), g["out_r_prediff_unweighted"], **TIGHT)


def test_lawler_fujita_matches_fixture():
    g = load_golden("lawler_fujita_48x40.npz")
    u, img = g["in_u"], g["in_image"]
    assert np.allclose(oracle.invert_u_overlap(u), g["out_invert_edge0"], **TIGHT)
    assert np.allclose(oracle.invert_u_overlap(u, iters=5, edge=3), g["out_invert_edge3_it5"], **TIGHT)
    assert np.allclose(oracle.undistort_image(img, u), g["out_undistorted"], **TIGHT)
