"""K4 parity: Lawler-Fujita inversion / resampling vs reference-made fixtures and the oracle."""
import numpy as np
import pytest

import oracle
from conftest import load_golden
from pygpa_b200 import synth
from pygpa_b200 import geometric_phase_analysis as GPA

pytestmark = pytest.mark.gpu


def test_golden_fixture():
    g = load_golden("lawler_fujita_48x40.npz")
    u, img = g["in_u"], g["in_image"]
    got = GPA.invert_u_overlap(u)
    assert got.shape == u.shape and got.dtype == np.float64
    assert np.abs(got - g["out_invert_edge0"]).max() < 1e-9
    got = GPA.invert_u_overlap(u, iters=5, edge=3)
    assert got.shape == (2, 48 + 6, 40 + 6)
    assert np.abs(got - g["out_invert_edge3_it5"]).max() < 1e-9
    assert np.abs(GPA.undistort_image(img, u) - g["out_undistorted"]).max() < 1e-9
    assert np.abs(GPA.invert_u(u, iters=7) - oracle.invert_u_overlap(u, iters=7)).max() < 1e-9
    for edge in (0, 3):        # the reference's invert_u: `- edge` in the iterations only (geometric_phase_analysis.py:257)
        got = GPA.invert_u(u, iters=6, edge=edge)
        assert got.shape == u.shape
        assert np.abs(got - oracle.invert_u(u, iters=6, edge=edge)).max() < 1e-9
    with pytest.raises(NotImplementedError):
        GPA.invert_u_overlap(u, mode='reflect')


@pytest.mark.parametrize("shape", [(130, 100), (64, 257)])
def test_oracle_parity_large_displacements(shape):
    """Displacements larger than scipy's 12-sample pad push coordinates far outside the frame:
    exercises the clamped-tap rule ('nearest') and the zero fill ('constant')."""
    u = synth.smooth_random_field(shape, 0.25, seed=shape[0])
    u *= 15.0 / np.abs(u).max()
    rng = np.random.default_rng(1)
    img = rng.normal(size=shape)
    for edge, iters in ((0, 35), (7, 4)):
        ref = oracle.invert_u_overlap(u, iters=iters, edge=edge)
        got = GPA.invert_u_overlap(u, iters=iters, edge=edge)
        assert np.abs(got - ref).max() < 1e-8
    ref = oracle.undistort_image(img, u)
    got = GPA.undistort_image(img, u)
    assert (ref == 0).any()                       # some samples fall outside the frame
    assert np.array_equal(ref == 0, got == 0)
    assert np.abs(got - ref).max() < 1e-8


def test_reconstruction_like_reference_test():
    """tests/test_geometric_phase_analysis.py:73-79: undistort_image(deformed, true_u) recovers the
    undeformed lattice to 2 % of its maximum."""
    shape = (300, 300)
    ks = synth.primary_ks(0.1, 7.0, 3)
    bump = synth.gaussian_bump(shape)
    original = synth.lattice_image(shape, ks)
    deformed = synth.lattice_image(shape, ks, bump)
    u_inv = GPA.invert_u_overlap(-bump)
    assert u_inv.shape == bump.shape
    rec = GPA.undistort_image(deformed, bump)
    assert np.all(np.abs(rec - original) / np.abs(original).max() < 0.02)
