"""The oracle's restatement of SciPy's cubic-spline map_coordinates (the arithmetic behind the
Lawler-Fujita step) against the installed SciPy.  CPU only."""
import numpy as np
import pytest
import scipy.ndimage as ndi

from oracle import spline_restatement as sr


@pytest.mark.parametrize("shape", [(9, 7), (40, 33), (5, 64)])
@pytest.mark.parametrize("mode", ["nearest", "constant"])
def test_map_coordinates_restatement(shape, mode):
    rng = np.random.default_rng(shape[0])
    a = rng.normal(size=shape)
    far = 3 * max(shape)
    co = np.stack([rng.uniform(-far, shape[0] + far, size=500), rng.uniform(-far, shape[1] + far, size=500)])
    co[:, :250] = np.stack([rng.uniform(-2, shape[0] + 1, size=250), rng.uniform(-2, shape[1] + 1, size=250)])
    co[:, 0] = (0.0, 0.0)
    co[:, 1] = (shape[0] - 1.0, shape[1] - 1.0)
    ref = ndi.map_coordinates(a, co, order=3, mode=mode)
    got = sr.map_coordinates_2d(a, co, mode)
    assert np.abs(got - ref).max() < 1e-13
    grid = np.mgrid[:shape[0], :shape[1]].astype(float)
    assert np.abs(sr.map_coordinates_2d(a, grid, mode) - a).max() < 1e-13     # interpolating spline


def test_prefilter_matches_scipy_spline_filter():
    rng = np.random.default_rng(0)
    a = rng.normal(size=(31, 18))
    assert np.abs(sr.spline_coefficients(a, 'constant') - ndi.spline_filter(a, order=3, mode='mirror')).max() < 1e-13
    pad = np.pad(a, 12, mode='edge')
    assert np.abs(sr.spline_coefficients(a, 'nearest') - ndi.spline_filter(pad, order=3, mode='reflect')).max() < 1e-13
