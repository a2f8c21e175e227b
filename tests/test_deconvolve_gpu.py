"""K8: gaussian_deconvolve (geometric_phase_analysis.py:892-904) on the GPU against the oracle's
restatement of scikit-image's Wiener filter.  PARITY UNPINNED at the reference boundary: scikit-image is
absent here, so there is no reference-generated fixture for this function (oracle/wiener_numpy.py)."""
import numpy as np
import pytest

import oracle
from pygpa_b200 import geometric_phase_analysis as GPA

pytestmark = pytest.mark.gpu


def _field(shape, seed):
    rng = np.random.default_rng(seed)
    return rng.normal(size=shape).cumsum(axis=-2).cumsum(axis=-1) / 50


@pytest.mark.parametrize("shape,sigma,dr", [((2, 70, 90), 5, 10), ((64, 64), 3, 6), ((3, 100, 45), 4, 8),
                                            ((2, 236, 216), 10, 20), ((1, 1000, 600), 10, 20)])
def test_matches_oracle(shape, sigma, dr):
    """Odd / even / prime-ish padded lengths (all Bluestein), 2-D input and stacks, transform lengths
    256 ... 4096."""
    data = _field(shape, sum(shape))
    got = GPA.gaussian_deconvolve(data, sigma, dr)
    ref = oracle.gaussian_deconvolve(data, sigma, dr)
    assert got.shape == data.shape and got.dtype == np.float64
    assert np.abs(got - ref).max() < 1e-10 * max(1.0, np.abs(ref).max())


def test_balance_and_maximum_transform_length():
    data = _field((1, 2048, 2048), 3)           # C3 frame: padded 2128 -> 8192-point transforms
    got = GPA.gaussian_deconvolve(data, 10, 20, balance=100.0)
    ref = oracle.gaussian_deconvolve(data, 10, 20, balance=100.0)
    assert np.abs(got - ref).max() < 1e-9 * np.abs(ref).max()


def test_constant_field_is_preserved_and_oversize_is_refused():
    const = np.full((40, 50), 3.25)
    assert np.abs(GPA.gaussian_deconvolve(const, 4, 8) - 3.25).max() < 1e-12      # W(0) = 1
    from pygpa_b200._lib import GpaError
    with pytest.raises(GpaError):
        GPA.gaussian_deconvolve(np.zeros((4100, 64)), 4, 8)
    with pytest.raises(GpaError):
        GPA.gaussian_deconvolve(np.zeros((12, 64)), 4, 8)                          # reflect padding wider than the frame
