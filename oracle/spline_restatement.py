"""Restatement of the third-party arithmetic behind the Lawler-Fujita step (oracle; test
infrastructure only).

The reference calls ``scipy.ndimage.map_coordinates`` with its defaults (order 3, prefilter)
in two modes: ``mode='nearest'`` for the fixed-point inversion of u
(pyGPA/geometric_phase_analysis.py:292-299) and ``mode='constant', cval=0`` for the final
resampling (:973).  SciPy (pinned here: 1.18.1, the version in this image; the algorithm is
unchanged since 1.6) is compiled C, so its rules are restated below and checked against the
installed SciPy in tests/test_spline_restatement.py.  The CUDA kernels follow this file.

Rules (verified to 2e-15 against SciPy, including coordinates far outside the array):
  * cubic B-spline, pole z = sqrt(3) - 2, gain (1 - z)(1 - 1/z) = 6 per axis, causal then
    anti-causal recursion per axis;
  * 'nearest': the array is first padded by 12 edge-replicated samples per side; the padded array
    is filtered with HALF-sample-symmetric ("reflect") initial conditions; the spline is evaluated
    at coordinate + 12 with the TAP INDICES clamped to the padded array (coordinates are not
    clamped);
  * 'constant': no padding; WHOLE-sample-symmetric ("mirror") initial conditions; taps that fall
    outside are mirrored (period 2n - 2); a coordinate outside [0, n-1] along any axis yields cval.
"""
import numpy as np

POLE = np.sqrt(3.0) - 2.0
NPAD = 12


def prefilter_axis(a, axis, boundary):
    """In-place-style cubic B-spline prefilter along one axis; boundary 'reflect' or 'mirror'."""
    c = np.moveaxis(np.array(a, dtype=np.float64), axis, 0) * ((1 - POLE) * (1 - 1 / POLE))
    n = c.shape[0]
    if n == 1:
        return np.array(a, dtype=np.float64)
    z = POLE
    zi = z ** np.arange(n).reshape((n,) + (1,) * (c.ndim - 1))
    if boundary == 'mirror':
        zn1 = z ** (n - 1)
        acc = c[0] + zn1 * c[n - 1] + (zi[1:n - 1] * (c[1:n - 1] + zn1 * c[n - 2:0:-1])).sum(axis=0)
        c[0] = acc / (1 - zn1 * zn1)
    else:
        zn = z ** n
        acc = (zi * (c + zn * c[::-1])).sum(axis=0)
        c[0] = acc * z / (1 - zn * zn) + c[0]
    for i in range(1, n):
        c[i] += z * c[i - 1]
    if boundary == 'mirror':
        c[n - 1] = (z * c[n - 2] + c[n - 1]) * z / (z * z - 1)
    else:
        c[n - 1] *= z / (z - 1)
    for i in range(n - 2, -1, -1):
        c[i] = z * (c[i + 1] - c[i])
    return np.moveaxis(c, 0, axis)


def spline_coefficients(a, mode):
    """Coefficient array map_coordinates interpolates from: (padded for 'nearest')."""
    a = np.asarray(a, dtype=np.float64)
    if mode == 'nearest':
        c = np.pad(a, NPAD, mode='edge')
        boundary = 'reflect'
    elif mode == 'constant':
        c, boundary = a, 'mirror'
    else:
        raise ValueError(mode)
    for ax in range(c.ndim):
        c = prefilter_axis(c, ax, boundary)
    return c


def _weights(t):
    return np.stack([(1 - t) ** 3 / 6, (3 * t ** 3 - 6 * t ** 2 + 4) / 6,
                     (-3 * t ** 3 + 3 * t ** 2 + 3 * t + 1) / 6, t ** 3 / 6])


def map_coordinates_2d(a, coords, mode, cval=0.0):
    """map_coordinates(a, coords, order=3, mode=mode, cval=cval) for 2-D a, vectorised."""
    a = np.asarray(a, dtype=np.float64)
    coef = spline_coefficients(a, mode)
    npad = NPAD if mode == 'nearest' else 0
    cx = np.asarray(coords[0], dtype=np.float64) + npad
    cy = np.asarray(coords[1], dtype=np.float64) + npad
    fx, fy = np.floor(cx), np.floor(cy)
    wx, wy = _weights(cx - fx), _weights(cy - fy)
    ix, iy = fx.astype(np.int64) - 1, fy.astype(np.int64) - 1
    n, m = coef.shape

    def taps(i0, length):
        idx = i0[None] + np.arange(4).reshape((4,) + (1,) * i0.ndim)
        if mode == 'nearest':
            return np.clip(idx, 0, length - 1)
        if length == 1:
            return np.zeros_like(idx)
        s2 = 2 * length - 2
        idx = np.abs(idx) % s2
        return np.where(idx >= length, s2 - idx, idx)
    tx, ty = taps(ix, n), taps(iy, m)
    out = np.zeros(cx.shape)
    for i in range(4):
        for j in range(4):
            out += wx[i] * wy[j] * coef[tx[i], ty[j]]
    if mode == 'constant':
        outside = (cx < 0) | (cx > n - 1) | (cy < 0) | (cy > m - 1)
        out = np.where(outside, cval, out)
    return out
