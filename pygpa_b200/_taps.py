"""Filter taps of the reference's Gaussian low-pass, per axis.

The reference filters in the Fourier domain with ``scipy.ndimage.fourier_gaussian``
(geometric_phase_analysis.py:44,75,87; cuGPA.py:57).  That is exactly a circular
convolution along each axis with the inverse DFT of exp(-2 pi^2 sigma^2 f^2) — for
sigma >~ 1 the periodised sampled Gaussian.  The CUDA path applies the same kernel
truncated to |d| <= R = ceil(trunc * sigma) (never more than the axis allows).
"""
from __future__ import annotations

import functools

import numpy as np

DEFAULT_TRUNC = 4.5   # phase error 3.2e-4 rad, k flips only at top-2 gaps < 3.3e-6
MAX_TAPS = 446        # kMaxTaps in csrc/lockin.cu


@functools.lru_cache(maxsize=64)
def axis_taps(n, sigma, trunc=DEFAULT_TRUNC):
    """(taps float32[2R+1], R) for an axis of length n."""
    sigma = float(sigma)
    f = np.fft.fftfreq(n)
    kernel = np.fft.ifft(np.exp(-2.0 * np.pi ** 2 * sigma ** 2 * f ** 2)).real   # circular, centred at 0
    r = int(np.ceil(trunc * sigma))
    r = max(0, min(r, (n - 1) // 2))
    if 2 * r + 1 > MAX_TAPS:
        raise ValueError(f"sigma={sigma} needs {2 * r + 1} taps; this build supports {MAX_TAPS}")
    d = np.arange(-r, r + 1)
    taps = np.ascontiguousarray(kernel[d % n], dtype=np.float32)
    taps.setflags(write=False)
    return taps, r
