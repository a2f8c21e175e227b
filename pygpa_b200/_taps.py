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


# ---- multirate factorisation G_sigma = G_a * G_b -------------------------------------------------
MR_WINDOW, MR_HL = 12, 5          # kMrW, kMrHL in csrc/lockin.cu


def _kernel_taps(n, sigma, r):
    f = np.fft.fftfreq(n)
    kernel = np.fft.ifft(np.exp(-2.0 * np.pi ** 2 * sigma ** 2 * f ** 2)).real
    return np.ascontiguousarray(kernel[np.arange(-r, r + 1) % n], dtype=np.float32)


@functools.lru_cache(maxsize=64)
def multirate_taps(n, m, sigma, trunc=DEFAULT_TRUNC):
    """Parameters of the multirate sweep for an (n, m) frame, or None when it does not apply.

    stride S: the largest of 8, 4, 2 with sigma_b = c S, c in [1.0, 1.1], sigma_b^2 <= 0.2 sigma^2
    (then the aliasing term exp(-2 pi^2 sigma_a^2 sigma_b^2 / (sigma^2 S^2)) is <= exp(-15.8) = 1.4e-7
    of the out-of-band content); both frame axes must be multiples of S and hold the filters.
    Returns dict(S, Ra_x, Ra_y, Rb, taps_ax, taps_ay, taps_bx, taps_by, sigma_a, sigma_b)."""
    sigma = float(sigma)
    for s in (8, 4, 2):
        if n % s or m % s or n // s < MR_WINDOW or m // s < MR_WINDOW:
            continue
        c = min(1.1, np.sqrt(0.2) * sigma / s)
        if c < 1.0:
            continue
        sigma_b = c * s
        sigma_a = float(np.sqrt(sigma ** 2 - sigma_b ** 2))
        rb = int(np.ceil(trunc * sigma_b))
        ra = int(np.ceil(trunc * sigma_a))
        # the decimating pass has statically scheduled kernels for an EVEN number of taps per phase:
        # round the filter length up to the next such size and spend the extra taps (at most one per
        # phase) on a slightly larger truncation radius
        jt = 2 * (-(-(-(-(2 * ra + 1) // s)) // 2))
        ra = max(ra, (s * jt - 1) // 2)
        if rb > MR_HL * s:
            continue
        if 2 * ra + 1 > min(n, m) or s * (-(-(2 * ra + 1) // s)) + 2 > MAX_TAPS:
            continue
        return dict(S=s, Ra_x=ra, Ra_y=ra, Rb=rb, sigma_a=sigma_a, sigma_b=sigma_b,
                    taps_ax=_kernel_taps(n, sigma_a, ra), taps_ay=_kernel_taps(m, sigma_a, ra),
                    taps_bx=_kernel_taps(n, sigma_b, rb), taps_by=_kernel_taps(m, sigma_b, rb))
    return None
