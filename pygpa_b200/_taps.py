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


def _pass2_tile_fits(s, j):
    """Shared-memory budget of the decimating pass-2 kernels (csrc/lockin.cu, launch_mr): the plane tile of
    S (W2 kP + J + kAhead + 1) fine rows x 32 columns plus two carrier buffers per warp group must fit 227 KB."""
    w2 = 4 if s == 8 else 8
    return s * (w2 * 16 + j + 3) * (32 + 4) * 8 <= 227 * 1024


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
        if not _pass2_tile_fits(s, -(-(2 * ra + 1) // s)):
            continue          # large sigma: fall through to a smaller stride (shorter tile in fine rows)
        return dict(S=s, Ra_x=ra, Ra_y=ra, Rb=rb, sigma_a=sigma_a, sigma_b=sigma_b,
                    taps_ax=_kernel_taps(n, sigma_a, ra), taps_ay=_kernel_taps(m, sigma_a, ra),
                    taps_bx=_kernel_taps(n, sigma_b, rb), taps_by=_kernel_taps(m, sigma_b, rb))
    return None


# ---- split pass 2: G_a = G_1 * G_2, anchor stage shared by the candidates of a plane ---------------
SPLIT_TOL = 1.3e-6       # worst-case transfer-function error; the 4.5 sigma truncation of G_a alone is 1.3e-6
SPLIT_TRUNC1 = 6.0       # stage A runs once per plane: its truncation is free
SPLIT_MAX_LOG_GAIN = float(np.log(8.0))   # largest re-amplification c of a candidate (C3: 2.5; stride 8 at sigma = 22: 3.5)


def _split_error(s, sigma_a, sigma_1, r1, h, dw, nf=2048):
    """max over input frequency f (cycles / fine pixel, relative to the anchor) of
    |c G_1t(f) h_t(S (f + delta)) - G_a(f + dw)|: the split pipeline's response to e^{2 pi i f x}
    against the ideal candidate-centred Gaussian, truncation and coarse-rate aliasing included."""
    sigma_2 = np.sqrt(sigma_a ** 2 - sigma_1 ** 2)
    f = (np.arange(nf) - nf // 2) / nf
    d = np.arange(-r1, r1 + 1)
    g1 = np.exp(-d ** 2 / (2 * sigma_1 ** 2)) / (sigma_1 * np.sqrt(2 * np.pi))
    G1 = np.exp(-2j * np.pi * np.outer(f, d)) @ g1
    m = np.arange(-h, h + 1)
    h2 = s * np.exp(-(s * m) ** 2 / (2 * sigma_2 ** 2)) / (sigma_2 * np.sqrt(2 * np.pi))
    delta = dw * sigma_a ** 2 / sigma_2 ** 2
    log_c = 2 * np.pi ** 2 * dw ** 2 * sigma_a ** 2 * sigma_1 ** 2 / sigma_2 ** 2
    if log_c > SPLIT_MAX_LOG_GAIN:  # the candidate sits far out on G_1's slope: its band leaves the anchor stage attenuated
        return float("inf")         # by 1/c and fp32 rounding noise comes back amplified by c
    c = np.exp(log_c)
    H2 = np.exp(-2j * np.pi * np.outer(s * (f + delta), m)) @ h2
    return float(np.abs(c * G1 * H2 - np.exp(-2 * np.pi ** 2 * sigma_a ** 2 * (f + dw) ** 2)).max())


@functools.lru_cache(maxsize=64)
def _split_plan(n, s, sigma_a, dw_max):
    def best_for(h):
        """(error, sigma_1, sigma_2, R1) of the best coarse-rate sigma for 2h+1 taps, or None."""
        best = None
        for s2c in np.arange(1.35, min(h / 4.4, 2.4) + 1e-9, 0.05):      # coarse-rate sigma of G_2
            sigma_2 = float(s2c * s)
            if sigma_2 >= 0.98 * sigma_a:
                break
            sigma_1 = float(np.sqrt(sigma_a ** 2 - sigma_2 ** 2))
            r1 = int(np.ceil(SPLIT_TRUNC1 * sigma_1))
            j1 = -(-(2 * r1 + 1) // s)
            j1 = max(18, j1 + (j1 & 1))                 # even, >= 18: the statically scheduled stage-A kernels
            r1 = (s * j1 - 1) // 2
            if r1 + s * (h + 1) > n or n // s <= 2 * (-(-r1 // s) + 1) or s * j1 + 2 > MAX_TAPS or not _pass2_tile_fits(s, j1):
                continue
            err = max(_split_error(s, sigma_a, sigma_1, r1, h, dw) for dw in (dw_max, 0.5 * dw_max))
            if best is None or err < best[0]:
                best = (err, sigma_1, sigma_2, r1)
        return best

    def plan(h, best):
        err, sigma_1, sigma_2, r1 = best
        d = np.arange(-r1, r1 + 1)
        t1 = (np.exp(-d ** 2 / (2 * sigma_1 ** 2)) / (sigma_1 * np.sqrt(2 * np.pi))).astype(np.float32)
        m = np.arange(-h, h + 1)
        t2 = (s * np.exp(-(s * m) ** 2 / (2 * sigma_2 ** 2)) / (sigma_2 * np.sqrt(2 * np.pi))).astype(np.float32)
        t1.setflags(write=False)
        t2.setflags(write=False)
        return dict(R1=r1, H=h, sigma_1=sigma_1, sigma_2=sigma_2, taps_1=t1, taps_2=t2, err=err,
                    c_max=float(np.exp(2 * np.pi ** 2 * dw_max ** 2 * sigma_a ** 2 * sigma_1 ** 2 / sigma_2 ** 2)))

    # the longest filter (23 coarse taps) first: if even that misses the tolerance — grids too wide for one
    # anchor — nothing shorter is tried; otherwise the shortest filter that meets it wins
    longest = best_for(11)
    if longest is None or longest[0] > SPLIT_TOL:
        return None
    for h in (6, 7, 8, 9, 10):                          # 13 ... 21 coarse taps per candidate
        best = best_for(h)
        if best is not None and best[0] <= SPLIT_TOL:
            return plan(h, best)
    return plan(11, longest)


def split_taps(n, mr, wx_rows):
    """Parameters of the split pass 2 (csrc/lockin.cu, k_mr_pass2b) for an axis of length n, the
    multirate plan ``mr`` and the candidate axis ``wx_rows``, or None when no factorisation
    G_a = G_1 * G_2 meets SPLIT_TOL with at most 23 coarse taps (wide grids: the candidates then
    keep their own full-rate pass 2).  The anchor is wx_rows[len // 2]."""
    wx = np.asarray(wx_rows, dtype=np.float64)
    if wx.size < 8:
        return None
    dw_max = float(np.abs(wx - wx[wx.size // 2]).max())
    # quantise dw_max upwards so that plans of neighbouring peaks share the cached search
    q = 2.0 ** (np.floor(np.log2(max(dw_max, 1e-12))) - 4)
    return _split_plan(int(n), int(mr["S"]), float(mr["sigma_a"]), float(np.ceil(dw_max / q) * q))
