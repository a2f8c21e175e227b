"""Synthetic lattice / moire frames for tests and benchmarks (own generator).

The reference's tests build their inputs with the third-party ``latticegen`` package
(tests/test_geometric_phase_analysis.py:25-41), which is not available offline.  This
module produces the same kind of image from first principles:

    image(r) = sum_j cos(2 pi k_j . (r + u(r)))     (+ weaker second-order terms)

with k_j = r_k (cos(xi0 + 60deg j), sin(xi0 + 60deg j)), axis 0 = x as everywhere in
pyGPA, plus seeded noise.  Everything is deterministic given the seeds.
"""
from __future__ import annotations

import numpy as np

__all__ = ["primary_ks", "gaussian_bump", "smooth_random_field", "lattice_image",
           "sweep_params", "make_config", "make_config_device", "frame_series_device"]


def primary_ks(r_k, xi0_deg=0.0, n=3):
    """First ``n`` of the six first-order hexagonal k-vectors, (n, 2), cycles/pixel."""
    ang = np.deg2rad(xi0_deg + 60.0 * np.arange(n))
    return r_k * np.stack([np.cos(ang), np.sin(ang)], axis=1)


def gaussian_bump(shape):
    """Displacement used by the reference tests (tests/test_geometric_phase_analysis.py:12-17),
    generalised to non-square frames: u_x = 0.5 xp exp(-((xp/(N/8))^2 + 1.2 (yp/(M/6))^2)/2), u_y = 0."""
    n, m = shape
    xp, yp = np.meshgrid(np.arange(n) - n // 2, np.arange(m) - m // 2, indexing='ij')
    ux = 0.5 * xp * np.exp(-0.5 * ((xp / (n / 8)) ** 2 + 1.2 * (yp / (m / 6)) ** 2))
    return np.stack([ux, np.zeros_like(ux)])


def smooth_random_field(shape, max_grad, seed, modes=8):
    """Smooth random displacement (2, N, M): a few low-frequency Fourier modes, scaled so
    that the largest component of the displacement gradient equals ``max_grad``."""
    rng = np.random.default_rng(seed)
    n, m = shape
    x = np.arange(n)[:, None] / n
    y = np.arange(m)[None, :] / m
    u = np.zeros((2, n, m))
    for c in range(2):
        for _ in range(modes):
            fx, fy = rng.integers(0, 3, size=2)
            if fx == 0 and fy == 0:
                fx = 1
            amp = rng.normal()
            ph = rng.uniform(0, 2 * np.pi)
            u[c] += amp * np.sin(2 * np.pi * (fx * x + fy * y) + ph)
    g = max(np.abs(np.gradient(u[0])).max(), np.abs(np.gradient(u[1])).max())
    return u * (max_grad / g)


def lattice_image(shape, ks, u=None, second_order=0.0, noise=0.0, seed=0):
    """Sum of cosines along ``ks`` (d, 2) sampled at r + u(r); optional second-order
    (k_i + k_j) terms with relative weight ``second_order`` and white Gaussian noise."""
    n, m = shape
    x = np.arange(n, dtype=np.float64)[:, None] * np.ones((1, m))
    y = np.arange(m, dtype=np.float64)[None, :] * np.ones((n, 1))
    if u is not None:
        x = x + u[0]
        y = y + u[1]
    ks = np.asarray(ks, dtype=np.float64)
    img = np.zeros(shape)
    for k in ks:
        img += np.cos(2 * np.pi * (k[0] * x + k[1] * y))
    if second_order:
        for i in range(len(ks)):
            k2 = ks[i] + ks[(i + 1) % len(ks)]
            img += second_order * np.cos(2 * np.pi * (k2[0] * x + k2[1] * y))
    if noise:
        img = img + noise * np.random.default_rng(seed).normal(size=shape)
    return img


def sweep_params(ks, n_grid, kwscale=2.5):
    """kw and kstep giving exactly ``n_grid`` candidates per axis: kw = mean|k|/kwscale as in
    extract_displacement_field (geometric_phase_analysis.py:915), kstep = 2 kw/(n - 0.5)."""
    kw = float(np.linalg.norm(ks, axis=1).mean() / kwscale)
    kstep = 2 * kw / (n_grid - 0.5)
    return kw, kstep


_CONFIGS = {
    # name: (N, r_k, max_grad, n_grid, seed_u, seed_noise)
    'C2': (1024, 0.05, 0.15, 21, 1, 2),
    'C3': (2048, 0.05, 0.30, 41, 3, 4),
}


def make_config(name, size=None, n_grid=None):
    """Synthetic input for BASELINE.json configs: returns dict with
    image (float64, zero-mean), ks (3,2), sigma, kw, kstep, n_grid, u."""
    n, r_k, max_grad, ng, su, sn = _CONFIGS[name]
    if size is not None:
        n = size
    if n_grid is not None:
        ng = n_grid
    ks = primary_ks(r_k, 7.0, 3)
    u = smooth_random_field((n, n), max_grad, su)
    img = lattice_image((n, n), ks, u, noise=0.3, seed=sn)
    img = img - img.mean()
    kw, kstep = sweep_params(ks, ng)
    return dict(image=img, ks=ks, sigma=10, kw=kw, kstep=kstep, n_grid=ng, u=u, name=name)


# ----------------------------------------------------------------------------------------------
# device-side generators for the large benchmark inputs (BASELINE configs 4 and 5): the same formulas
# evaluated with torch on the GPU, because an 8192 x 8192 frame takes ~40 s of NumPy on the host.
# Synthetic-input plumbing only; nothing here is on the measured path.
# ----------------------------------------------------------------------------------------------
def _smooth_field_device(shape, max_grad, seed, device, modes=8):
    import torch
    rng = np.random.default_rng(seed)
    n, m = shape
    x = torch.arange(n, dtype=torch.float64, device=device)[:, None] / n
    y = torch.arange(m, dtype=torch.float64, device=device)[None, :] / m
    u = torch.zeros((2, n, m), dtype=torch.float64, device=device)
    for c in range(2):
        for _ in range(modes):
            fx, fy = rng.integers(0, 3, size=2)
            if fx == 0 and fy == 0:
                fx = 1
            amp = rng.normal()
            ph = rng.uniform(0, 2 * np.pi)
            u[c] += amp * torch.sin(2 * np.pi * (int(fx) * x + int(fy) * y) + ph)
    g = max(float(torch.diff(u, dim=1).abs().max()), float(torch.diff(u, dim=2).abs().max()))
    return u * (max_grad / g)


def _lattice_device(shape, ks, u, noise, seed, device):
    import torch
    n, m = shape
    x = torch.arange(n, dtype=torch.float64, device=device)[:, None] + u[0]
    y = torch.arange(m, dtype=torch.float64, device=device)[None, :] + u[1]
    img = torch.zeros((n, m), dtype=torch.float64, device=device)
    for k in np.asarray(ks, dtype=np.float64):
        img += torch.cos(2 * np.pi * (float(k[0]) * x + float(k[1]) * y))
    if noise:
        gen = torch.Generator(device=device)
        gen.manual_seed(int(seed))
        img += noise * torch.randn((n, m), dtype=torch.float64, device=device, generator=gen)
    return img


def make_config_device(name, device, size=None, n_grid=None):
    """C5: stitched 8192 x 8192 mosaic — a C3-type smooth twist-gradient field plus a
    constant extra rotation per tile of a 4 x 4 mosaic, blended over 128 px — generated on `device`.
    'C2' / 'C3' give the same kind of frame as make_config (not the same random numbers).  Returns dict with
    image (float64 CUDA tensor, zero mean), ks, sigma, kw, kstep, n_grid."""
    import torch
    table = {"C2": (1024, 0.15, 21, 1, 2, 0), "C3": (2048, 0.30, 41, 3, 4, 0), "C5": (8192, 0.30, 41, 5, 6, 4)}
    n, max_grad, ng, su, sn, tiles = table[name]
    n = size or n
    ng = n_grid or ng
    ks = primary_ks(0.05, 7.0, 3)
    u = _smooth_field_device((n, n), max_grad, su, device)
    if tiles:
        rng = np.random.default_rng(su + 100)
        theta = rng.normal(scale=0.01, size=(tiles, tiles))            # extra twist per tile, radians
        t = n // tiles
        pos = torch.arange(n, dtype=torch.float64, device=device)
        idx = torch.clamp((pos / t).floor().long(), max=tiles - 1)
        # smooth-step blend of the per-tile twist over 128 px around every tile boundary
        frac = (pos - idx * t) / t
        edge = 64.0 / t
        wgt = torch.clamp((frac - (1 - edge)) / (2 * edge), 0, 1)       # 0 inside the tile, -> 0.5 at the boundary
        wgt = wgt * wgt * (3 - 2 * wgt)
        nxt = torch.clamp(idx + 1, max=tiles - 1)
        th = torch.from_numpy(theta).to(device)
        tw = (th[idx][:, idx] * (1 - wgt)[:, None] * (1 - wgt)[None, :] + th[nxt][:, idx] * wgt[:, None] * (1 - wgt)[None, :]
              + th[idx][:, nxt] * (1 - wgt)[:, None] * wgt[None, :] + th[nxt][:, nxt] * wgt[:, None] * wgt[None, :])
        cx = (pos - (idx.double() + 0.5) * t)
        u = u.clone()
        u[0] += -tw * cx[None, :]
        u[1] += tw * cx[:, None]
    img = _lattice_device((n, n), ks, u, 0.3, sn, device)
    img -= img.mean()
    kw, kstep = sweep_params(ks, ng)
    return dict(image=img, ks=ks, sigma=10, kw=kw, kstep=kstep, n_grid=ng, name=name)


def frame_series_device(n_frames, device, size=1024, t0=0, total=512):
    """C4: frames t0 .. t0 + n_frames - 1 of a `total`-frame LEEM-like series on `device`:
    the C2 lattice with u_t = u (1 + 0.2 sin(2 pi t / total)) + drift(t), a Gaussian illumination envelope, noise
    seeded per frame, quantised to 16 bit.  Returns (frames (n_frames, size, size) float64 CUDA tensor, ks)."""
    import torch
    ks = primary_ks(0.05, 7.0, 3)
    u = _smooth_field_device((size, size), 0.15, 1, device)
    ax = (torch.arange(size, dtype=torch.float64, device=device) - size / 2) / size
    env = torch.exp(-(ax[:, None] ** 2 + ax[None, :] ** 2) / 0.5)
    frames = torch.empty((n_frames, size, size), dtype=torch.float64, device=device)
    for i in range(n_frames):
        t = t0 + i
        ut = u * (1 + 0.2 * np.sin(2 * np.pi * t / total))
        ut[0] += 2.0 * t / total
        ut[1] -= 1.0 * t / total
        f = (3.0 + _lattice_device((size, size), ks, ut, 0.3, 1000 + t, device)) * env
        f = torch.clamp(f / 8.0, 0, 1)
        frames[i] = torch.round(f * 65535.0)
    return frames, ks
