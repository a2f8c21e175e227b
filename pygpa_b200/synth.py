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
           "sweep_params", "make_config"]


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
