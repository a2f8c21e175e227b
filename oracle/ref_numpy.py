"""NumPy/SciPy restatement of the pyGPA hot path (oracle; test infrastructure only).

Every function names the reference lines it follows (paths relative to the
reference checkout, ``pyGPA/...``).  Conventions are the reference's: images are
(N, M) with axis 0 = "x"; k-vectors are in cycles/pixel and ``k[0]`` multiplies
the axis-0 index; all arithmetic is float64 / complex128.
"""
from __future__ import annotations

import numpy as np
import scipy.ndimage as ndi
from scipy.fft import dctn, idctn

TWO_PI = 2.0 * np.pi

__all__ = [
    "wrap_to_pi", "gaussian_transfer", "lockin_fixed", "candidate_axes", "wfr_sweep",
    "wfr_sweep_klist", "wfr2_grad", "wfr4", "wfr4_allowed", "fit_plane", "iterate_GPA", "phase_unwrap", "phase_unwrap_prediff", "weighted_lstsq",
    "reconstruct_u_inv", "reconstruct_u_inv_from_phases", "invert_u", "invert_u_overlap",
    "undistort_image", "extract_displacement_field", "fixed_reference_pipeline",
]


def wrap_to_pi(x):
    """mathtools.py:72-75 / phase_unwrap.py:135-138 — values mapped to [-pi, pi)."""
    return np.mod(np.asarray(x) + np.pi, TWO_PI) - np.pi


def gaussian_transfer(shape, sigma):
    """The frequency response ``scipy.ndimage.fourier_gaussian`` applies
    (geometric_phase_analysis.py:44,75,87; cuGPA.py:57)."""
    return ndi.fourier_gaussian(np.ones(shape), sigma=sigma)


def _carrier(shape, wx, wy):
    # geometric_phase_analysis.py:72-73: exp(2 pi i (x*wx + y*wy)), x = axis 0.
    x = np.arange(shape[0])[:, None]
    y = np.arange(shape[1])[None, :]
    return np.exp(TWO_PI * 1j * (x * wx + y * wy))


def lockin_fixed(image, kvec, sigma=22, transfer=None):
    """Fixed-reference spatial lock-in.

    geometric_phase_analysis.py:20-45 (GPA), 48-76 (optGPA); cuGPA.py:11-38.
    demodulate -> FFT -> Gaussian low-pass in the Fourier domain -> inverse FFT.
    """
    image = np.asarray(image)
    if transfer is None:
        transfer = gaussian_transfer(image.shape, sigma)
    spectrum = np.fft.fft2(image * _carrier(image.shape, kvec[0], kvec[1]))
    return np.fft.ifft2(spectrum * transfer)


def candidate_axes(kx, ky, kw, kstep):
    """The candidate grid of the adaptive sweep, geometric_phase_analysis.py:803-804.
    The two ``np.arange`` calls are reproduced verbatim because their length is
    float-rounding dependent."""
    return np.arange(kx - kw, kx + kw, kstep), np.arange(ky - kw, ky + kw, kstep)


def _phase_gradient(sf, mode=None):
    """Gradient of -angle(sf) as wfr2_grad_opt takes it.

    mode None  -> np.gradient (geometric_phase_analysis.py:807, cuGPA.py:63-65)
    mode 'diff'-> forward difference padded with NaN (cuGPA.py:58-62: axis 0 first).
    Returns (..., 2) with [...,0] = d/d axis0.
    """
    ph = -np.angle(sf)
    if mode is None:
        g0, g1 = np.gradient(ph)
    elif mode == 'diff':
        g0 = np.diff(ph, axis=0, append=np.nan)
        g1 = np.diff(ph, axis=1, append=np.nan)
    else:
        raise ValueError("grad mode must be None or 'diff'")
    return np.stack([g0, g1], axis=-1)


def wfr_sweep_klist(image, sigma, klist, kref, grad_mode=None, want_grad=True,
                    return_diag=False):
    """Adaptive GPA over an ordered list of candidate k-vectors.

    Semantics of wfr2_grad_opt (geometric_phase_analysis.py:763-813; cuGPA.py:41-87),
    and with want_grad=False of optwfr2 / wfr3 (669-686, 647-666):
      for each candidate in order: sf = lock-in at that k; pixels where |sf| is
      STRICTLY larger than |stored lock-in| take the candidate: store sf re-referenced
      to kref, the candidate k, and grad(-angle(sf)) + 2 pi (k - kref).
      Finally grad <- wrapToPi(2 grad)/2.

    Returns dict(lockin (N,M) c16, w (2,N,M), grad (N,M,2), kidx (N,M) int32 [-1 where
    nothing ever won]); with return_diag also 'amp1','amp2' = largest and second
    largest candidate amplitude per pixel (near-tie classification).
    """
    image = np.asarray(image, dtype=np.float64)
    klist = np.asarray(klist, dtype=np.float64).reshape(-1, 2)
    shape = image.shape
    transfer = gaussian_transfer(shape, sigma)
    x = np.arange(shape[0])[:, None]
    y = np.arange(shape[1])[None, :]
    lockin = np.zeros(shape, dtype=np.complex128)
    w = np.zeros(shape + (2,))
    grad = np.zeros(shape + (2,))
    kidx = np.full(shape, -1, dtype=np.int32)
    amp1 = np.zeros(shape)
    amp2 = np.zeros(shape)
    for idx, (wx, wy) in enumerate(klist):
        sf = np.fft.ifft2(np.fft.fft2(image * np.exp(TWO_PI * 1j * (x * wx + y * wy))) * transfer)
        a = np.abs(sf)
        take = a > np.abs(lockin)
        if want_grad:
            g = _phase_gradient(sf, grad_mode)
            grad[take] = g[take] + TWO_PI * np.array([wx - kref[0], wy - kref[1]])
        rot = np.exp(-TWO_PI * 1j * ((wx - kref[0]) * x + (wy - kref[1]) * y))
        lockin[take] = (sf * rot)[take]
        w[take] = (wx, wy)
        kidx[take] = idx
        if return_diag:
            amp2 = np.maximum(amp2, np.minimum(amp1, a))
            amp1 = np.maximum(amp1, a)
    out = {'lockin': lockin, 'w': np.moveaxis(w, -1, 0), 'kidx': kidx}
    if want_grad:
        out['grad'] = wrap_to_pi(2 * grad) / 2
    if return_diag:
        out['amp1'], out['amp2'] = amp1, amp2
    return out


def wfr_sweep(image, sigma, kx, ky, kw, kstep, grad_mode=None, want_grad=True,
              return_diag=False):
    """wfr2_grad_opt / optwfr2 on the Cartesian grid, wx outer, wy inner
    (geometric_phase_analysis.py:803-804).  kidx = ix*ny + iy."""
    wxs, wys = candidate_axes(kx, ky, kw, kstep)
    klist = np.stack(np.meshgrid(wxs, wys, indexing='ij'), axis=-1).reshape(-1, 2)
    out = wfr_sweep_klist(image, sigma, klist, (kx, ky), grad_mode, want_grad, return_diag)
    out['wxs'], out['wys'] = wxs, wys
    return out


def wfr2_grad(image, sigma, kx, ky, kw, kstep, grad=None):
    """geometric_phase_analysis.py:722-760: like wfr_sweep, but the gradient function (None = np.gradient, 'diff' =
    np.diff with a NaN appended — axis 1 first there, :739-743 — or a callable) is applied to the phase of every
    candidate's RE-REFERENCED lock-in and wrapped per candidate.  Returns dict(lockin, w (2,N,M), grad, kidx)."""
    image = np.asarray(image, dtype=np.float64)
    shape = image.shape
    transfer = gaussian_transfer(shape, sigma)
    x = np.arange(shape[0])[:, None]
    y = np.arange(shape[1])[None, :]
    if isinstance(grad, str) and grad == 'diff':
        def grad_func(phase):
            return np.stack([np.diff(phase, axis=1, append=np.nan), np.diff(phase, axis=0, append=np.nan)], axis=-1)
    elif grad is None:
        def grad_func(phase):
            return np.stack(np.gradient(phase), axis=-1)
    else:
        grad_func = grad
    lockin = np.zeros(shape, dtype=np.complex128)
    w = np.zeros(shape + (2,))
    g_out = np.zeros(shape + (2,))
    kidx = np.full(shape, -1, dtype=np.int32)
    wxs, wys = candidate_axes(kx, ky, kw, kstep)
    for ix, wx in enumerate(wxs):
        for iy, wy in enumerate(wys):
            sf = np.fft.ifft2(np.fft.fft2(image * np.exp(TWO_PI * 1j * (x * wx + y * wy))) * transfer)
            sf = sf * np.exp(-TWO_PI * 1j * ((wx - kx) * x + (wy - ky) * y))
            g = wrap_to_pi(grad_func(-np.angle(sf)) * 2) / 2
            take = np.abs(sf) > np.abs(lockin)
            lockin[take] = sf[take]
            w[take] = (wx, wy)
            g_out[take] = g[take]
            kidx[take] = ix * len(wys) + iy
    return {'lockin': lockin, 'w': np.moveaxis(w, -1, 0), 'grad': g_out, 'kidx': kidx}


def wfr4_allowed(klist, dk):
    """The neighbourhood test of wfr4 (geometric_phase_analysis.py:854) as a (K, K) boolean table:
    allowed[c, i] = ``np.linalg.norm(klist[c] - klist[i]) < 2*np.sqrt(2)*dk`` (c = the k-vector a
    pixel currently holds, i = the new candidate), same float64 expression as the reference."""
    klist = np.asarray(klist, dtype=np.float64).reshape(-1, 2)
    return np.linalg.norm(klist[:, None, :] - klist[None, :, :], axis=-1) < 2 * np.sqrt(2) * dk


def wfr4(image, sigma, klist, kref, dk, return_diag=False):
    """geometric_phase_analysis.py:839-862: ordered sweep in which a pixel only accepts a new
    candidate if its amplitude is larger AND its k lies within 2 sqrt(2) dk of the k the pixel
    currently holds (initially klist[0]).  Sequential per pixel, order dependent.

    return_diag adds 'kidx' (index of the held candidate, -1 = never accepted) and 'margin': the
    smallest relative amplitude gap |a - |stored|| / max(a, |stored|) over all amplitude decisions
    that mattered (candidates inside the neighbourhood) — a pixel with a small margin is a
    near-tie and may legitimately follow another path in a different arithmetic."""
    image = np.asarray(image, dtype=np.float64)
    klist = np.asarray(klist, dtype=np.float64).reshape(-1, 2)
    shape = image.shape
    transfer = gaussian_transfer(shape, sigma)
    x = np.arange(shape[0])[:, None]
    y = np.arange(shape[1])[None, :]
    lockin = np.zeros(shape, dtype=np.complex128)
    w = np.zeros(shape + (2,))
    w[..., 0], w[..., 1] = klist[0, 0], klist[0, 1]
    kidx = np.full(shape, -1, dtype=np.int32)
    margin = np.full(shape, np.inf)
    for idx, (wx, wy) in enumerate(klist):
        sf = np.fft.ifft2(np.fft.fft2(image * np.exp(TWO_PI * 1j * (x * wx + y * wy))) * transfer)
        sf = sf * np.exp(-TWO_PI * 1j * ((wx - kref[0]) * x + (wy - kref[1]) * y))
        a, stored = np.abs(sf), np.abs(lockin)
        near = np.linalg.norm(w - np.array([wx, wy]), axis=-1) < 2 * np.sqrt(2) * dk
        take = (a > stored) & near
        if return_diag:
            gap = np.abs(a - stored) / np.maximum(np.maximum(a, stored), 1e-300)
            margin = np.where(near, np.minimum(margin, gap), margin)
        lockin[take] = sf[take]
        w[take] = (wx, wy)
        kidx[take] = idx
    out = {'lockin': lockin, 'w': np.moveaxis(w, -1, 0)}
    if return_diag:
        out['kidx'], out['margin'] = kidx, margin
    return out


# ----------------------------------------------------------------------------------
# weighted least-squares phase unwrapping (Ghiglia-Romero PCG), phase_unwrap.py
# ----------------------------------------------------------------------------------

def _poisson_scale(shape):
    """phase_unwrap.py:106-115.  Note the reference divides the axis-0 index by M and
    the axis-1 index by N; reproduced as is."""
    n, m = shape
    i = np.arange(n)[:, None]
    j = np.arange(m)[None, :]
    scale = 2.0 * (np.cos(np.pi * i / m) + np.cos(np.pi * j / n) - 2.0)
    scale[0, 0] = 1.0
    return scale


def _apply_q(p, wwx, wwy):
    """phase_unwrap.py:118-132: A^T W^T W A p (weighted 5-point Laplacian)."""
    fx = wwx * np.diff(p, axis=1)
    fy = wwy * np.diff(p, axis=0)
    return np.diff(fx, axis=1, prepend=0, append=0) + np.diff(fy, axis=0, prepend=0, append=0)


def _pcg(dx, dy, wwx, wwy, kmax, return_iters=False):
    """Shared PCG loop of phase_unwrap (phase_unwrap.py:168-208) and
    phase_unwrap_prediff (311-350).  dx: (N, M-1) wrapped diffs along axis 1,
    dy: (N-1, M) along axis 0, wwx/wwy matching edge weights (or None for 1)."""
    fx = dx if wwx is None else wwx * dx
    fy = dy if wwy is None else wwy * dy
    if wwx is None:
        wwx = np.ones_like(dx)
        wwy = np.ones_like(dy)
    r = np.diff(fx, axis=1, prepend=0, append=0) + np.diff(fy, axis=0, prepend=0, append=0)
    r0 = np.linalg.norm(r)
    phi = np.zeros((dx.shape[0], dy.shape[1]))
    scale = _poisson_scale(r.shape)
    k = 0
    rz_prev = None
    p = None
    while not np.all(r == 0.0):
        z = idctn(dctn(r) / scale)
        k += 1
        rz = np.tensordot(r, z)
        p = z if k == 1 else z + (rz / rz_prev) * p
        rz_prev = rz
        qp = _apply_q(p, wwx, wwy)
        alpha = rz / np.tensordot(p, qp)
        phi += alpha * p
        r -= alpha * qp
        if k >= kmax or np.linalg.norm(r) < 1e-9 * r0:
            break
    return (phi, k) if return_iters else phi


def _edge_weights(weight):
    ww = np.asarray(weight, dtype=np.float64) ** 2           # phase_unwrap.py:162
    return (np.minimum(ww[:, :-1], ww[:, 1:]),               # :166
            np.minimum(ww[:-1, :], ww[1:, :]))               # :167


def phase_unwrap(psi, weight=None, kmax=100, return_iters=False):
    """phase_unwrap.py:141-208."""
    psi = np.asarray(psi, dtype=np.float64)
    dx = wrap_to_pi(np.diff(psi, axis=1))
    dy = wrap_to_pi(np.diff(psi, axis=0))
    if weight is None:
        wwx = np.ones_like(dx)
        wwy = np.ones_like(dy)
    else:
        wwx, wwy = _edge_weights(weight)
    return _pcg(dx, dy, wwx, wwy, kmax, return_iters)


def phase_unwrap_prediff(dx, dy, weight=None, kmax=100, return_iters=False):
    """phase_unwrap.py:282-350 (gradients supplied by the caller)."""
    dx = wrap_to_pi(np.asarray(dx, dtype=np.float64))
    dy = wrap_to_pi(np.asarray(dy, dtype=np.float64))
    if weight is None:
        return _pcg(dx, dy, None, None, kmax, return_iters)
    wwx, wwy = _edge_weights(weight)
    return _pcg(dx, dy, wwx, wwy, kmax, return_iters)


# ----------------------------------------------------------------------------------
# phase -> displacement
# ----------------------------------------------------------------------------------

def weighted_lstsq(b, K, w):
    """myweighed_lstsq, geometric_phase_analysis.py:97-113: per pixel (i, j) the
    minimum-norm least-squares solution of (w[:, i, j, None] * K) x = w[:, i, j] * b[:, i, j]
    (LAPACK gelsd with rcond = machine eps, as numba's np.linalg.lstsq calls it).
    w may be larger than b in its trailing dims (the reference indexes it with b's
    indices)."""
    b = np.asarray(b, dtype=np.float64)
    K = np.asarray(K, dtype=np.float64)
    d, n, m = b.shape
    wl = np.asarray(w, dtype=np.float64)[:, :n, :m]
    a = wl.reshape(d, -1).T[:, :, None] * K[None, :, :]            # (P, d, 2)
    rhs = (wl * b).reshape(d, -1).T[:, :, None]                      # (P, d, 1)
    sol = np.linalg.pinv(a, rcond=np.finfo(np.float64).eps) @ rhs    # (P, 2, 1)
    return sol[:, :, 0].T.reshape(2, n, m)


def reconstruct_u_inv(kvecs, b, weights=None, use_only_ks=None):
    """geometric_phase_analysis.py:157-193."""
    K = TWO_PI * np.asarray(kvecs, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    b = b - b.mean(axis=(1, 2), keepdims=True)                       # :182
    if use_only_ks is not None:
        assert len(use_only_ks) == 2                                  # :190
        sel = list(use_only_ks)
        u = np.linalg.inv(K[sel]) @ b[sel].reshape(2, -1)
        return u.reshape((2,) + b.shape[1:])
    if weights is None:
        # :185 hard-codes three k-vectors through reshape((3, -1))
        u = np.linalg.lstsq(K, b.reshape(3, -1), rcond=None)[0]
        return u.reshape((2,) + b.shape[1:])
    return weighted_lstsq(b, K, weights)


def reconstruct_u_inv_from_phases(kvecs, phases, weights, weighted_unwrap=True,
                                  pre_diff=False):
    """geometric_phase_analysis.py:196-245."""
    K = TWO_PI * np.asarray(kvecs, dtype=np.float64)
    phases = np.asarray(phases, dtype=np.float64)
    if pre_diff:
        dbdx = wrap_to_pi(phases[..., 0])[:, :, :-1]
        dbdy = wrap_to_pi(phases[..., 1])[:, :-1]
    else:
        dbdx = wrap_to_pi(np.diff(phases, axis=2))
        dbdy = wrap_to_pi(np.diff(phases, axis=1))
    dudx = weighted_lstsq(dbdx, K, weights)
    dudy = weighted_lstsq(dbdy, K, weights)
    if weighted_unwrap:
        wn = np.linalg.norm(weights, axis=0)
        us = [phase_unwrap_prediff(dudx[i], dudy[i], wn, kmax=10) for i in range(2)]
    else:
        us = [phase_unwrap_prediff(dudx[i], dudy[i]) for i in range(2)]
    return np.array(us)


def invert_u_overlap(us, iters=35, edge=0, mode='nearest'):
    """geometric_phase_analysis.py:262-300: fixed-point inversion of the displacement,
    u_it <- u(r + u_it), cubic-spline resampling (scipy map_coordinates, order 3)."""
    us = np.asarray(us, dtype=np.float64)
    n, m = us.shape[1:]
    gx, gy = np.mgrid[-edge:n + edge, -edge:m + edge]
    cur = [ndi.map_coordinates(c, [gx, gy], mode=mode) for c in us]
    for _ in range(iters):      # iters-1 plain rounds + the final one (cval is inert)
        cur = [ndi.map_coordinates(c, [gx + cur[0], gy + cur[1]], mode=mode) for c in us]
    return np.stack(cur)


def invert_u(us, iters=35, edge=0, mode='nearest'):
    """geometric_phase_analysis.py:248-259: the variant on the (N, M) grid; `- edge` enters the iterations only."""
    us = np.asarray(us, dtype=np.float64)
    gx, gy = np.mgrid[:us.shape[1], :us.shape[2]]
    cur = [ndi.map_coordinates(c, [gx, gy], mode=mode) for c in us]
    for _ in range(iters):
        cur = [ndi.map_coordinates(c, [gx + cur[0] - edge, gy + cur[1] - edge], mode=mode) for c in us]
    return np.stack(cur)


def undistort_image(deformed, u):
    """geometric_phase_analysis.py:935-974 (Lawler-Fujita)."""
    u = np.asarray(u, dtype=np.float64)
    u_inv = invert_u_overlap(-u)
    gx, gy = np.mgrid[:u.shape[1], :u.shape[2]]
    return ndi.map_coordinates(np.asarray(deformed, dtype=np.float64),
                               [gx + u_inv[0], gy + u_inv[1]])


def extract_displacement_field(image, kvecs, sigma=None, kwscale=2.5, ksteps=3,
                               return_gs=False, sweep=None):
    """geometric_phase_analysis.py:907-932 without the skimage deconvolution branch.
    ``sweep(image, sigma, kx, ky, kw=, kstep=)`` defaults to the oracle's optwfr2."""
    image = np.asarray(image, dtype=np.float64)
    kvecs = np.asarray(kvecs, dtype=np.float64)
    norms = np.linalg.norm(kvecs, axis=1)
    kw = norms.mean() / kwscale
    if sigma is None:
        sigma = int(np.ceil(1 / norms.min()))
    kstep = kw / ksteps
    if sweep is None:
        def sweep(im, s, kx, ky, kw, kstep):
            return wfr_sweep(im, s, kx, ky, kw, kstep, want_grad=False)
    gs = [sweep(image - image.mean(), sigma, pk[0], pk[1], kw=kw, kstep=kstep) for pk in kvecs]
    phases = np.stack([np.angle(g['lockin']) for g in gs])
    mask = np.zeros(image.shape, dtype=bool)
    dr = 2 * sigma
    mask[dr:-dr, dr:-dr] = True
    weights = np.stack([np.abs(g['lockin']) for g in gs]) * (mask + 1e-6)
    u = reconstruct_u_inv_from_phases(kvecs, phases, weights)
    return (u, gs) if return_gs else u


def fit_plane(image):
    """mathtools.py:30-47: plane a[0] x + a[1] y + a[2] through `image`, Huber loss with scipy's
    default f_scale = 1, minimised by scipy.optimize.least_squares from a zero start."""
    import scipy.optimize as spo
    image = np.asarray(image, dtype=np.float64)
    xx, yy = np.meshgrid(np.arange(image.shape[0]), np.arange(image.shape[1]), indexing='ij')

    def resid(x):
        return (image - (x[0] * xx + x[1] * yy + x[2])).ravel()
    return spo.least_squares(resid, np.zeros(3), loss='huber').x


def iterate_GPA(image, kvecs, sigma, edge=5, iters=3, kmax_iter=25, kmax=200):
    """geometric_phase_analysis.py:116-154."""
    kvecs = np.asarray(kvecs, dtype=np.float64)
    corr = np.zeros_like(kvecs)
    sl = (slice(edge, -edge), slice(edge, -edge)) if edge > 0 else (slice(None), slice(None))
    for i in range(iters + 1):
        rs = [lockin_fixed(image, k, sigma) for k in kvecs + corr]
        prs = [np.angle(r)[sl] for r in rs]
        w = np.stack([np.abs(r)[sl] for r in rs])
        if i < iters:
            prs = [phase_unwrap(p, np.sqrt(we / we.max()), kmax=kmax_iter) for p, we in zip(prs, w)]
            corr -= np.stack([fit_plane(p)[:2] / TWO_PI for p in prs])
        else:
            prs = np.stack([phase_unwrap(p, np.sqrt(we / we.max()), kmax=kmax) for p, we in zip(prs, w)])
    return prs, w, corr


def fixed_reference_pipeline(image, kvecs, sigma, kmax=100, weighted=True):
    """Config-1 path as iterate_GPA assembles it (geometric_phase_analysis.py:133-151):
    lock-in per k -> angle/abs -> weighted unwrap with sqrt(w/max w) -> reconstruct_u_inv."""
    rs = [lockin_fixed(image, k, sigma) for k in kvecs]
    amps = np.stack([np.abs(r) for r in rs])
    phases = np.stack([phase_unwrap(np.angle(r), np.sqrt(a / a.max()), kmax=kmax)
                       for r, a in zip(rs, amps)])
    u = reconstruct_u_inv(kvecs, phases, amps if weighted else None)
    return {'lockin': np.stack(rs), 'phases': phases, 'u': u}
