"""Drop-in for the hot-path functions of ``pyGPA.geometric_phase_analysis`` on B200.

Same names, arguments and return values as the reference (NumPy in, NumPy float64 /
complex128 out); the arithmetic runs in libgpa_b200.so's sm_100a kernels.  Functions of
the reference module that are not on the adaptive-GPA hot path (k-vector extraction,
plotting, property extraction, ...) are intentionally absent — keep using pyGPA for those.
"""
from __future__ import annotations

import numpy as np
import torch

from . import cuGPA as _cu
from . import engine, solvers
from .cuGPA import _sweep, _to_host
from .mathtools import fit_plane, wrapToPi  # noqa: F401  (re-exported like the reference does)

__all__ = ["GPA", "optGPA", "vecGPA", "wfr", "wfr2", "optwfr2", "wfr2_only_lockin", "wfr2_only_lockin_vec",
           "wfr2_grad_opt", "wfr2_grad", "wfr2_grad_vec", "wfr3", "wfr4", "fit_delta_k", "iterate_GPA", "myweighed_lstsq", "reconstruct_u_inv", "reconstruct_u_inv_from_phases",
           "extract_displacement_field", "invert_u", "invert_u_overlap", "undistort_image", "gaussian_deconvolve"]


def optGPA(image, kvec, sigma=22):
    """Spatial lock-in with reference vector ``kvec`` (geometric_phase_analysis.py:48-76)."""
    device = engine.require_cuda()
    img = engine.image_to_device(image, device)
    host = _to_host(engine.lockin_fixed(img, kvec, sigma, out_f64=True))
    torch.cuda.current_stream().synchronize()
    return host.numpy()


def GPA(image, kx, ky, sigma=22):
    """geometric_phase_analysis.py:20-45."""
    return optGPA(image, (kx, ky), sigma)


def vecGPA(image, kvecs, sigma=22):
    """geometric_phase_analysis.py:79-89: lock-in for a list of k-vectors, (K, N, M)."""
    device = engine.require_cuda()
    img = engine.image_to_device(image, device)
    kvecs = np.asarray(kvecs, dtype=np.float64).reshape(-1, 2)
    outs = [_to_host(engine.lockin_fixed(img, k, sigma, out_f64=True)) for k in kvecs]
    torch.cuda.current_stream().synchronize()
    return np.stack([o.numpy() for o in outs])


def wfr2_grad_opt(image, sigma, kx, ky, kw, kstep):
    """Adaptive GPA with the phase gradient (geometric_phase_analysis.py:763-813)."""
    return _sweep(image, sigma, kx, ky, kw, kstep, engine.GRAD_CENTRAL, want_w=True)


def wfr2_grad(image, sigma, kx, ky, kw, kstep, grad=None):
    """geometric_phase_analysis.py:722-760.  The reference computes the gradient of the RE-REFERENCED lock-in of every
    candidate and wraps it per candidate.  grad=None: the winner's value is the same as wfr2_grad_opt's, which is what
    runs (fused).  grad='diff' (np.diff with a NaN appended; note the reference pairs axis 1 with the first gradient
    component here, unlike cuGPA) and a callable run unfused: one fixed lock-in per distinct winning candidate and the
    gradient function on the host (cuGPA._grad_by_winner)."""
    if grad is None:
        return _sweep(image, sigma, kx, ky, kw, kstep, engine.GRAD_CENTRAL, want_w=True)
    if isinstance(grad, str):
        if grad != 'diff':
            raise ValueError("grad must be None, 'diff' or a callable")

        def grad(phase):          # verbatim from the reference (:739-743)
            dbdx = np.diff(phase, axis=1, append=np.nan)
            dbdy = np.diff(phase, axis=0, append=np.nan)
            return np.stack([dbdx, dbdy], axis=-1)
    return _cu._grad_by_winner(image, sigma, kx, ky, kw, kstep, grad, rereference=True)


def optwfr2(image, sigma, kx, ky, kw, kstep):
    """geometric_phase_analysis.py:669-686: dict with 'lockin' and 'w'.  This is the default
    ``wfr_func`` of extract_displacement_field."""
    return _sweep(image, sigma, kx, ky, kw, kstep, engine.GRAD_NONE, want_w=True, want_grad=False)


wfr2 = optwfr2   # geometric_phase_analysis.py:615-644 computes the same thing less efficiently


def wfr2_only_lockin(image, sigma, kx, ky, kw, kstep):
    """geometric_phase_analysis.py:689-702."""
    return _sweep(image, sigma, kx, ky, kw, kstep, engine.GRAD_NONE, want_w=False, want_grad=False)['lockin']


def wfr(image, sigma, kx, ky, kw, kstep):
    """geometric_phase_analysis.py:583-612: dict of real arrays 'wx', 'wy', 'phase', 'r'."""
    g = optwfr2(image, sigma, kx, ky, kw, kstep)
    return {'wx': g['w'][0], 'wy': g['w'][1], 'phase': np.angle(g['lockin']), 'r': np.abs(g['lockin'])}


def wfr3(image, sigma, klist, kref):
    """Adaptive GPA over an explicit ordered list of candidate k-vectors
    (geometric_phase_analysis.py:647-666)."""
    device = engine.require_cuda()
    img = engine.image_to_device(image, device)
    klist = np.asarray(klist, dtype=np.float64).reshape(-1, 2)
    plan = engine.SweepPlan(img.shape, klist[:, 0].copy(), klist[:, 1].copy(), sigma, engine.CAND_LIST, device=device)
    res = plan.run(img, kref, engine.GRAD_NONE, out_f64=True, want_w=True, want_kidx=False)
    host = {k: _to_host(res[k]) for k in ("lockin", "w")}
    torch.cuda.current_stream().synchronize()
    return {k: v.numpy() for k, v in host.items()}


def wfr4(image, sigma, klist, kref, dk):
    """Ordered k-list sweep where a pixel accepts a stronger candidate only if its k lies within
    2 sqrt(2) dk of the k the pixel currently holds (geometric_phase_analysis.py:839-862)."""
    device = engine.require_cuda()
    img = engine.image_to_device(image, device)
    res = engine.wfr4_sweep(img, sigma, klist, kref, dk)
    host = {k: _to_host(res[k]) for k in ("lockin", "w")}
    torch.cuda.current_stream().synchronize()
    return {k: v.numpy() for k, v in host.items()}


def wfr2_only_lockin_vec(image, sigma, kx, ky, kw, kstep):
    """geometric_phase_analysis.py:705-719 batches the candidates of one wx through dask; the
    per-pixel result is that of wfr2_only_lockin, and on the GPU every candidate of a tile is
    already processed in one kernel."""
    return wfr2_only_lockin(image, sigma, kx, ky, kw, kstep)


def wfr2_grad_vec(image, sigma, kx, ky, kw, kstep):
    """geometric_phase_analysis.py:816-836: dask-batched wfr2_grad_opt, same result."""
    return wfr2_grad_opt(image, sigma, kx, ky, kw, kstep)


# ----------------------------------------------------------------------------------------------
# k-vector refinement (K1 fixed lock-in + K2 + K6)
# ----------------------------------------------------------------------------------------------
def fit_delta_k(phases):
    """geometric_phase_analysis.py:92-94: slope of the Huber plane through an unwrapped phase, in cycles."""
    return fit_plane(phases)[:2] / (2 * np.pi)


def iterate_GPA(image, kvecs, sigma, edge=5, iters=3, kmax_iter=25, kmax=200, verbose=False):
    """Iterate the GPA procedure, moving the reference vectors to the extracted average
    (geometric_phase_analysis.py:116-154).  Returns (prs, w, corr): the final unwrapped phases
    (d, N-2 edge, M-2 edge), their weights |lock-in| and the correction with kvecs + corr the k-vectors
    used last.  Everything per pixel stays on the device; only the three plane coefficients per
    k-vector and iteration come back to update the k-vectors."""
    dev = engine.require_cuda()
    img = engine.image_to_device(image, dev)
    kvecs = np.asarray(kvecs, dtype=np.float64)
    corr = np.zeros_like(kvecs)
    for i in range(iters + 1):
        last = i == iters
        prs, ws, deltas = [], [], []
        for ks in kvecs + corr:
            r = engine.lockin_fixed(img, ks, sigma)
            ph, amp, amax = solvers.lockin_phase_amp(r, edge if edge > 0 else 0)
            un = solvers.unwrap(psi=ph, weight=solvers.weight_sqrt_norm(amp, amax), kmax=kmax if last else kmax_iter)
            if last:
                prs.append(un)
                ws.append(amp)
            else:
                deltas.append(solvers.fit_plane_huber(un)[:2] / (2 * np.pi))
        if not last:
            delta_ks = np.stack(deltas)
            if verbose:
                print(delta_ks)
            corr -= delta_ks
    return _host(torch.stack(prs)), _host(torch.stack(ws)), corr


# ----------------------------------------------------------------------------------------------
# phase -> displacement (K3 + K2)
# ----------------------------------------------------------------------------------------------
def _host(t):
    h = _to_host(t)
    torch.cuda.current_stream().synchronize()
    return h.numpy()


def myweighed_lstsq(b, K, w):
    """Per-pixel weighted least squares, minimise |w (K x - b)| (geometric_phase_analysis.py:97-113).
    b (d, n, m), K (d, 2), w (d, >=n, >=m) -> (2, n, m).  Note K is taken as given (the reference's
    callers pass 2 pi kvecs)."""
    dev = engine.require_cuda()
    K = np.asarray(K, dtype=np.float64)
    return _host(solvers.lstsq(solvers.to_device_f64(b, dev), solvers.SRC_PLAIN, K / (2 * np.pi),
                               solvers.to_device_f64(w, dev)))


def reconstruct_u_inv(kvecs, b, weights=None, use_only_ks=None):
    """Reconstruct the displacement field from unwrapped GPA phases
    (geometric_phase_analysis.py:157-193): mean-subtracted phases, then (i) the global
    pseudo-inverse of 2 pi kvecs, (ii) with weights the per-pixel weighted least squares, or
    (iii) with use_only_ks the exact inverse for two chosen k-vectors."""
    dev = engine.require_cuda()
    kvecs = np.asarray(kvecs, dtype=np.float64)
    bd = solvers.to_device_f64(b, dev)
    d = kvecs.shape[0]
    K = 2 * np.pi * kvecs
    if use_only_ks is not None:
        assert len(use_only_ks) == 2
        sel = list(use_only_ks)
        mat = np.zeros((2, d))
        mat[:, sel] = np.linalg.inv(K[sel])
        return _host(solvers.lstsq(bd, solvers.SRC_PLAIN, kvecs, matrix=mat, subtract_mean=True))
    if weights is None:
        if d != 3:   # the reference reshapes to (3, -1) (geometric_phase_analysis.py:185)
            raise ValueError("cannot reshape: the unweighted branch of reconstruct_u_inv needs exactly 3 k-vectors")
        return _host(solvers.lstsq(bd, solvers.SRC_PLAIN, kvecs, matrix=np.linalg.pinv(K), subtract_mean=True))
    return _host(solvers.lstsq(bd, solvers.SRC_PLAIN, kvecs, solvers.to_device_f64(weights, dev), subtract_mean=True))


def reconstruct_u_inv_from_phases(kvecs, phases, weights, weighted_unwrap=True, pre_diff=False):
    """Displacement field from wrapped phases: project the wrapped phase differences onto
    Cartesian displacement gradients per pixel, then integrate by weighted least squares
    (geometric_phase_analysis.py:196-245)."""
    dev = engine.require_cuda()
    u = solvers.displacement_from_phases(np.asarray(kvecs, dtype=np.float64), solvers.to_device_f64(phases, dev),
                                         solvers.to_device_f64(weights, dev), weighted_unwrap, pre_diff)
    return _host(u)


def invert_u_overlap(us, iters=35, edge=0, mode='nearest'):
    """Find the inverse of the displacement us, u_it(r + us(r)) = r, by fixed-point iteration of
    cubic-spline resampling on a grid grown by `edge` (geometric_phase_analysis.py:262-300)."""
    if mode != 'nearest':
        raise NotImplementedError("the B200 Lawler-Fujita kernel implements scipy mode='nearest' (the reference default)")
    dev = engine.require_cuda()
    us = np.asarray(us, dtype=np.float64)
    if us.ndim != 3 or us.shape[0] != 2:
        raise ValueError("us must have shape (2, N, M)")
    return _host(solvers.invert_u(solvers.to_device_f64(us, dev), iters=iters, edge=edge))


def invert_u(us, iters=35, edge=0, mode='nearest'):
    """geometric_phase_analysis.py:248-259: one evaluation of us on the pixel grid, then `iters` fixed-point rounds
    u_it <- us(r - edge + u_it) (the reference subtracts `edge` in the iterations only; reproduced as is)."""
    if mode != 'nearest':
        raise NotImplementedError("the B200 Lawler-Fujita kernel implements scipy mode='nearest' (the reference default)")
    dev = engine.require_cuda()
    us = np.asarray(us, dtype=np.float64)
    if us.ndim != 3 or us.shape[0] != 2:
        raise ValueError("us must have shape (2, N, M)")
    return _host(solvers.invert_u(solvers.to_device_f64(us, dev), iters=iters, edge=edge, overlap=False))


def undistort_image(deformed, u):
    """Reconstruct an undistorted image from a deformed image and the displacement field u
    (Lawler-Fujita; geometric_phase_analysis.py:935-974)."""
    dev = engine.require_cuda()
    deformed = np.asarray(deformed, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    if u.shape != (2,) + deformed.shape:
        raise ValueError("u must have shape (2,) + deformed.shape")
    return _host(solvers.undistort(solvers.to_device_f64(deformed, dev), solvers.to_device_f64(u, dev)))


def gaussian_deconvolve(data, sigma, dr=20, balance=5000):
    """Deconvolve a stack of images `data` (..., N, M) with a Gaussian kernel of std sigma: reflect padding
    by 2 dr, Wiener filter (scikit-image's, with its Laplacian regulariser), crop
    (geometric_phase_analysis.py:892-904)."""
    dev = engine.require_cuda()
    return _host(solvers.gaussian_deconvolve(solvers.to_device_f64(data, dev), sigma, dr, balance))


_DEVICE_SWEEPS = {}


def extract_displacement_field(image, kvecs, sigma=None, kwscale=2.5, ksteps=3, return_gs=False,
                               wfr_func=optwfr2, deconvolve=False):
    """Top-level convenience function (geometric_phase_analysis.py:907-932).

    With one of this package's sweeps as ``wfr_func`` (the default) the whole chain — sweep per
    k-vector, phases and masked weights, per-pixel least squares, PCG integration — stays on the
    GPU and only u comes back.  Any other callable is invoked exactly like the reference does and
    its NumPy results are uploaded for the tail.  deconvolve=True applies gaussian_deconvolve to u
    (on the device as well)."""
    dev = engine.require_cuda()
    image = np.asarray(image, dtype=np.float64)
    kvecs = np.asarray(kvecs, dtype=np.float64)
    norms = np.linalg.norm(kvecs, axis=1)
    kw = norms.mean() / kwscale
    if sigma is None:
        sigma = int(np.ceil(1 / norms.min()))
    kstep = kw / ksteps
    dr = int(2 * sigma)
    centred = image - image.mean()
    mode = _DEVICE_SWEEPS.get(wfr_func)
    if mode is not None and not return_gs:
        img = engine.image_to_device(centred, dev)
        phs, wts = [], []
        for pk in kvecs:
            wxs, wys = engine.grid_axes(pk[0], pk[1], kw, kstep)
            plan = engine.SweepPlan(img.shape, wxs, wys, sigma, device=dev)
            res = plan.run(img, pk, engine.GRAD_NONE, want_kidx=False)
            ph, wt = solvers.phase_weight(res["lockin"], dr)
            phs.append(ph)
            wts.append(wt)
        u = solvers.displacement_from_phases(kvecs, torch.stack(phs), torch.stack(wts))
        if deconvolve:
            u = solvers.gaussian_deconvolve(u, sigma, dr)
        return _host(u)
    gs = [wfr_func(centred, sigma, pk[0], pk[1], kw=kw, kstep=kstep) for pk in kvecs]
    phases = np.stack([np.angle(g['lockin']) for g in gs])
    mask = np.zeros_like(image, dtype=bool)
    mask[dr:-dr, dr:-dr] = 1.
    weights = np.stack([np.abs(g['lockin']) for g in gs]) * (mask + 1e-6)
    u = reconstruct_u_inv_from_phases(kvecs, phases, weights)
    if deconvolve:
        u = gaussian_deconvolve(u, sigma, dr)
    if return_gs:
        return u, gs
    return u


for _f in (optwfr2, wfr2_grad_opt, wfr2_grad, _cu.wfr2_grad_opt, _cu.wfr2_grad_single):
    _DEVICE_SWEEPS[_f] = True
