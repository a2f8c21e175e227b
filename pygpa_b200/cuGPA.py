"""Drop-in for ``pyGPA.cuGPA`` (reference: pyGPA/cuGPA.py) on B200.

Same function names, arguments and return values as the CuPy module — NumPy arrays in,
NumPy arrays out (float64 / complex128) — but every device operation is a hand-written
sm_100a kernel of libgpa_b200.so.  Pass these functions wherever the reference accepts a
``wfr_func`` (``extract_displacement_field(..., wfr_func=cuGPA.wfr2_grad_opt)``,
tests/test_cuGPA.py:46-49).

Differences, all deliberate (DESIGN.md "Deviations"):
  * arithmetic is fp32 with a Gaussian truncated at 4.5 sigma instead of complex128 FFTs:
    phases agree to < 1e-3 rad, the selected k-vector is identical except at near-ties;
  * ``grad`` None and 'diff' run inside the fused kernel; a callable gets NumPy arrays (there is no CuPy here) and is
    applied, unfused, to the phase image of every distinct winning candidate (see _grad_by_winner);
  * ``cuGPA()`` returns a CuPy array only when CuPy is importable, else a NumPy array.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine

__all__ = ["cuGPA", "wfr2_grad_opt", "wfr2_grad_single", "wfr2_only_lockin", "wfr2_only_grad",
           "wfr2_grad_opt_peaks", "clear_plans"]


def _grad_mode(grad):
    if grad is None:
        return engine.GRAD_CENTRAL
    if isinstance(grad, str) and grad == 'diff':
        return engine.GRAD_FORWARD
    if callable(grad):
        return None          # unfused: _grad_by_winner
    raise ValueError("grad must be None (np.gradient), 'diff' or a callable")


def _grad_by_winner(image, sigma, kx, ky, kw, kstep, grad_func, rereference, want_w=True):
    """Sweep with a user gradient function.  The reference applies grad_func to the phase image of EVERY candidate and
    keeps, per pixel, the value of the last candidate that won (cuGPA.py:66-82, geometric_phase_analysis.py:745-757) —
    i.e. of the final winner.  A callable cannot run inside the fused kernel, so: arg-max sweep on the GPU, then ONE
    fixed lock-in (gpa_lockin_fixed) per DISTINCT winning candidate, grad_func on its phase image on the host, and
    the pixels that candidate won are filled.  rereference: True = geometric_phase_analysis.wfr2_grad (gradient of the
    re-referenced phase, wrapped per candidate), False = cuGPA.wfr2_grad_opt (raw phase + 2 pi (w - k), wrapped at
    the end).  grad_func gets a NumPy float64 array (CuPy does not exist here) and returns an (N, M, 2) array or a
    pair of (N, M) arrays."""
    device = engine.require_cuda()
    img = engine.image_to_device(image, device)
    plan = _plan_for(img.shape, sigma, kx, ky, kw, kstep, device)
    res = plan.run(img, (kx, ky), engine.GRAD_NONE, out_f64=True, want_w=want_w, want_kidx=True)
    host = {k: _to_host(res[k]) for k in ("lockin", "w", "kidx") if res.get(k) is not None}
    torch.cuda.current_stream().synchronize()
    out = {k: v.numpy() for k, v in host.items()}
    kidx = out.pop("kidx")
    n, m = kidx.shape
    wxs, wys = plan.wx, plan.wy
    xx, yy = np.ogrid[0:n, 0:m]
    grad = np.zeros((n, m, 2))
    two_pi = 2 * np.pi

    def wrap(v):
        return (v + np.pi) % two_pi - np.pi
    for c in np.unique(kidx[kidx >= 0]):
        wx, wy = float(wxs[c // wys.size]), float(wys[c % wys.size])
        sf_t = _to_host(engine.lockin_fixed(img, (wx, wy), sigma, out_f64=True))
        torch.cuda.current_stream().synchronize()
        sf = sf_t.numpy()
        if rereference:
            sf = sf * np.exp(-2j * np.pi * ((wx - kx) * xx + (wy - ky) * yy))
        g = grad_func(-np.angle(sf))
        g = np.stack(g, axis=-1) if isinstance(g, (tuple, list)) else np.asarray(g)
        if rereference:
            g = wrap(g * 2) / 2
        else:
            g = g + two_pi * np.array([wx - kx, wy - ky])
        t = kidx == c
        grad[t] = g[t]
    out["grad"] = grad if rereference else wrap(2 * grad) / 2
    return out


def _to_host(t):
    """CUDA tensor -> NumPy array through pinned memory (one DMA, no CPU conversion pass)."""
    if t is None:
        return None
    host = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
    host.copy_(t, non_blocking=True)
    return host


# The reference is stateless.  The only state kept here is derived, never results: the
# geometry of a call (candidate axes, tap tables, scratch sizing), which costs a few ms of host time to build.
# The cache is a small LRU (8 entries; the scratch itself is one shared grow-only buffer per device, see
# engine.workspace) and clear_plans() drops everything, scratch included.
_PLAN_CACHE_SIZE = 8
_plans = {}
_host_sweeps = {}


def clear_plans():
    """Drop every cached plan, batched executor and the device scratch buffers."""
    _plans.clear()
    for hs in _host_sweeps.values():
        hs.close()
    _host_sweeps.clear()
    engine.release_workspaces()


def _plan_for(shape, sigma, kx, ky, kw, kstep, device):
    key = (tuple(shape), float(sigma), float(kx), float(ky), float(kw), float(kstep), str(device))
    plan = _plans.pop(key, None)
    if plan is None:
        while len(_plans) >= _PLAN_CACHE_SIZE:
            _plans.pop(next(iter(_plans)))          # least recently used
        wxs, wys = engine.grid_axes(kx, ky, kw, kstep)
        plan = engine.SweepPlan(shape, wxs, wys, sigma, engine.CAND_GRID, device=device)
    _plans[key] = plan                              # most recently used last
    return plan


def _sweep(image, sigma, kx, ky, kw, kstep, grad_mode, want_w, want_grad=True):
    device = engine.require_cuda()
    img = engine.image_to_device(image, device)
    plan = _plan_for(img.shape, sigma, kx, ky, kw, kstep, device)
    res = plan.run(img, (kx, ky), grad_mode if want_grad else engine.GRAD_NONE, out_f64=True,
                   want_w=want_w, want_kidx=False)
    host = {k: _to_host(res[k]) for k in ("lockin", "w", "grad") if res.get(k) is not None}
    torch.cuda.current_stream().synchronize()
    return {k: v.numpy() for k, v in host.items()}


def cuGPA(image, kvec, sigma=22):
    """Spatial lock-in with a fixed reference vector (cuGPA.py:11-38).  The reference returns
    the result as a device (CuPy) array; so does this when CuPy is available."""
    device = engine.require_cuda()
    img = engine.image_to_device(image, device)
    res = engine.lockin_fixed(img, kvec, sigma, out_f64=True)
    try:
        import cupy as cp   # noqa: WPS433
        return cp.asarray(res)
    except ImportError:
        host = _to_host(res)
        torch.cuda.current_stream().synchronize()
        return host.numpy()


def wfr2_grad_opt(image, sigma, kx, ky, kw, kstep, grad=None):
    """Adaptive GPA: dict with 'lockin' (N,M) c16, 'w' (2,N,M) f8, 'grad' (N,M,2) f8
    (cuGPA.py:41-87)."""
    if _grad_mode(grad) is None:
        return _grad_by_winner(image, sigma, kx, ky, kw, kstep, grad, rereference=False)
    return _sweep(image, sigma, kx, ky, kw, kstep, _grad_mode(grad), want_w=True)


def wfr2_grad_single(image, sigma, kx, ky, kw, kstep, grad=None):
    """cuGPA.py:90-133: like wfr2_grad_opt without 'w' (the reference's arithmetic is promoted
    to double by NumPy/CuPy type rules, so the outputs are float64 / complex128 there too)."""
    if _grad_mode(grad) is None:
        return _grad_by_winner(image, sigma, kx, ky, kw, kstep, grad, rereference=False, want_w=False)
    return _sweep(image, sigma, kx, ky, kw, kstep, _grad_mode(grad), want_w=False)


def wfr2_only_lockin(image, sigma, kvec, kw, kstep):
    """cuGPA.py:136-158: only the complex lock-in signal (note the kvec tuple signature)."""
    kx, ky = kvec
    return _sweep(image, sigma, kx, ky, kw, kstep, engine.GRAD_NONE, want_w=False, want_grad=False)['lockin']


def wfr2_only_grad(image, sigma, kvec, kw, kstep, grad=None):
    """cuGPA.py:161-202: only the phase gradient (N, M, 2)."""
    kx, ky = kvec
    if _grad_mode(grad) is None:
        return _grad_by_winner(image, sigma, kx, ky, kw, kstep, grad, rereference=False, want_w=False)['grad']
    return _sweep(image, sigma, kx, ky, kw, kstep, _grad_mode(grad), want_w=False)['grad']


def wfr2_grad_opt_peaks(image, sigma, kvecs, kw, kstep, grad=None, shape=None):
    """wfr2_grad_opt for SEVERAL primary k-vectors of one frame in one call — the loop of
    extract_displacement_field (geometric_phase_analysis.py:916-921) as a batch: returns
    [wfr2_grad_opt(image, sigma, k[0], k[1], kw, kstep, grad) for k in kvecs] (same dicts, same dtypes), but the
    image is uploaded once and the device-to-host copy of peak p's arrays overlaps the sweep of peak p+1.

    With torch.distributed initialised (one process per GPU) the call is SPMD: every rank calls it, the k-grid is
    sharded over the GPUs, every GPU copies its rows of the results out over its own PCIe link, and rank 0 gets the
    arrays (the other ranks pass image=None and shape=<frame shape>, and get None).  The returned arrays are views of a page-locked buffer
    owned by the cached executor: valid until the next call with the same geometry (copy them to keep them)."""
    import torch.distributed as tdist
    from . import dist as gdist
    device = engine.require_cuda()
    rank = tdist.get_rank() if tdist.is_initialized() else 0
    kvecs = [tuple(map(float, k)) for k in kvecs]
    if rank == 0:
        shape = tuple(np.shape(image))
    elif shape is None:
        raise ValueError("ranks other than 0 must pass shape=(N, M)")
    shape = tuple(int(v) for v in shape)
    key = (shape, float(sigma), tuple(kvecs), float(kw), float(kstep), str(grad), str(device))
    hs = _host_sweeps.get(key)
    if hs is None:
        for old in _host_sweeps.values():           # one batched executor at a time: it pins 48 B per pixel and peak
            old.close()
        _host_sweeps.clear()
        hs = gdist.HostSweep(shape, sigma, kvecs, kw, kstep, grad=grad)
        _host_sweeps[key] = hs
    return hs(image if rank == 0 else None)
