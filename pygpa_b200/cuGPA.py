"""Drop-in for ``pyGPA.cuGPA`` (reference: pyGPA/cuGPA.py) on B200.

Same function names, arguments and return values as the CuPy module — NumPy arrays in,
NumPy arrays out (float64 / complex128) — but every device operation is a hand-written
sm_100a kernel of libgpa_b200.so.  Pass these functions wherever the reference accepts a
``wfr_func`` (``extract_displacement_field(..., wfr_func=cuGPA.wfr2_grad_opt)``,
tests/test_cuGPA.py:46-49).

Differences, all deliberate (DESIGN.md "Deviations"):
  * arithmetic is fp32 with a Gaussian truncated at 4.5 sigma instead of complex128 FFTs:
    phases agree to < 1e-3 rad, the selected k-vector is identical except at near-ties;
  * ``grad`` may be None or 'diff'; a callable cannot run inside the fused kernel and
    raises NotImplementedError (there is no CPU fallback);
  * ``cuGPA()`` returns a CuPy array only when CuPy is importable, else a NumPy array.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine

__all__ = ["cuGPA", "wfr2_grad_opt", "wfr2_grad_single", "wfr2_only_lockin", "wfr2_only_grad"]


def _grad_mode(grad):
    if grad is None:
        return engine.GRAD_CENTRAL
    if isinstance(grad, str) and grad == 'diff':
        return engine.GRAD_FORWARD
    raise NotImplementedError(
        "grad must be None (np.gradient) or 'diff'; a user callable cannot be fused into the "
        "CUDA sweep and pygpa_b200 has no CPU fallback")


def _to_host(t):
    """CUDA tensor -> NumPy array through pinned memory (one DMA, no CPU conversion pass)."""
    if t is None:
        return None
    host = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
    host.copy_(t, non_blocking=True)
    return host


_plans = {}


def _plan_for(shape, sigma, kx, ky, kw, kstep, device):
    """SweepPlan cache: the geometry (candidate axes, taps, scratch sizing) of a call depends only on
    these arguments, and building it costs a few ms of host time (cudaMemGetInfo, tap tables)."""
    key = (tuple(shape), float(sigma), float(kx), float(ky), float(kw), float(kstep), str(device))
    plan = _plans.get(key)
    if plan is None:
        if len(_plans) >= 64:
            _plans.clear()
        wxs, wys = engine.grid_axes(kx, ky, kw, kstep)
        plan = engine.SweepPlan(shape, wxs, wys, sigma, engine.CAND_GRID, device=device)
        _plans[key] = plan
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
    return _sweep(image, sigma, kx, ky, kw, kstep, _grad_mode(grad), want_w=True)


def wfr2_grad_single(image, sigma, kx, ky, kw, kstep, grad=None):
    """cuGPA.py:90-133: like wfr2_grad_opt without 'w' (the reference's arithmetic is promoted
    to double by NumPy/CuPy type rules, so the outputs are float64 / complex128 there too)."""
    return _sweep(image, sigma, kx, ky, kw, kstep, _grad_mode(grad), want_w=False)


def wfr2_only_lockin(image, sigma, kvec, kw, kstep):
    """cuGPA.py:136-158: only the complex lock-in signal (note the kvec tuple signature)."""
    kx, ky = kvec
    return _sweep(image, sigma, kx, ky, kw, kstep, engine.GRAD_NONE, want_w=False, want_grad=False)['lockin']


def wfr2_only_grad(image, sigma, kvec, kw, kstep, grad=None):
    """cuGPA.py:161-202: only the phase gradient (N, M, 2)."""
    kx, ky = kvec
    return _sweep(image, sigma, kx, ky, kw, kstep, _grad_mode(grad), want_w=False)['grad']
