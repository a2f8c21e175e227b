"""Drop-in for the hot-path functions of ``pyGPA.geometric_phase_analysis`` on B200.

Same names, arguments and return values as the reference (NumPy in, NumPy float64 /
complex128 out); the arithmetic runs in libgpa_b200.so's sm_100a kernels.  Functions of
the reference module that are not on the adaptive-GPA hot path (k-vector extraction,
plotting, property extraction, ...) are intentionally absent — keep using pyGPA for those.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from .cuGPA import _sweep, _to_host
from .mathtools import wrapToPi  # noqa: F401  (re-exported like the reference does)

__all__ = ["GPA", "optGPA", "vecGPA", "wfr", "wfr2", "optwfr2", "wfr2_only_lockin", "wfr2_grad_opt",
           "wfr2_grad", "wfr3"]


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
    """geometric_phase_analysis.py:722-760.  The reference computes the gradient of the
    RE-REFERENCED lock-in of every candidate and wraps it per candidate; the winner's value
    is the same as wfr2_grad_opt's, which is what runs here (grad=None only)."""
    if grad is not None:
        raise NotImplementedError("wfr2_grad on B200 supports grad=None only")
    return _sweep(image, sigma, kx, ky, kw, kstep, engine.GRAD_CENTRAL, want_w=True)


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
