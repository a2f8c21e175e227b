"""Drop-in for ``pyGPA.phase_unwrap`` (reference: pyGPA/phase_unwrap.py) on B200: weighted
least-squares phase unwrapping (Ghiglia & Romero PCG with a DCT Poisson preconditioner),
float64 on the device."""
from __future__ import annotations

import numpy as np
import torch

from . import solvers
from .cuGPA import _to_host
from .engine import require_cuda

__all__ = ["phase_unwrap", "phase_unwrap_prediff", "phase_unwrap_ref", "phase_unwrap_ref_prediff"]


def _finish(t):
    host = _to_host(t)
    torch.cuda.current_stream().synchronize()
    return host.numpy()


def phase_unwrap(psi, weight=None, kmax=100):
    """Unwrap the phase image psi, optionally weighted (phase_unwrap.py:141-208)."""
    dev = require_cuda()
    psi = np.asarray(psi)
    if psi.ndim != 2:
        raise ValueError("psi must be 2-D")
    w = None if weight is None else solvers.to_device_f64(np.broadcast_to(np.asarray(weight, dtype=np.float64), psi.shape), dev)
    return _finish(solvers.unwrap(psi=solvers.to_device_f64(psi, dev), weight=w, kmax=kmax))


def phase_unwrap_prediff(dx, dy, weight=None, kmax=100):
    """Unwrap from phase gradients: dx (N, M-1) = diff along axis 1, dy (N-1, M) along axis 0
    (phase_unwrap.py:282-350)."""
    dev = require_cuda()
    w = None if weight is None else solvers.to_device_f64(weight, dev)
    return _finish(solvers.unwrap(dx=solvers.to_device_f64(dx, dev), dy=solvers.to_device_f64(dy, dev), weight=w, kmax=kmax))


def phase_unwrap_ref(psi, weight, kmax=100):
    """phase_unwrap.py:26-78 — same algorithm with the scaling recomputed every iteration."""
    return phase_unwrap(psi, weight, kmax)


def phase_unwrap_ref_prediff(dx, dy, weight=None, kmax=100):
    """phase_unwrap.py:211-279."""
    return phase_unwrap_prediff(dx, dy, weight, kmax)
