"""Drop-in for ``pyGPA.phase_unwrap`` (reference: pyGPA/phase_unwrap.py) on B200: weighted
least-squares phase unwrapping (Ghiglia & Romero PCG with a DCT Poisson preconditioner),
float64 on the device."""
from __future__ import annotations

import numpy as np
import torch

from . import solvers
from .cuGPA import _to_host
from .engine import require_cuda

__all__ = ["phase_unwrap", "phase_unwrap_prediff", "phase_unwrap_ref", "phase_unwrap_ref_prediff",
           "solvePoisson", "solvePoisson_precomped", "precomp_Poissonscaling", "applyQ"]


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


def precomp_Poissonscaling(rho):
    """phase_unwrap.py:106-115: 2 (cos(pi I / M) + cos(pi J / N) - 2) with [0, 0] = 1 (N / M swapped as in the
    reference), evaluated on the device."""
    n, m = np.shape(rho)
    return _finish(solvers.poisson_scale(n, m, require_cuda()))


def solvePoisson_precomped(rho, scale):
    """phase_unwrap.py:95-103: idctn(dctn(rho) / scale) with scipy's unnormalised type-2 transforms, on the device."""
    dev = require_cuda()
    return _finish(solvers.solve_poisson(solvers.to_device_f64(rho, dev), solvers.to_device_f64(scale, dev)))


def solvePoisson(rho):
    """phase_unwrap.py:81-92: as solvePoisson_precomped, with the [0, 0] coefficient set to 0 instead of kept
    (the solution then has the DC term of idctn removed: the constant dctn(rho)[0, 0] / (4 N M) = mean(rho))."""
    dev = require_cuda()
    rho_d = solvers.to_device_f64(rho, dev)
    return _finish(solvers.solve_poisson(rho_d)) - float(np.mean(rho))


def applyQ(p, WWx, WWy):
    """phase_unwrap.py:118-132: (A^T)(W^T)(W)(A) p for edge weights WWx (N, M-1), WWy (N-1, M)."""
    dev = require_cuda()
    return _finish(solvers.apply_q(solvers.to_device_f64(p, dev), solvers.to_device_f64(WWx, dev), solvers.to_device_f64(WWy, dev)))
