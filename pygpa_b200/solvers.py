"""Device-side drivers of K2 (PCG phase unwrap) and K3 (per-pixel least squares).
torch tensors in, torch tensors out; all float64 like the reference."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .engine import _count, _ptr, _stream, require_cuda, workspace

SRC_PLAIN, SRC_DIFF1, SRC_DIFF0, SRC_PREDIFF0, SRC_PREDIFF1 = range(5)
LSQ_WEIGHTED, LSQ_MATRIX = 0, 1
MAX_D = 8


def to_device_f64(a, device=None):
    device = device or require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device, non_blocking=True)


def unwrap(psi=None, dx=None, dy=None, weight=None, kmax=100, return_iters=False):
    """PCG unwrap on the device.  Give psi (N, M) or the gradients dx (N, M-1), dy (N-1, M)."""
    lib = _lib.load()
    if psi is not None:
        n, m = psi.shape
        dev = psi.device
    else:
        n, m = dx.shape[0], dy.shape[1]
        if dx.shape != (n, m - 1) or dy.shape != (n - 1, m):
            raise ValueError(f"dx {tuple(dx.shape)} / dy {tuple(dy.shape)} do not describe one (N, M) grid")
        dev = dx.device
    if weight is not None and tuple(weight.shape) != (n, m):
        raise ValueError("weight must have the shape of the unwrapped phase")
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.gpa_unwrap_workspace_bytes(n, m, ctypes.byref(nbytes)))
    ws = workspace(nbytes.value, dev)
    phi = torch.empty((n, m), dtype=torch.float64, device=dev)
    iters = ctypes.c_int(0)
    _lib.check(lib.gpa_unwrap_pcg(_ptr(psi), _ptr(dx), _ptr(dy), _ptr(weight), n, m, int(kmax), _ptr(phi),
                                  ctypes.byref(iters) if return_iters else None, _ptr(ws), ws.numel(), _stream()))
    _count(5 + 9 * max(1, int(kmax)))
    return (phi, iters.value) if return_iters else phi


def _uw_ws(n, m, dev):
    nbytes = ctypes.c_size_t(0)
    _lib.check(_lib.load().gpa_unwrap_workspace_bytes(n, m, ctypes.byref(nbytes)))
    return workspace(nbytes.value, dev)


def dctn(x, inverse=False):
    """scipy.fft.dctn / idctn (type 2, unnormalised) of a float64 CUDA matrix, on the device."""
    lib = _lib.load()
    n, m = x.shape
    ws = _uw_ws(n, m, x.device)
    out = torch.empty_like(x)
    _lib.check(lib.gpa_dctn(_ptr(x.contiguous()), n, m, int(bool(inverse)), _ptr(out), _ptr(ws), ws.numel(), _stream()))
    _count(4)
    return out


def poisson_scale(n, m, device=None):
    """precomp_Poissonscaling (phase_unwrap.py:106-115) for an (n, m) grid, float64 CUDA tensor."""
    device = device or require_cuda()
    out = torch.empty((n, m), dtype=torch.float64, device=device)
    _lib.check(_lib.load().gpa_poisson_scale(n, m, _ptr(out), _stream()))
    _count(1)
    return out


def divide(a, b):
    out = torch.empty_like(a)
    _lib.check(_lib.load().gpa_divide_f64(_ptr(a.contiguous()), _ptr(b.contiguous()), _ptr(out), a.numel(), _stream()))
    _count(1)
    return out


def solve_poisson(rho, scale=None):
    """idctn(dctn(rho) / scale) (solvePoisson_precomped, phase_unwrap.py:95-103); scale=None uses precomp_Poissonscaling."""
    n, m = rho.shape
    if scale is None:
        scale = poisson_scale(n, m, rho.device)
    return dctn(divide(dctn(rho), scale), inverse=True)


def apply_q(p, wwx, wwy):
    """applyQ (phase_unwrap.py:118-132) on float64 CUDA tensors: p (N, M), wwx (N, M-1), wwy (N-1, M)."""
    lib = _lib.load()
    n, m = p.shape
    if tuple(wwx.shape) != (n, m - 1) or tuple(wwy.shape) != (n - 1, m):
        raise ValueError("wwx must be (N, M-1) and wwy (N-1, M)")
    ws = workspace(512 + 8 * (-(-m // 64)) * (-(-n // 32)) + 256, p.device)
    q = torch.empty_like(p)
    _lib.check(lib.gpa_apply_q(_ptr(p.contiguous()), _ptr(wwx.contiguous()), _ptr(wwy.contiguous()), n, m, _ptr(q), _ptr(ws), ws.numel(), _stream()))
    _count(1)
    return q


def lstsq(src, kind, kvecs, weights=None, matrix=None, subtract_mean=False):
    """Per-pixel least squares on the device, (2, n, m) float64 (see include/gpa_b200.h, K3)."""
    lib = _lib.load()
    kv = np.ascontiguousarray(kvecs, dtype=np.float64).reshape(-1, 2)
    d = kv.shape[0]
    if d > MAX_D:
        raise ValueError(f"at most {MAX_D} k-vectors are supported")
    if src.shape[0] != d:
        raise ValueError("first axis of the phases must match the number of k-vectors")
    n, m = int(src.shape[1]), int(src.shape[2])
    on, om = {SRC_PLAIN: (n, m), SRC_DIFF1: (n, m - 1), SRC_DIFF0: (n - 1, m),
              SRC_PREDIFF0: (n, m - 1), SRC_PREDIFF1: (n - 1, m)}[kind]
    out = torch.empty((2, on, om), dtype=torch.float64, device=src.device)
    ws_t, ws_n = None, 0
    if subtract_mean:
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.gpa_lstsq_workspace_bytes(d, ctypes.byref(nbytes)))
        ws_t = workspace(nbytes.value, src.device)
        ws_n = ws_t.numel()
    mat = None
    if matrix is not None:
        mat = np.ascontiguousarray(matrix, dtype=np.float64)
        assert mat.shape == (2, d)
    wn, wm = (int(weights.shape[1]), int(weights.shape[2])) if weights is not None else (0, 0)
    _lib.check(lib.gpa_lstsq_u(_ptr(src), kind, _ptr(weights), wn, wm, _lib.as_pd(kv), d, n, m,
                               LSQ_MATRIX if matrix is not None else LSQ_WEIGHTED,
                               _lib.as_pd(mat) if mat is not None else None, int(subtract_mean),
                               _ptr(out), _ptr(ws_t), ws_n, _stream()))
    _count(3 if subtract_mean else 1)
    return out


def phasegradient_to_J(grads, weights, K, sub=None, order=None, do_wrap=False, nmperpixel=1.0, add_identity=False):
    """K5a: (d, N, M, 2) phase gradients + (d, >=N, >=M) weights -> J (N, M, 2, 2) on the device
    (property_extract.py:69-101; K, sub, order are the host-side quantities described in gpa_b200.h)."""
    lib = _lib.load()
    d, n, m = int(grads.shape[0]), int(grads.shape[1]), int(grads.shape[2])
    if tuple(grads.shape) != (d, n, m, 2) or weights.shape[0] != d:
        raise ValueError("grads must be (d, N, M, 2) and weights (d, N, M)")
    K = np.ascontiguousarray(K, dtype=np.float64).reshape(d, 2)
    sub_a = None if sub is None else np.ascontiguousarray(sub, dtype=np.float64).reshape(d, 2)
    ord_a = None if order is None else np.ascontiguousarray(order, dtype=np.int32).reshape(d)
    out = torch.empty((n, m, 2, 2), dtype=torch.float64, device=grads.device)
    _lib.check(lib.gpa_phasegradient_to_j(_ptr(grads), _ptr(weights), int(weights.shape[1]), int(weights.shape[2]),
                                          _lib.as_pd(K), _lib.as_pd(sub_a) if sub_a is not None else None,
                                          ord_a.ctypes.data_as(ctypes.POINTER(ctypes.c_int)) if ord_a is not None else None,
                                          int(do_wrap), d, n, m, float(nmperpixel), int(add_identity), _ptr(out), _stream()))
    _count(1)
    return out


def props_from_jac(jac, refangle=0.0, refscale=1.0, diff=False, add_identity=False):
    """K5b: (..., 2, 2) Jacobians -> (4, ...) properties on the device (property_extract.py:137-178)."""
    lib = _lib.load()
    if jac.shape[-2:] != (2, 2):
        raise ValueError("Jac must have shape (..., 2, 2)")
    lead = tuple(jac.shape[:-2])
    npix = int(np.prod(lead)) if lead else 1
    out = torch.empty((4,) + lead, dtype=torch.float64, device=jac.device)
    _lib.check(lib.gpa_props_from_jac(_ptr(jac), npix, float(refangle), float(refscale), int(bool(diff)),
                                      int(bool(add_identity)), _ptr(out), _stream()))
    _count(1)
    return out


def lockin_phase_amp(lockin, edge=0):
    """(angle, abs, max abs) of a complex lock-in tensor cropped by `edge` on every side
    (iterate_GPA, geometric_phase_analysis.py:134-139); the max stays on the device."""
    lib = _lib.load()
    n, m = int(lockin.shape[0]), int(lockin.shape[1])
    shape = (n - 2 * edge, m - 2 * edge)
    ph = torch.empty(shape, dtype=torch.float64, device=lockin.device)
    amp = torch.empty(shape, dtype=torch.float64, device=lockin.device)
    amax = torch.empty(1, dtype=torch.float64, device=lockin.device)
    _lib.check(lib.gpa_lockin_phase_amp(_ptr(lockin), int(lockin.dtype == torch.complex128), n, m, int(edge),
                                        _ptr(ph), _ptr(amp), _ptr(amax), _stream()))
    _count(1)
    return ph, amp, amax


def weight_sqrt_norm(amp, amax):
    """sqrt(amp / max amp) (geometric_phase_analysis.py:141)."""
    lib = _lib.load()
    out = torch.empty_like(amp)
    _lib.check(lib.gpa_weight_sqrt_norm(_ptr(amp), _ptr(amax), amp.numel(), _ptr(out), _stream()))
    _count(1)
    return out


def fit_plane_huber(img, f_scale=1.0, max_iter=500, tol=1e-11, return_iters=False):
    """Huber plane fit of a float64 CUDA image: host array [a_x, a_y, b] with img ~ a_x x + a_y y + b
    (mathtools.py:30-47)."""
    lib = _lib.load()
    n, m = int(img.shape[0]), int(img.shape[1])
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.gpa_fit_plane_workspace_bytes(ctypes.byref(nbytes)))
    ws = workspace(nbytes.value, img.device)
    theta = np.zeros(3)
    iters = ctypes.c_int(0)
    _lib.check(lib.gpa_fit_plane_huber(_ptr(img), n, m, float(f_scale), int(max_iter), float(tol), _lib.as_pd(theta),
                                       ctypes.byref(iters), _ptr(ws), ws.numel(), _stream()))
    _count(iters.value)
    return (theta, iters.value) if return_iters else theta


def gaussian_deconvolve(data, sigma, dr=20, balance=5000.0):
    """Wiener deconvolution of every trailing (N, M) plane of a float64 CUDA tensor with the Gaussian of
    std sigma (geometric_phase_analysis.py:892-904), on the device."""
    lib = _lib.load()
    n, m = int(data.shape[-2]), int(data.shape[-1])
    planes = int(data.numel() // (n * m))
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.gpa_deconvolve_workspace_bytes(n, m, int(dr), ctypes.byref(nbytes)))
    ws = workspace(nbytes.value, data.device)
    data = data.contiguous()
    out = torch.empty_like(data)
    _lib.check(lib.gpa_gaussian_deconvolve(_ptr(data), planes, n, m, float(sigma), int(dr), float(balance), _ptr(out),
                                           _ptr(ws), ws.numel(), _stream()))
    _count(4 + 7 * planes)
    return out


def norm_axis0(w):
    lib = _lib.load()
    out = torch.empty(w.shape[1:], dtype=torch.float64, device=w.device)
    _lib.check(lib.gpa_norm_axis0(_ptr(w), int(w.shape[0]), out.numel(), _ptr(out), _stream()))
    _count(1)
    return out


def phase_weight(lockin, border, eps=1e-6):
    """(phases, weights) float64 from a complex lock-in tensor (extract_displacement_field glue)."""
    lib = _lib.load()
    n, m = lockin.shape
    ph = torch.empty((n, m), dtype=torch.float64, device=lockin.device)
    w = torch.empty((n, m), dtype=torch.float64, device=lockin.device)
    _lib.check(lib.gpa_phase_weight(_ptr(lockin), int(lockin.dtype == torch.complex128), n, m, int(border),
                                    float(eps), _ptr(ph), _ptr(w), _stream()))
    _count(1)
    return ph, w


def displacement_from_phases(kvecs, phases, weights, weighted_unwrap=True, pre_diff=False):
    """reconstruct_u_inv_from_phases (geometric_phase_analysis.py:196-245) on device tensors:
    wrapped differences -> two per-pixel least-squares solves -> two PCG integrations (kmax=10)."""
    k0, k1 = (SRC_PREDIFF0, SRC_PREDIFF1) if pre_diff else (SRC_DIFF1, SRC_DIFF0)
    dudx = lstsq(phases, k0, kvecs, weights)
    dudy = lstsq(phases, k1, kvecs, weights)
    if weighted_unwrap:
        wn = norm_axis0(weights)
        us = [unwrap(dx=dudx[i], dy=dudy[i], weight=wn, kmax=10) for i in range(2)]
    else:
        us = [unwrap(dx=dudx[i], dy=dudy[i]) for i in range(2)]
    return torch.stack(us)


def _lawler_ws(n, m, edge, device):
    lib = _lib.load()
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.gpa_lawler_workspace_bytes(n, m, edge, ctypes.byref(nbytes)))
    return workspace(nbytes.value, device)


def invert_u(u, iters=35, edge=0, scale=1.0, overlap=True):
    """Fixed-point inverse of the displacement field (2, N, M), on the device: invert_u_overlap -> (2, N+2e, M+2e);
    overlap=False: the reference's invert_u -> (2, N, M), `- edge` in the iterations only."""
    lib = _lib.load()
    n, m = int(u.shape[1]), int(u.shape[2])
    ws = _lawler_ws(n, m, edge, u.device)
    e_out = edge if overlap else 0
    out = torch.empty((2, n + 2 * e_out, m + 2 * e_out), dtype=torch.float64, device=u.device)
    fn = lib.gpa_invert_u if overlap else lib.gpa_invert_u_plain
    _lib.check(fn(_ptr(u), n, m, float(scale), int(iters), int(edge), _ptr(out), _ptr(ws), ws.numel(), _stream()))
    _count(9)
    return out


def undistort(img, u, iters=35):
    """undistort_image on the device: invert -u, then resample img (N, M) with zeros outside."""
    lib = _lib.load()
    n, m = int(img.shape[0]), int(img.shape[1])
    ws = _lawler_ws(n, m, 0, img.device)
    out = torch.empty((n, m), dtype=torch.float64, device=img.device)
    _lib.check(lib.gpa_undistort_image(_ptr(img), _ptr(u), n, m, int(iters), _ptr(out), _ptr(ws), ws.numel(), _stream()))
    _count(14)
    return out
