"""Drop-in for ``pyGPA.unit_cell_averaging`` on B200: average an image
over the unit cell of its lattice (drizzle scatter-add) and expand the cell back to a full image.

    calc_ucell_parameters   unit_cell_averaging.py:45-53   (host, O(1))
    unit_cell_average       unit_cell_averaging.py:132-205 (K7 scatter kernel)
    expand_unitcell         unit_cell_averaging.py:234-249 (K4 cubic-spline gather)

NumPy in, NumPy out, float64.  The scatter adds with fp64 atomics, so results agree with the
reference's serial loop to rounding (1e-13 relative), not bit for bit.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, engine, solvers
from .cuGPA import _to_host
from .engine import _count, _ptr, _stream, workspace

__all__ = ["forward_transform", "backward_transform", "cart_in_uc", "calc_ucell_parameters", "unit_cell_average",
           "expand_unitcell"]


def forward_transform(vecs, ks):
    """unit_cell_averaging.py:7-10 (host)."""
    return vecs @ ks.T


def backward_transform(vecs, ks):
    """unit_cell_averaging.py:13-16 (host)."""
    return vecs @ np.linalg.inv(ks).T


def cart_in_uc(vecs, ks, rmin=0):
    """unit_cell_averaging.py:29-34 (host)."""
    return backward_transform(forward_transform(vecs, ks) % 1., ks) - rmin


def calc_ucell_parameters(ks, z):
    """unit_cell_averaging.py:45-53: (rmin, rsize) of the zoomed unit-cell array."""
    corners = np.array([[0., 0.], [0., 1.], [1., 0.], [1., 1.]])
    cornervals = backward_transform(corners, ks)
    rmin = cornervals.min(axis=0)
    rsize = tuple((z * np.ceil(cornervals.max(axis=0) - np.floor(rmin))).astype(int))
    return rmin, rsize


def _geometry(ks, z):
    ks = np.ascontiguousarray(ks, dtype=np.float64)
    if ks.shape != (2, 2):
        raise ValueError("ks must be the two k-vectors spanning the unit cell, shape (2, 2)")
    rmin, rsize = calc_ucell_parameters(ks, z)
    return ks, np.ascontiguousarray(np.linalg.inv(ks)), np.ascontiguousarray(rmin, dtype=np.float64), rsize


def _ws(rs0, rs1, device):
    nbytes = ctypes.c_size_t(0)
    _lib.check(_lib.load().gpa_uc_workspace_bytes(int(rs0), int(rs1), ctypes.byref(nbytes)))
    return workspace(nbytes.value, device)


def _host(t):
    h = _to_host(t)
    torch.cuda.current_stream().synchronize()
    return h.numpy()


def unit_cell_average_device(image, ks, u=None, z=1):
    """float64 CUDA tensors in, the (rs0, rs1) cell as a CUDA tensor out."""
    lib = _lib.load()
    ks, kinv, rmin, rsize = _geometry(ks, z)
    n, m = int(image.shape[0]), int(image.shape[1])
    if u is not None and tuple(u.shape) != (2, n, m):
        raise ValueError("u must have shape (2,) + image.shape")
    out = torch.empty(rsize, dtype=torch.float64, device=image.device)
    ws = _ws(rsize[0], rsize[1], image.device)
    _lib.check(lib.gpa_uc_average(_ptr(image), _ptr(u), n, m, _lib.as_pd(ks), _lib.as_pd(kinv), _lib.as_pd(rmin),
                                  float(z), int(rsize[0]), int(rsize[1]), _ptr(out), _ptr(ws), ws.numel(), _stream()))
    _count(2)
    return out


def unit_cell_average(image, ks, u=None, z=1, only_generate_func=False):
    """Average `image` over the unit cell spanned by the k-vectors `ks` (2, 2), optionally through the
    deformation field u (2, N, M), on a cell array zoomed by z; NaN pixels are ignored and cells nothing
    maps to are NaN (unit_cell_averaging.py:132-205)."""
    dev = engine.require_cuda()
    if only_generate_func:
        # the reference hands back its inner numba function, which takes u as (N, M, 2)
        return lambda img, uu: unit_cell_average(img, ks, np.moveaxis(np.asarray(uu), -1, 0), z=z)
    ud = None if u is None else solvers.to_device_f64(u, dev)
    return _host(unit_cell_average_device(solvers.to_device_f64(image, dev), ks, ud, z))


def expand_unitcell(unit_cell_image, ks, shape, z=1, z2=1, u=0):
    """Recreate a full image of `shape` from a unit-cell image made with zoom z, optionally on a grid z2
    times finer and through the distortion u (unit_cell_averaging.py:234-249)."""
    dev = engine.require_cuda()
    lib = _lib.load()
    ks, kinv, rmin, _rsize = _geometry(ks, z)
    cell = solvers.to_device_f64(unit_cell_image, dev)
    h, w = int(shape[0]), int(shape[1])
    ud, uc = None, 0.0
    if np.ndim(u) == 0:
        uc = float(u)
    else:
        ud = solvers.to_device_f64(np.broadcast_to(np.asarray(u, dtype=np.float64), (2, h, w)), dev)
    out = torch.empty((h, w), dtype=torch.float64, device=dev)
    ws = _ws(cell.shape[0], cell.shape[1], dev)
    _lib.check(lib.gpa_uc_expand(_ptr(cell), int(cell.shape[0]), int(cell.shape[1]), h, w, _ptr(ud), uc, float(z2),
                                 _lib.as_pd(ks), _lib.as_pd(kinv), _lib.as_pd(rmin), float(z), _ptr(out), _ptr(ws),
                                 ws.numel(), _stream()))
    _count(8)
    return _host(out)
