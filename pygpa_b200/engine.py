"""Device-side driver of the lock-in / sweep kernels (K1).

PyTorch is used for device buffers, streams and pinned host memory only; every kernel
that runs is one of libgpa_b200.so's.  Functions here take and return torch CUDA tensors;
the NumPy-in / NumPy-out mirrors of the reference API live in cuGPA.py and
geometric_phase_analysis.py.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from ._taps import DEFAULT_TRUNC, axis_taps, multirate_taps, split_taps

GRAD_CENTRAL, GRAD_FORWARD, GRAD_NONE = 0, 1, 2
CAND_GRID, CAND_LIST = 0, 1

_workspaces = {}
launch_count = 0     # kernels of ours enqueued so far (bench.py reports the delta)


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.GpaError("pygpa_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    _lib.load()
    return torch.device("cuda", torch.cuda.current_device())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _ws_key(device):
    return (device, torch.cuda.current_stream(device).cuda_stream)


def workspace(nbytes, device):
    """Grow-only scratch buffer per (device, current stream) — the C ABI never allocates.  Every entry point enqueues on
    the current torch stream, so work issued on two streams never shares scratch, and a buffer is only ever replaced by the
    stream that uses it (the caching allocator recycles it in that stream's order)."""
    key = _ws_key(device)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        _workspaces[key] = None
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def set_pruning(on):
    """Toggle the exact branch-and-bound candidate pruning of the multirate arg-max (default on)."""
    _lib.check(_lib.load().gpa_set_pruning(int(bool(on))))


def set_tma(on):
    """Toggle the TMA box loads of the interpolation kernel's coarse tiles (default on; results are identical)."""
    _lib.check(_lib.load().gpa_set_tma(int(bool(on))))


def release_workspaces():
    _workspaces.clear()


def image_to_device(image, device=None):
    """Host image (any real dtype) -> contiguous float32 CUDA tensor, cast on the device."""
    device = device or require_cuda()
    if isinstance(image, torch.Tensor):
        t = image.to(device)
        return t if t.dtype == torch.float32 and t.is_contiguous() else t.float().contiguous()
    arr = np.ascontiguousarray(image)
    if arr.ndim != 2:
        raise ValueError("image must be 2-D")
    if np.iscomplexobj(arr):
        raise TypeError("image must be real")
    if arr.dtype == np.float32:
        return torch.from_numpy(arr).to(device, non_blocking=True)
    if arr.dtype != np.float64:
        arr = arr.astype(np.float64)
    staged = torch.from_numpy(arr).to(device, non_blocking=True)
    out = torch.empty(arr.shape, dtype=torch.float32, device=device)
    lib = _lib.load()
    _lib.check(lib.gpa_cast_f64_to_f32(_ptr(staged), _ptr(out), arr.size, _stream()))
    _count(1)
    return out


def _count(n):
    global launch_count
    launch_count += n


def _plan_planes(n, m, n_rows, n_planes, rx, ry, device, planes_in_flight):
    lib = _lib.load()
    nbytes = ctypes.c_size_t(0)

    def need(p):
        _lib.check(lib.gpa_lockin_workspace_bytes(n, m, n_rows, n_planes, rx, ry, p, ctypes.byref(nbytes)))
        return nbytes.value
    if planes_in_flight is None:
        free, _total = torch.cuda.mem_get_info(device)
        cached = _workspaces.get(_ws_key(device))
        budget = int(0.6 * free) + (cached.numel() if cached is not None else 0)
        base, full = need(1), need(n_planes)
        if full <= budget or n_planes == 1:
            planes_in_flight = n_planes
        else:
            per = (full - base) // (n_planes - 1)
            planes_in_flight = int(max(1, min(n_planes, (budget - base) // per + 1)))
    return planes_in_flight, need(planes_in_flight)


def lockin_fixed(img_dev, kvec, sigma, trunc=DEFAULT_TRUNC, out_f64=False):
    """Fixed-reference lock-in of a float32 CUDA image; returns a complex CUDA tensor (N, M)."""
    device = img_dev.device
    lib = _lib.load()
    n, m = img_dev.shape
    tx, rx = axis_taps(n, sigma, trunc)
    ty, ry = axis_taps(m, sigma, trunc)
    _p, nbytes = _plan_planes(n, m, 1, 1, rx, ry, device, 1)
    ws = workspace(nbytes, device)
    out = torch.empty((n, m), dtype=torch.complex128 if out_f64 else torch.complex64, device=device)
    _lib.check(lib.gpa_lockin_fixed(_ptr(img_dev), n, m, float(kvec[0]), float(kvec[1]),
                                    _lib.as_pf(tx), rx, _lib.as_pf(ty), ry, int(out_f64), _ptr(out),
                                    _ptr(ws), ws.numel(), _stream()))
    _count(4)
    return out


class SweepPlan:
    """Geometry + scratch of one sweep call (candidate axes, taps, workspace)."""

    def __init__(self, shape, wx_rows, wy_planes, sigma, cand_mode=CAND_GRID, trunc=DEFAULT_TRUNC,
                 planes_in_flight=None, device=None, method="auto", private_ws=False, split_y=True):
        self.device = device or require_cuda()
        self._private = bool(private_ws)       # own scratch: what argmax() leaves survives other plans' calls
        self._ws = None
        self.n, self.m = int(shape[0]), int(shape[1])
        self.wx = np.ascontiguousarray(wx_rows, dtype=np.float64)
        self.wy = np.ascontiguousarray(wy_planes, dtype=np.float64)
        if self.wx.ndim != 1 or self.wy.ndim != 1 or self.wx.size == 0 or self.wy.size == 0:
            raise ValueError("candidate axes must be non-empty 1-D arrays")
        self.cand_mode = cand_mode
        self.tx, self.rx = axis_taps(self.n, sigma, trunc)
        self.ty, self.ry = axis_taps(self.m, sigma, trunc)
        self.in_flight, self.ws_bytes = _plan_planes(self.n, self.m, self.wx.size, self.wy.size,
                                                     self.rx, self.ry, self.device, planes_in_flight)
        self.n_cand = self.wx.size * self.wy.size if cand_mode == CAND_GRID else self.wy.size
        # arg-max method: the multirate form when the frame and sigma allow it (and the candidate
        # grid is large enough to amortise its extra passes), else the direct form
        if method not in ("auto", "direct", "multirate", "multirate-single"):
            raise ValueError("method must be 'auto', 'direct', 'multirate' or 'multirate-single'")
        self.mr = None
        if method != "direct" and trunc == DEFAULT_TRUNC and self.wx.size <= 65535:
            self.mr = multirate_taps(self.n, self.m, float(sigma), trunc)
            if self.mr is not None and method == "auto" and cand_mode == CAND_LIST:
                self.mr = None
        if method in ("multirate", "multirate-single") and self.mr is None:
            raise ValueError("the multirate sweep does not apply to this frame size / sigma")
        # split pass 2 (anchor stage shared by the candidates of a plane) when the grid is narrow enough;
        # 'multirate-single' keeps every candidate's own full-rate pass 2
        self.split = None
        self.split_y = None                     # the same along axis 1: one anchor plane + a coarse-rate stage per plane
        if self.mr is not None and cand_mode == CAND_GRID and method != "multirate-single":
            self.split = split_taps(self.n, self.mr, self.wx)
            if split_y:
                self.split_y = split_taps(self.m, self.mr, self.wy)
        if self.mr is not None:
            self.mr_in_flight, self.mr_ws_bytes = self._plan_mr(planes_in_flight)
            self.ws_bytes = max(self.ws_bytes, self.mr_ws_bytes)

    def _plan_mr(self, planes_in_flight):
        lib = _lib.load()
        mr, nbytes = self.mr, ctypes.c_size_t(0)

        def need(p):
            _lib.check(lib.gpa_sweep_mr_workspace_bytes(self.n, self.m, self.wx.size, self.wy.size, self.cand_mode,
                                                        mr["S"], mr["Ra_x"], mr["Ra_y"], mr["Rb"], *self._split_geom(),
                                                        p, ctypes.byref(nbytes)))
            return nbytes.value
        p = planes_in_flight
        if p is None:
            free, _total = torch.cuda.mem_get_info(self.device)
            cached = _workspaces.get(_ws_key(self.device))
            budget = int(0.6 * free) + (cached.numel() if cached is not None else 0)
            base, full = need(1), need(self.wy.size)
            if full <= budget or self.wy.size == 1:
                p = self.wy.size
            else:
                per = (full - base) // (self.wy.size - 1)
                p = int(max(1, min(self.wy.size, (budget - base) // per + 1)))
        return p, need(p)

    def _workspace(self):
        if not self._private:
            return workspace(self.ws_bytes, self.device)
        if self._ws is None or self._ws.numel() < self.ws_bytes:
            self._ws = torch.empty(int(self.ws_bytes), dtype=torch.uint8, device=self.device)
        return self._ws

    def _geom(self):
        return (self.n, self.m, _lib.as_pd(self.wx), self.wx.size, _lib.as_pd(self.wy), self.wy.size, self.cand_mode)

    def _taps(self):
        return (_lib.as_pf(self.tx), self.rx, _lib.as_pf(self.ty), self.ry)

    def _split_geom(self):
        gx = (self.split["R1"], self.split["H"]) if self.split else (0, 0)
        gy = (self.split_y["R1"], self.split_y["H"]) if self.split_y else (0, 0)
        return gx + gy

    def _split_args(self):
        sp, spy = self.split, self.split_y
        sa = float(self.mr["sigma_a"])
        ax = (_lib.as_pf(sp["taps_1"]), sp["R1"], _lib.as_pf(sp["taps_2"]), sp["H"], sa, sp["sigma_1"]) if sp else (None, 0, None, 0, sa, 0.0)
        ay = (_lib.as_pf(spy["taps_1"]), spy["R1"], _lib.as_pf(spy["taps_2"]), spy["H"], spy["sigma_1"]) if spy else (None, 0, None, 0, 0.0)
        return ax + ay

    def argmax(self, img_dev, key, plane_begin=0, plane_end=None, plane_step=1):
        """key (N, M) int64 CUDA tensor, updated in place with the candidates of planes
        plane_begin, plane_begin + plane_step, ... < plane_end (a step > 1 needs the multirate form)."""
        lib = _lib.load()
        plane_end = self.wy.size if plane_end is None else plane_end
        ws = self._workspace()
        if self.mr is None and plane_step != 1:
            raise ValueError("interleaved plane shares are only available in the multirate form")
        if self.mr is not None:
            mr = self.mr
            _lib.check(lib.gpa_sweep_argmax_mr(_ptr(img_dev), *self._geom(), plane_begin, plane_end, plane_step, mr["S"],
                                               _lib.as_pf(mr["taps_ax"]), mr["Ra_x"], _lib.as_pf(mr["taps_ay"]), mr["Ra_y"],
                                               _lib.as_pf(mr["taps_bx"]), _lib.as_pf(mr["taps_by"]), mr["Rb"],
                                               *self._split_args(), _ptr(key), _ptr(ws), ws.numel(), _stream()))
            n_local = -(-(plane_end - plane_begin) // plane_step)
            chunks = -(-n_local // self.mr_in_flight) if plane_end > plane_begin else 0
            _count(2 + (6 if self.split else 4) * chunks + (1 if self.split and chunks else 0) + (2 if self.split_y and chunks else 0))
            return
        _lib.check(lib.gpa_sweep_argmax(_ptr(img_dev), *self._geom(), plane_begin, plane_end, *self._taps(),
                                        _ptr(key), _ptr(ws), ws.numel(), _stream()))
        chunks = -(-(plane_end - plane_begin) // self.in_flight) if plane_end > plane_begin else 0
        _count(2 + 2 * chunks)

    def finalize(self, img_dev, key, kref, grad_mode=GRAD_CENTRAL, out_f64=False, want_w=False, want_kidx=True,
                 plane_begin=0, plane_end=None, planes_valid=False, out=None, plane_step=1):
        """Winner's lock-in / gradient / w / k-index for pixels won by planes plane_begin,
        plane_begin + plane_step, ... < plane_end.
        planes_valid: the workspace still holds what argmax() left for exactly this range (then the
        multirate path interpolates from its coarse grids and the direct path skips pass 1)."""
        lib = _lib.load()
        plane_end = self.wy.size if plane_end is None else plane_end
        n, m, dev = self.n, self.m, self.device
        real = torch.float64 if out_f64 else torch.float32
        cplx = torch.complex128 if out_f64 else torch.complex64
        if out is None:
            out = {}
            full = plane_begin == 0 and plane_end == self.wy.size and plane_step == 1
            make = torch.empty if full else torch.zeros
            out["lockin"] = make((n, m), dtype=cplx, device=dev)
            out["grad"] = make((n, m, 2), dtype=real, device=dev) if grad_mode != GRAD_NONE else None
            out["w"] = make((2, n, m), dtype=real, device=dev) if want_w else None
            out["kidx"] = make((n, m), dtype=torch.int32, device=dev) if want_kidx else None
        ws = self._workspace()
        if plane_step != 1 and not (self.mr is not None and planes_valid):
            raise ValueError("an interleaved plane share can only be finalized from its resident coarse grids")
        if self.mr is not None and planes_valid and -(-(plane_end - plane_begin) // plane_step) <= self.mr_in_flight:
            mr = self.mr
            _lib.check(lib.gpa_sweep_finalize_mr(*self._geom(), plane_begin, plane_end, plane_step, mr["S"], mr["Ra_x"], mr["Ra_y"],
                                                 _lib.as_pf(mr["taps_bx"]), _lib.as_pf(mr["taps_by"]), mr["Rb"],
                                                 *self._split_geom(),
                                                 _ptr(key), float(kref[0]), float(kref[1]), grad_mode, int(out_f64),
                                                 _ptr(out["lockin"]), _ptr(out.get("grad")), _ptr(out.get("w")),
                                                 _ptr(out.get("kidx")), _ptr(ws), ws.numel(), _stream()))
            _count(1)
            return out
        if self.mr is not None:
            planes_valid = False       # the workspace holds coarse grids, not full-resolution planes
        _lib.check(lib.gpa_sweep_finalize(_ptr(img_dev), *self._geom(), plane_begin, plane_end, int(planes_valid),
                                          *self._taps(), _ptr(key), float(kref[0]), float(kref[1]), grad_mode,
                                          int(out_f64), _ptr(out["lockin"]), _ptr(out.get("grad")),
                                          _ptr(out.get("w")), _ptr(out.get("kidx")), _ptr(ws), ws.numel(), _stream()))
        chunks = -(-(plane_end - plane_begin) // self.in_flight) if plane_end > plane_begin else 0
        reuse = planes_valid and chunks == 1
        _count(chunks if reuse else 2 + 2 * chunks)
        return out

    def run(self, img_dev, kref, grad_mode=GRAD_CENTRAL, out_f64=False, want_w=False, want_kidx=True):
        """Whole sweep on this GPU: zero key, arg-max over all planes, finalize."""
        if self.mr is not None:
            key = torch.zeros((self.n, self.m), dtype=torch.int64, device=self.device)
            self.argmax(img_dev, key)
            out = self.finalize(img_dev, key, kref, grad_mode, out_f64, want_w, want_kidx,
                                planes_valid=self.mr_in_flight == self.wy.size)
            out["key"] = key
            return out
        lib = _lib.load()
        n, m, dev = self.n, self.m, self.device
        real = torch.float64 if out_f64 else torch.float32
        out = {"key": torch.empty((n, m), dtype=torch.int64, device=dev),
               "lockin": torch.empty((n, m), dtype=torch.complex128 if out_f64 else torch.complex64, device=dev),
               "grad": torch.empty((n, m, 2), dtype=real, device=dev) if grad_mode != GRAD_NONE else None,
               "w": torch.empty((2, n, m), dtype=real, device=dev) if want_w else None,
               "kidx": torch.empty((n, m), dtype=torch.int32, device=dev) if want_kidx else None}
        ws = self._workspace()
        _lib.check(lib.gpa_wfr_sweep(_ptr(img_dev), *self._geom(), *self._taps(), float(kref[0]), float(kref[1]),
                                     grad_mode, int(out_f64), _ptr(out["key"]), _ptr(out["lockin"]),
                                     _ptr(out["grad"]), _ptr(out["w"]), _ptr(out["kidx"]), _ptr(ws), ws.numel(),
                                     _stream()))
        chunks = -(-self.wy.size // self.in_flight)
        _count(2 + 2 * chunks + (1 if chunks == 1 else 2 + 2 * chunks))
        return out


def wfr4_sweep(img_dev, sigma, klist, kref, dk, trunc=DEFAULT_TRUNC, out_f64=True):
    """Ordered k-list sweep with the neighbourhood acceptance rule of wfr4
    (geometric_phase_analysis.py:839-862) on a float32 CUDA image; dict of CUDA tensors."""
    lib = _lib.load()
    device = img_dev.device
    n, m = img_dev.shape
    klist = np.ascontiguousarray(klist, dtype=np.float64).reshape(-1, 2)
    kx, ky = klist[:, 0].copy(), klist[:, 1].copy()
    K = kx.size
    if K == 0:
        raise ValueError("klist is empty")
    # the reference's test (:854), evaluated once per pair of list entries: same float64 expression
    allowed = np.linalg.norm(klist[:, None, :] - klist[None, :, :], axis=-1) < 2 * np.sqrt(2) * dk
    allowed_dev = torch.from_numpy(np.ascontiguousarray(allowed, dtype=np.uint8)).to(device)
    tx, rx = axis_taps(n, sigma, trunc)
    ty, ry = axis_taps(m, sigma, trunc)
    in_flight, nbytes = _plan_planes(n, m, K, K, rx, ry, device, None)
    ws = workspace(nbytes, device)
    out = {"key": torch.empty((n, m), dtype=torch.int64, device=device),
           "lockin": torch.empty((n, m), dtype=torch.complex128 if out_f64 else torch.complex64, device=device),
           "w": torch.empty((2, n, m), dtype=torch.float64 if out_f64 else torch.float32, device=device),
           "kidx": torch.empty((n, m), dtype=torch.int32, device=device)}
    _lib.check(lib.gpa_wfr4_sweep(_ptr(img_dev), n, m, _lib.as_pd(kx), _lib.as_pd(ky), K, _ptr(allowed_dev),
                                  _lib.as_pf(tx), rx, _lib.as_pf(ty), ry, float(kref[0]), float(kref[1]), int(out_f64),
                                  _ptr(out["key"]), _ptr(out["lockin"]), _ptr(out["w"]), _ptr(out["kidx"]),
                                  _ptr(ws), ws.numel(), _stream()))
    chunks = -(-K // in_flight)
    _count(3 + 2 * chunks + (1 if chunks == 1 else 2 + 2 * chunks))
    return out


def grid_axes(kx, ky, kw, kstep):
    """The reference's candidate axes, verbatim NumPy expression because the lengths are
    rounding dependent (geometric_phase_analysis.py:803-804)."""
    return np.arange(kx - kw, kx + kw, kstep), np.arange(ky - kw, ky + kw, kstep)
