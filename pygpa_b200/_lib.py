"""ctypes binding of libgpa_b200.so (C ABI declared in include/gpa_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, an exception
is raised.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C pygpa_b200/csrc``.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpa_b200.so")

c_int, c_double, c_size_t, c_void_p = ctypes.c_int, ctypes.c_double, ctypes.c_size_t, ctypes.c_void_p
_pd = ctypes.POINTER(ctypes.c_double)
_pf = ctypes.POINTER(ctypes.c_float)

# name -> (restype, argtypes).  Device pointers are passed as c_void_p integers.
SIGNATURES = {
    "gpa_last_error": (ctypes.c_char_p, []),
    "gpa_version": (c_int, []),
    "gpa_device_sm_count": (c_int, []),
    "gpa_profile_enable": (c_int, [c_int]),
    "gpa_profile_read": (c_int, [ctypes.c_char_p, ctypes.POINTER(c_double), ctypes.POINTER(c_int), c_int]),
    "gpa_fp32_peak_tflops": (c_int, [c_void_p, c_size_t, ctypes.POINTER(c_double), c_void_p]),
    "gpa_cast_f64_to_f32": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "gpa_key_to_kidx": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "gpa_phase_weight": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p]),
    "gpa_gaussian_taps": (c_int, [c_int, c_double, c_int, _pf]),
    "gpa_default_radius": (c_int, [c_int, c_double, c_double]),
    "gpa_multirate_plan": (c_int, [c_int, c_int, c_double, _pd, _pd, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "gpa_split_plan": (c_int, [c_int, c_int, c_double, _pd, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), _pd, _pf, _pf]),
    "gpa_lockin_workspace_bytes": (c_int, [c_int] * 7 + [ctypes.POINTER(c_size_t)]),
    "gpa_lockin_fixed": (c_int, [c_void_p, c_int, c_int, c_double, c_double, _pf, c_int, _pf, c_int,
                                 c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gpa_sweep_argmax": (c_int, [c_void_p, c_int, c_int, _pd, c_int, _pd, c_int, c_int, c_int, c_int,
                                 _pf, c_int, _pf, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gpa_set_pruning": (c_int, [c_int]),
    "gpa_set_tma": (c_int, [c_int]),
    "gpa_set_dct_pipeline": (c_int, [c_int]),
    "gpa_sweep_mr_workspace_bytes": (c_int, [c_int] * 14 + [ctypes.POINTER(c_size_t)]),
    "gpa_sweep_argmax_mr": (c_int, [c_void_p, c_int, c_int, _pd, c_int, _pd, c_int, c_int, c_int, c_int, c_int, c_int,
                                    _pf, c_int, _pf, c_int, _pf, _pf, c_int, _pf, c_int, _pf, c_int, c_double, c_double,
                                    _pf, c_int, _pf, c_int, c_double, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gpa_sweep_finalize_mr": (c_int, [c_int, c_int, _pd, c_int, _pd, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                      _pf, _pf, c_int, c_int, c_int, c_int, c_int, c_void_p, c_double, c_double, c_int, c_int,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gpa_sweep_finalize": (c_int, [c_void_p, c_int, c_int, _pd, c_int, _pd, c_int, c_int, c_int, c_int, c_int,
                                   _pf, c_int, _pf, c_int, c_void_p, c_double, c_double, c_int, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gpa_wfr_sweep": (c_int, [c_void_p, c_int, c_int, _pd, c_int, _pd, c_int, c_int,
                              _pf, c_int, _pf, c_int, c_double, c_double, c_int, c_int,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gpa_wfr4_sweep": (c_int, [c_void_p, c_int, c_int, _pd, _pd, c_int, c_void_p, _pf, c_int, _pf, c_int,
                               c_double, c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_size_t, c_void_p]),
    "gpa_lstsq_workspace_bytes": (c_int, [c_int, ctypes.POINTER(c_size_t)]),
    "gpa_lstsq_u": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, _pd, c_int, c_int, c_int,
                            c_int, _pd, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gpa_phasegradient_to_j": (c_int, [c_void_p, c_void_p, c_int, c_int, _pd, _pd, ctypes.POINTER(c_int), c_int, c_int,
                                       c_int, c_int, c_double, c_int, c_void_p, c_void_p]),
    "gpa_props_from_jac": (c_int, [c_void_p, c_size_t, c_double, c_double, c_int, c_int, c_void_p, c_void_p]),
    "gpa_lockin_phase_amp": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gpa_weight_sqrt_norm": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "gpa_fit_plane_workspace_bytes": (c_int, [ctypes.POINTER(c_size_t)]),
    "gpa_fit_plane_huber": (c_int, [c_void_p, c_int, c_int, c_double, c_int, c_double, _pd, ctypes.POINTER(c_int),
                                    c_void_p, c_size_t, c_void_p]),
    "gpa_uc_workspace_bytes": (c_int, [c_int, c_int, ctypes.POINTER(c_size_t)]),
    "gpa_uc_average": (c_int, [c_void_p, c_void_p, c_int, c_int, _pd, _pd, _pd, c_double, c_int, c_int, c_void_p,
                               c_void_p, c_size_t, c_void_p]),
    "gpa_uc_expand": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_double, c_double, _pd, _pd, _pd,
                              c_double, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gpa_deconvolve_workspace_bytes": (c_int, [c_int, c_int, c_int, ctypes.POINTER(c_size_t)]),
    "gpa_gaussian_deconvolve": (c_int, [c_void_p, c_int, c_int, c_int, c_double, c_int, c_double, c_void_p,
                                        c_void_p, c_size_t, c_void_p]),
    "gpa_norm_axis0": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_void_p]),
    "gpa_dctn": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gpa_poisson_scale": (c_int, [c_int, c_int, c_void_p, c_void_p]),
    "gpa_divide_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gpa_apply_q": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gpa_unwrap_workspace_bytes": (c_int, [c_int, c_int, ctypes.POINTER(c_size_t)]),
    "gpa_unwrap_pcg": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                               ctypes.POINTER(c_int), c_void_p, c_size_t, c_void_p]),
    "gpa_peer_alloc": (c_int, [c_size_t, ctypes.POINTER(c_void_p), ctypes.POINTER(ctypes.c_ubyte)]),
    "gpa_peer_open": (c_int, [ctypes.POINTER(ctypes.c_ubyte), ctypes.POINTER(c_void_p)]),
    "gpa_peer_close": (c_int, [c_void_p]),
    "gpa_peer_free": (c_int, [c_void_p]),
    "gpa_peer_signal": (c_int, [ctypes.POINTER(c_void_p), c_int, ctypes.c_ulonglong, c_void_p]),
    "gpa_peer_wait": (c_int, [c_void_p, c_int, ctypes.c_ulonglong, c_double, c_void_p, c_void_p]),
    "gpa_peer_copy": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "gpa_host_register": (c_int, [c_void_p, c_size_t]),
    "gpa_host_unregister": (c_int, [c_void_p]),
    "gpa_key_merge": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_size_t, c_void_p]),
    "gpa_sweep_arm_gossip": (c_int, [ctypes.POINTER(c_void_p), c_int, ctypes.c_uint]),
    "gpa_sweep_arm_two_phase": (c_int, [ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), c_void_p, c_void_p,
                                        c_int, ctypes.c_ulonglong, c_double, c_void_p]),
    "gpa_key_to_w": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "gpa_sweep_finalize_mr_sharded": (c_int, [c_int, c_int, _pd, c_int, _pd, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                              _pf, _pf, c_int, c_int, c_int, c_int, c_int, c_void_p, c_double, c_double, c_int, c_int,
                                              ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), c_int, c_int, c_int,
                                              c_void_p, c_size_t, c_void_p]),
    "gpa_lawler_workspace_bytes": (c_int, [c_int, c_int, c_int, ctypes.POINTER(c_size_t)]),
    "gpa_invert_u": (c_int, [c_void_p, c_int, c_int, c_double, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gpa_invert_u_plain": (c_int, [c_void_p, c_int, c_int, c_double, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gpa_resample_image": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gpa_undistort_image": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None


class GpaError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GpaError(
            f"{LIB_PATH} not found: the CUDA extension has not been built "
            "(run `make -C pygpa_b200/csrc`). pygpa_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().gpa_last_error()
        raise GpaError(f"libgpa_b200 error {rc}: {msg.decode() if msg else '?'}")


def as_pd(arr):
    return arr.ctypes.data_as(_pd)


def as_pf(arr):
    return arr.ctypes.data_as(_pf)
