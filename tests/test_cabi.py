"""The C-ABI shared library loads and exports everything include/gpa_b200.h declares.
No compute calls (no GPU needed)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from pygpa_b200 import _lib


def _declared():
    text = open(os.path.join(ROOT, "include", "gpa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpa_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gpa_b200.h but not exported"
    # and the ctypes prototypes cover exactly the header
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.gpa_version() >= 100
    nbytes = ctypes.c_size_t(0)
    rc = lib.gpa_lockin_workspace_bytes(1, 1, 1, 1, 0, 0, 1, ctypes.byref(nbytes))   # 1x1 frame: invalid
    assert rc == -1
    assert b"2x2" in lib.gpa_last_error()
    rc = lib.gpa_lockin_workspace_bytes(64, 64, 3, 3, 40, 4, 1, ctypes.byref(nbytes))  # filter wider than the frame
    assert rc == -1 and b"fit the frame" in lib.gpa_last_error()
    rc = lib.gpa_lockin_workspace_bytes(2048, 2048, 41, 41, 45, 45, 41, ctypes.byref(nbytes))
    assert rc == 0
    planes = 41 * (2048 + 91) * 2048 * 8
    assert planes <= nbytes.value < planes * 1.05
    with pytest.raises(_lib.GpaError):
        _lib.check(-1)


def test_new_entry_points_validate_arguments_without_a_gpu():
    """Argument checks of the 8(f) entry points run before any CUDA call: invalid input -> GPA_ERR_INVALID
    (-1) and a message, never a crash."""
    lib = _lib.load()
    nbytes = ctypes.c_size_t(0)
    null = ctypes.c_void_p(0)
    assert lib.gpa_props_from_jac(null, 10, 0.0, 1.0, 0, 0, null, null) == -1
    assert b"null" in lib.gpa_last_error()
    k = (ctypes.c_double * 6)(*([0.1] * 6))
    assert lib.gpa_phasegradient_to_j(null, null, 4, 4, k, None, None, 0, 3, 4, 4, 1.0, 0, null, null) == -1
    assert lib.gpa_deconvolve_workspace_bytes(4100, 64, 8, ctypes.byref(nbytes)) == -1      # 4132 > 4096 samples
    assert b"exceeds" in lib.gpa_last_error()
    assert lib.gpa_deconvolve_workspace_bytes(12, 64, 8, ctypes.byref(nbytes)) == -1        # reflect padding too wide
    assert lib.gpa_deconvolve_workspace_bytes(2048, 2048, 20, ctypes.byref(nbytes)) == 0
    assert nbytes.value >= 2 * 2128 * 2128 * 16
    assert lib.gpa_uc_workspace_bytes(0, 5, ctypes.byref(nbytes)) == -1
    assert lib.gpa_uc_workspace_bytes(100, 120, ctypes.byref(nbytes)) == 0 and nbytes.value >= 4 * 100 * 120 * 8
    assert lib.gpa_fit_plane_workspace_bytes(ctypes.byref(nbytes)) == 0 and nbytes.value > 19 * 1184 * 8
    th = (ctypes.c_double * 3)()
    assert lib.gpa_fit_plane_huber(null, 8, 8, 1.0, 10, 1e-9, th, None, null, 0, null) == -1
    assert lib.gpa_wfr4_sweep(null, 64, 64, k, k, 3, null, None, 2, None, 2, 0.0, 0.0, 1, null, null, null, null,
                              null, 0, null) == -1
    assert lib.gpa_unwrap_workspace_bytes(1014, 1014, ctypes.byref(nbytes)) == 0             # Bluestein tables included
    assert nbytes.value > 7 * 1014 * 1014 * 8 + 2 * 2048 * 16


def test_split_pass2_geometry_is_validated_without_a_gpu():
    """gpa_sweep_mr_workspace_bytes with the split pass 2: bad stage radii are rejected, a valid split
    needs more scratch than the single-stage form (halo rows of P1 + the anchor stage's output)."""
    lib = _lib.load()
    nbytes, single = ctypes.c_size_t(0), ctypes.c_size_t(0)
    base = (2048, 2048, 41, 41, 0, 4, 43, 43, 20)
    assert lib.gpa_sweep_mr_workspace_bytes(*base, 0, 0, 0, 0, 41, ctypes.byref(single)) == 0
    assert lib.gpa_sweep_mr_workspace_bytes(*base, 43, 7, 0, 0, 41, ctypes.byref(nbytes)) == 0
    extra = nbytes.value - single.value
    nbytes_x = nbytes.value
    assert 41 * 2 * (512 + 14) * 512 * 8 <= extra <= 41 * (2 * (512 + 14) * 512 + 600 * 512) * 8
    assert lib.gpa_sweep_mr_workspace_bytes(*base, 43, 12, 0, 0, 41, ctypes.byref(nbytes)) == -1       # more than 23 coarse taps
    assert b"coarse radius" in lib.gpa_last_error()
    assert lib.gpa_sweep_mr_workspace_bytes(2048, 2048, 41, 41, 1, 4, 43, 43, 20, 43, 7, 0, 0, 41, ctypes.byref(nbytes)) == -1   # k-list
    assert lib.gpa_sweep_mr_workspace_bytes(88, 88, 41, 41, 0, 4, 43, 43, 20, 43, 7, 0, 0, 41, ctypes.byref(nbytes)) == -1       # edge bands overlap
    both = ctypes.c_size_t(0)
    assert lib.gpa_sweep_mr_workspace_bytes(*base, 43, 7, 43, 7, 41, ctypes.byref(both)) == 0          # + split pass 1
    assert 2 * 2192 * 544 * 8 <= both.value - nbytes_x <= 2 * 2208 * 544 * 8 + 3 * 41 * 600 * 8 + 8192
    assert lib.gpa_sweep_mr_workspace_bytes(*base, 43, 7, 43, 12, 41, ctypes.byref(both)) == -1
    r1, h, s1 = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_double(0)
    assert lib.gpa_split_plan(2048, 3, 8.98, None, 0, ctypes.byref(r1), ctypes.byref(h), ctypes.byref(s1), None, None) == -1


def test_integration_stub_prototypes_match_the_binding():
    """The ctypes prototypes printed in INTEGRATION.md (the stub a pyGPA maintainer would add) are the
    ones pygpa_b200/_lib.py declares — same arity, same C types in the same order."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    env = {"ctypes": ctypes, "_pd": ctypes.POINTER(ctypes.c_double), "_pf": ctypes.POINTER(ctypes.c_float),
           "_vp": ctypes.c_void_p, "_i": ctypes.c_int, "_d": ctypes.c_double}
    found = re.findall(r"^L\.(gpa_[a-z0-9_]+)\.argtypes = (\[.*?\])\n(?=[A-Za-z\n])", text, flags=re.S | re.M)
    assert len(found) >= 6
    for name, expr in found:
        got = eval(expr, env)          # noqa: S307 - our own documentation
        want = _lib.SIGNATURES[name][1]
        assert len(got) == len(want), f"{name}: INTEGRATION.md lists {len(got)} arguments, the binding {len(want)}"
        for a, b in zip(got, want):
            assert ctypes.sizeof(a) == ctypes.sizeof(b) and (a is b or a._type_ == b._type_), f"{name}: {a} vs {b}"


def test_header_prototypes_match_the_binding_arity():
    """Every prototype of include/gpa_b200.h has as many parameters as the ctypes binding passes."""
    text = open(os.path.join(ROOT, "include", "gpa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = re.findall(r"\b(gpa_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text)
    assert len(protos) >= 40
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert name in _lib.SIGNATURES, f"{name} is declared but not bound"
        assert n == len(_lib.SIGNATURES[name][1]), f"{name}: header has {n} parameters, the binding {len(_lib.SIGNATURES[name][1])}"
