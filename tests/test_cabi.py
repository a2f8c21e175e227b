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
