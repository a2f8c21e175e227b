import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests need a CUDA device and the built library: on a CPU box they are skipped (not failed) unless
    the run selects them explicitly with -m gpu (the GPU box's run, where a missing device must fail loudly)."""
    import torch
    if "gpu" in (config.getoption("-m") or ""):
        return
    lib = os.path.join(ROOT, "pygpa_b200", "libgpa_b200.so")
    if torch.cuda.is_available() and os.path.exists(lib):
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and pygpa_b200/libgpa_b200.so (no CPU fallback by design)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


@pytest.fixture(scope="session")
def golden():
    return load_golden
