"""Import the UNMODIFIED reference (pyGPA) from /root/reference for oracle pinning.

Only used in the build container (the GPU box has no /root/reference): by
oracle/gen_golden.py and tests/test_oracle_vs_reference.py.  The reference's
top-level imports of packages that are absent offline and unused on the hot path
(matplotlib, dask, moisan2011, skimage, latticegen) are satisfied with empty modules.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PYGPA_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyGPA"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def load():
    """Returns (geometric_phase_analysis, phase_unwrap) modules of the reference."""
    if not available():
        raise ImportError("reference checkout not found at " + REFERENCE_ROOT)
    for name, attrs in [
        ("matplotlib", {}), ("matplotlib.pyplot", {}), ("mpl_toolkits", {}),
        ("mpl_toolkits.axes_grid1", {}),
        ("mpl_toolkits.axes_grid1.inset_locator", {"inset_axes": None}),
        ("dask", {}), ("dask.array", {}), ("moisan2011", {"per": None}),
        ("skimage", {}), ("skimage.feature", {"peak_local_max": None}),
        ("skimage.restoration", {"wiener": None}), ("skimage.morphology", {"disk": None}),
        ("latticegen", {}), ("latticegen.transformations", {"rotate": None}),
    ]:
        try:
            __import__(name)
        except Exception:
            _stub(name, **attrs)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import pyGPA.geometric_phase_analysis as gpa
    import pyGPA.phase_unwrap as pu
    return gpa, pu
