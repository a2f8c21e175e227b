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


def _rotate(v, angle):
    """Stand-in for latticegen.transformations.rotate (third party, absent offline): only reached
    through calc_diff_from_isotropic (geometric_phase_analysis.py:318), which enumerates ALL rotations
    of one vector by multiples of 2 pi / symmetry and keeps the nearest — the set, and so the result,
    is the same for either sign convention of the rotation."""
    import numpy as np
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[c, -s], [s, c]]) @ np.asarray(v)


def load_property_extract():
    """The reference's pyGPA.property_extract (its latticegen / dask imports stubbed)."""
    gpa, _ = load()
    tr = sys.modules["latticegen.transformations"]
    tr.rotate = _rotate
    # property_extract.py:692-693 wraps two more latticegen helpers in numba.njit at import time (lazy:
    # never compiled unless the Kerelsky fit is called, which the oracle does not do)
    tr.rotation_matrix = lambda angle: None
    tr.strain_matrix = lambda epsilon: None
    # property_extract.py:863 decorates one function with dask.array.as_gufunc (unused here)
    sys.modules["dask.array"].as_gufunc = lambda **kw: (lambda f: f)
    sys.modules["dask"].array = sys.modules["dask.array"]
    sys.modules["latticegen"].transformations = sys.modules["latticegen.transformations"]
    gpa.rotate = _rotate          # `from latticegen.transformations import rotate` bound the stub's None
    import pyGPA.property_extract as pe
    return pe


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
