"""CPU oracle for the adaptive-GPA hot path (TEST INFRASTRUCTURE ONLY).

This package is a NumPy/SciPy restatement of the algorithms in the reference
``pyGPA`` (geometric_phase_analysis.py / phase_unwrap.py / cuGPA.py).  It exists to
check the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package ``pygpa_b200`` never does (tests/test_host_logic.py::test_product_never_imports_the_oracle
enforces that).

Parity status: PINNED (except oracle/wiener_numpy.py, whose third-party arithmetic — scikit-image's
Wiener filter — is absent here: that module says "parity unpinned" in its header).  ``tests/golden/*.npz`` were produced by the unmodified
reference functions (``oracle/gen_golden.py`` imports ``/root/reference``); the
oracle is asserted against them in ``tests/test_oracle_golden.py`` and, when the
reference checkout is present, against the live reference in
``tests/test_oracle_vs_reference.py``.
"""
from .ref_numpy import *  # noqa: F401,F403
from .props_numpy import *  # noqa: F401,F403
from .ucell_numpy import *  # noqa: F401,F403
from .wiener_numpy import *  # noqa: F401,F403
