"""pygpa_b200 — B200-native (sm_100a) implementation of pyGPA's adaptive-GPA hot path.

Modules mirror the reference package for the functions on that path:
  cuGPA                      <- pyGPA/cuGPA.py
  geometric_phase_analysis   <- pyGPA/geometric_phase_analysis.py (lock-in, wfr*, reconstruct_u_inv*, Lawler-Fujita)
  phase_unwrap               <- pyGPA/phase_unwrap.py
Importing the package does not touch the GPU; the first call loads libgpa_b200.so and
raises if it is missing or no CUDA device is present (no CPU fallback).
"""
__version__ = "0.1.0"
