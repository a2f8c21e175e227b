"""The one mathtools helper the hot path uses (reference: pyGPA/mathtools.py:72-75)."""
import numpy as np


def wrapToPi(x):
    """Wrap all values of x to the interval [-pi, pi)."""
    return (x + np.pi) % (2 * np.pi) - np.pi
