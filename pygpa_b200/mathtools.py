"""The mathtools helpers the hot path uses (reference: pyGPA/mathtools.py)."""
import numpy as np


def wrapToPi(x):
    """Wrap all values of x to the interval [-pi, pi) (mathtools.py:72-75)."""
    return (x + np.pi) % (2 * np.pi) - np.pi


def fit_plane(image, verbose=False):
    """Fit the plane a[0]*x + a[1]*y + a[2] through `image` (x, y = array indices) with a Huber loss
    (mathtools.py:30-47).  The reference minimises with scipy.optimize.least_squares; here the same
    minimiser is reached by iteratively reweighted least squares on the GPU (K6, refine.cu)."""
    from . import solvers
    dev = solvers.require_cuda()
    theta, iters = solvers.fit_plane_huber(solvers.to_device_f64(image, dev), return_iters=True)
    if verbose:
        print(f"Huber plane fit converged after {iters} reweighting steps")
    return theta
