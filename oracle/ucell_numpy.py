"""NumPy restatement of pyGPA/unit_cell_averaging.py (oracle; test infrastructure only):
calc_ucell_parameters (:45-53), unit_cell_average (:132-205) with add_to_position / float_overlap
(:208-217, :37-42), expand_unitcell (:234-249)."""
from __future__ import annotations

import numpy as np
import scipy.ndimage as ndi

__all__ = ["calc_ucell_parameters", "cart_in_uc", "unit_cell_average", "expand_unitcell"]


def calc_ucell_parameters(ks, z):
    """unit_cell_averaging.py:45-53: lower corner and zoomed array size of the unit cell."""
    corners = np.array([[0., 0.], [0., 1.], [1., 0.], [1., 1.]])
    cornervals = corners @ np.linalg.inv(ks).T
    rmin = cornervals.min(axis=0)
    rsize = tuple((z * np.ceil(cornervals.max(axis=0) - np.floor(rmin))).astype(int))
    return rmin, rsize


def cart_in_uc(vecs, ks, rmin=0):
    """unit_cell_averaging.py:29-34."""
    return ((vecs @ ks.T) % 1.) @ np.linalg.inv(ks).T - rmin


def unit_cell_average(image, ks, u=None, z=1):
    """unit_cell_averaging.py:132-205.  Quirk kept: float_overlap pairs the axis-0 offset with the
    axis-1 fraction and vice versa (overlap[li, lj] = (lj ? f0 : 1-f0) * (li ? f1 : 1-f1))."""
    image = np.asarray(image, dtype=np.float64)
    ks = np.asarray(ks, dtype=np.float64)
    rmin, rsize = calc_ucell_parameters(ks, z)
    n, m = image.shape
    rr = np.stack(np.meshgrid(np.arange(n), np.arange(m), indexing='ij'), axis=-1).astype(np.float64)
    if u is not None:
        rr = rr + np.moveaxis(np.asarray(u, dtype=np.float64), 0, -1)
    R = cart_in_uc(rr, ks, rmin) * z
    keep = ~np.isnan(image)
    R, vals = R[keep], image[keep]
    Rf = np.floor(R)
    f = R - Rf
    Ri = Rf.astype(np.int64)
    res = np.zeros(rsize)
    weights = np.zeros(rsize)
    for li in range(2):
        for lj in range(2):
            ov = (f[:, 0] if lj else 1 - f[:, 0]) * (f[:, 1] if li else 1 - f[:, 1])
            a, b = Ri[:, 0] + li, Ri[:, 1] + lj
            ok = (a >= 0) & (a < rsize[0]) & (b >= 0) & (b < rsize[1])
            np.add.at(res, (a[ok], b[ok]), (vals * ov)[ok])
            np.add.at(weights, (a[ok], b[ok]), ov[ok])
    with np.errstate(invalid='ignore', divide='ignore'):
        return res / weights


def expand_unitcell(unit_cell_image, ks, shape, z=1, z2=1, u=0):
    """unit_cell_averaging.py:234-249."""
    ks = np.asarray(ks, dtype=np.float64)
    rr = np.mgrid[:shape[0], :shape[1]] / z2
    rrt = np.moveaxis(rr + u, 0, -1)
    rmin, _ = calc_ucell_parameters(ks, z)
    X = cart_in_uc(rrt, ks, rmin) * z
    return ndi.map_coordinates(np.nan_to_num(unit_cell_image), np.moveaxis(X, -1, 0), cval=0)
