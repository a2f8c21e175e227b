"""Drop-in for the per-pixel part of ``pyGPA.property_extract`` on B200:
the consumers of the adaptive sweep's phase-gradient maps.

    phasegradient2J / phasegradient2Jac   property_extract.py:69-101, 55-66
    props_from_Jac / props_from_J         property_extract.py:137-178, 218-219

Same names, arguments, NumPy-in / NumPy-out types as the reference.  The O(#k) host algebra of
the reference (ordering the k-vectors, the isotropic reference lattice) stays NumPy on the host;
everything per pixel runs in libgpa_b200.so (K5, pygpa_b200/csrc/props.cu).  Kerelsky fits,
u2J and the dask gufunc wrappers of the reference module are not part of this package.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine, solvers
from .cuGPA import _to_host

__all__ = ["periodic_average", "periodic_difference", "calc_diff_from_isotropic", "phasegradient2J",
           "phasegradient2Jac", "props_from_Jac", "props_from_J"]

_TWO_PI = 2 * np.pi


def periodic_average(X, period=_TWO_PI):
    """mathtools.py:6-10 (host scalars)."""
    return np.angle(np.exp(1j * _TWO_PI / period * np.asarray(X)).mean()) * period / _TWO_PI


def periodic_difference(X, Y, period=_TWO_PI):
    """mathtools.py:13-17 (host scalars)."""
    return np.angle(np.exp(1j * _TWO_PI / period * (np.asarray(X) - Y))) * period / _TWO_PI


def calc_diff_from_isotropic(ani_ks, symmetry=6):
    """geometric_phase_analysis.py:310-323: dks with ani_ks + dks an isotropic lattice (mean length,
    periodic-mean orientation).  The reference enumerates the `symmetry` rotations of the mean
    vector with latticegen's rotate; only the set of rotated vectors matters."""
    ani_ks = np.asarray(ani_ks, dtype=np.float64)
    dt = periodic_average(np.arctan2(*ani_ks.T[::-1]), period=_TWO_PI / symmetry)
    r = np.linalg.norm(ani_ks, axis=1).mean()
    k_hex = r * np.array([np.cos(dt), np.sin(dt)])
    rots = [np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]) for a in _TWO_PI / symmetry * np.arange(symmetry)]
    ks_hex = np.array([R @ k_hex for R in rots])
    alldiffs = ks_hex - ani_ks[:, None]
    argmins = np.linalg.norm(alldiffs, axis=-1).argmin(axis=1)
    return alldiffs[np.arange(len(ani_ks)), argmins]


def _host(t):
    h = _to_host(t)
    torch.cuda.current_stream().synchronize()
    return h.numpy()


def _solve_setup(kvecs, iso_ref, sort):
    """Host part of property_extract.py:78-94: (K, sub, order)."""
    kvecs = np.asarray(kvecs, dtype=np.float64)
    if kvecs.shape != (3, 2):
        raise ValueError("phasegradient2J needs exactly three k-vectors (the reference hard-codes np.arange(3))")
    angles = np.arctan2(*kvecs.T[::-1])
    if sort == 0:
        lkvecs, order = kvecs, np.arange(3)
    else:
        order = np.argsort(sort * periodic_difference(angles, periodic_average(angles)))
        lkvecs = kvecs[order]
    if iso_ref:
        dks = calc_diff_from_isotropic(lkvecs)
        return _TWO_PI * (lkvecs + dks), _TWO_PI * dks, order
    return _TWO_PI * kvecs, None, None       # the reference ignores `sort` without iso_ref


def phasegradient2J_device(kvecs, grads, weights, nmperpixel, iso_ref=True, sort=0, add_identity=False):
    """Device-resident form: CUDA tensors in, (N, M, 2, 2) CUDA tensor out."""
    K, sub, order = _solve_setup(kvecs, iso_ref, sort)
    return solvers.phasegradient_to_J(grads, weights, K, sub, order, do_wrap=iso_ref, nmperpixel=nmperpixel,
                                      add_identity=add_identity)


def phasegradient2J(kvecs, grads, weights, nmperpixel, iso_ref=True, sort=0):
    """J (N, M, 2, 2) directly from the sweep's phase gradients (property_extract.py:69-101)."""
    dev = engine.require_cuda()
    return _host(phasegradient2J_device(kvecs, solvers.to_device_f64(grads, dev), solvers.to_device_f64(weights, dev),
                                        nmperpixel, iso_ref, sort))


def phasegradient2Jac(kvecs, grads, weights, nmperpixel):
    """property_extract.py:55-66: identity + phasegradient2J."""
    dev = engine.require_cuda()
    return _host(phasegradient2J_device(kvecs, solvers.to_device_f64(grads, dev), solvers.to_device_f64(weights, dev),
                                        nmperpixel, add_identity=True))


def props_from_Jac(Jac, refangle=0., refscale=1., diff=False):
    """(angle, aniangle, alpha, kappa) of a (..., 2, 2) Jacobian field (property_extract.py:137-178)."""
    dev = engine.require_cuda()
    return _host(solvers.props_from_jac(solvers.to_device_f64(Jac, dev), refangle, refscale, diff))


def props_from_J(J, refangle=0., refscale=1):
    """property_extract.py:218-219."""
    dev = engine.require_cuda()
    return _host(solvers.props_from_jac(solvers.to_device_f64(J, dev), refangle, refscale, False, add_identity=True))
