"""NumPy restatement of the property-extraction consumers of the sweep's phase gradient
(oracle; test infrastructure only).  Reference: pyGPA/property_extract.py:69-101
(phasegradient2J), 55-66 (phasegradient2Jac), 137-178 (props_from_Jac), 218-219 (props_from_J);
pyGPA/geometric_phase_analysis.py:303-323 (average_lattice_vector, calc_diff_from_isotropic);
pyGPA/mathtools.py:6-18 (periodic_average, periodic_difference).

``latticegen.transformations.rotate`` (a third-party helper that is absent offline, no version pin
in the reference's requirements.txt) is only used to enumerate the `symmetry` rotations of one
vector by multiples of 2 pi / symmetry; the SET of rotated vectors does not depend on the sign
convention of the rotation, and only the set enters the result (nearest member), so a plain
counter-clockwise rotation matrix restates it."""
from __future__ import annotations

import numpy as np

from .ref_numpy import TWO_PI, weighted_lstsq, wrap_to_pi

__all__ = ["periodic_average", "periodic_difference", "calc_diff_from_isotropic", "phasegradient2J",
           "phasegradient2Jac", "props_from_Jac", "props_from_J", "svd2x2_lapack"]


def periodic_average(x, period=TWO_PI):
    """mathtools.py:6-10 (unit weights)."""
    y = np.angle(np.exp(1j * TWO_PI / period * np.asarray(x)).mean())
    return y * period / TWO_PI


def periodic_difference(x, y, period=TWO_PI):
    """mathtools.py:13-17."""
    z = np.angle(np.exp(1j * TWO_PI / period * (np.asarray(x) - y)))
    return z * period / TWO_PI


def _rotate(v, angle):
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[c, -s], [s, c]]) @ v


def calc_diff_from_isotropic(ani_ks, symmetry=6):
    """geometric_phase_analysis.py:303-323: dks such that ani_ks + dks is an isotropic lattice with
    the mean length and the periodic-mean orientation of ani_ks."""
    ani_ks = np.asarray(ani_ks, dtype=np.float64)
    dt = periodic_average(np.arctan2(*ani_ks.T[::-1]), period=TWO_PI / symmetry)
    r = np.linalg.norm(ani_ks, axis=1).mean()
    k_hex = r * np.array([np.cos(dt), np.sin(dt)])
    ks_hex = np.array([_rotate(k_hex, i * TWO_PI / symmetry) for i in range(symmetry)])
    alldiffs = ks_hex - ani_ks[:, None]
    argmins = np.linalg.norm(alldiffs, axis=-1).argmin(axis=1)
    return alldiffs[np.arange(len(ani_ks)), argmins]


def phasegradient2J(kvecs, grads, weights, nmperpixel, iso_ref=True, sort=0):
    """property_extract.py:69-101.  grads (3, N, M, 2) as the sweep returns them per peak,
    weights (3, N, M) -> J (N, M, 2, 2), J[..., i, j] = d u_i / d x_j per nm.
    Quirks kept: `order = np.arange(3)` hard-codes three k-vectors; with sort != 0 the gradients and
    k-vectors are re-ordered but the weights are not."""
    kvecs = np.asarray(kvecs, dtype=np.float64)
    grads = np.asarray(grads, dtype=np.float64)
    angles = np.arctan2(*kvecs.T[::-1])
    if sort == 0:
        lkvecs, order = kvecs, np.arange(3)
    else:
        order = np.argsort(sort * periodic_difference(angles, periodic_average(angles)))
        lkvecs = kvecs[order]
    if iso_ref:
        dks = calc_diff_from_isotropic(lkvecs)
        K = TWO_PI * (lkvecs + dks)
        iso_grads = wrap_to_pi(np.stack([g - TWO_PI * dk for g, dk in zip(grads[order], dks)]))
    else:
        K = TWO_PI * kvecs
        iso_grads = grads
    dudx = weighted_lstsq(iso_grads[..., 0], K, weights)
    dudy = weighted_lstsq(iso_grads[..., 1], K, weights)
    J = np.stack([dudx, dudy], axis=-1) / nmperpixel
    return np.moveaxis(J, 0, -2)


def phasegradient2Jac(kvecs, grads, weights, nmperpixel):
    """property_extract.py:55-66."""
    return np.eye(2) + phasegradient2J(kvecs, grads, weights, nmperpixel)


def props_from_Jac(Jac, refangle=0., refscale=1., diff=False):
    """property_extract.py:137-178, verbatim arithmetic on numpy's SVD (LAPACK gesdd).  Note that the
    result depends on the relative sign LAPACK gives the two singular-vector pairs (det U = -1 for
    every matrix with Jac[1, 0] != 0): see svd2x2_lapack."""
    u, s, v = np.linalg.svd(np.asarray(Jac, dtype=np.float64))
    signs = np.sign(u[..., None, [0, 1], [0, 1]])
    v = signs * v
    u = np.swapaxes(signs * u, -1, -2)
    u_p = np.swapaxes(u @ v, -1, -2)
    angle = np.rad2deg(np.arctan2(u_p[..., 1, 0], u_p[..., 0, 0]))
    aniangle = np.rad2deg(np.arctan2(u[..., 1, 0], u[..., 0, 0]))
    if diff:
        aniangle += 90
        alpha = s[..., 0]
    else:
        alpha = s[..., 1]
    kappa = s[..., 0] / s[..., 1]
    aniangle = aniangle % 180
    return np.array([angle + refangle, aniangle, alpha * refscale, kappa])


def props_from_J(J, refangle=0., refscale=1):
    """property_extract.py:218-219."""
    return props_from_Jac(np.asarray(J) + np.eye(2), refangle=refangle, refscale=refscale)


# ------------------------------------------------------------------------------------------------
# LAPACK's 2x2 SVD, restated: what the CUDA kernel has to reproduce, sign conventions included
# ------------------------------------------------------------------------------------------------
_EPS = np.finfo(np.float64).eps / 2          # dlamch('Epsilon')
_UNFL = np.finfo(np.float64).tiny


def _sign(a, b):
    """Fortran SIGN(a, b)."""
    return abs(a) if not np.signbit(b) else -abs(a)


def _dlasv2(f, g, h):
    """LAPACK dlasv2 (SVD of the upper triangular [[f, g], [0, h]]), published algorithm:
    [[csl, snl], [-snl, csl]] @ [[f, g], [0, h]] @ [[csr, -snr], [snr, csr]] = diag(ssmax, ssmin)."""
    ft, fa, ht, ha = f, abs(f), h, abs(h)
    pmax = 1
    swap = ha > fa
    if swap:
        pmax = 3
        ft, ht = ht, ft
        fa, ha = ha, fa
    gt, ga = g, abs(g)
    if ga == 0:
        ssmin, ssmax, clt, crt, slt, srt = ha, fa, 1., 1., 0., 0.
    else:
        gasmal = True
        if ga > fa:
            pmax = 2
            if fa / ga < _EPS:
                gasmal = False
                ssmax = ga
                ssmin = fa / (ga / ha) if ha > 1 else (fa / ga) * ha
                clt, slt, srt, crt = 1., ht / gt, 1., ft / gt
        if gasmal:
            d = fa - ha
            l = 1. if d == fa else d / fa
            m = gt / ft
            t = 2. - l
            mm, tt = m * m, t * t
            s = np.sqrt(tt + mm)
            r = abs(m) if l == 0 else np.sqrt(l * l + mm)
            a = 0.5 * (s + r)
            ssmin, ssmax = ha / a, fa * a
            if mm == 0:
                t = _sign(2., ft) * _sign(1., gt) if l == 0 else gt / _sign(d, ft) + m / t
            else:
                t = (m / (s + t) + m / (r + l)) * (1. + a)
            l = np.sqrt(t * t + 4.)
            crt, srt = 2. / l, t / l
            clt = (crt + srt * m) / a
            slt = (ht / ft) * srt / a
    if swap:
        csl, snl, csr, snr = srt, crt, slt, clt
    else:
        csl, snl, csr, snr = clt, slt, crt, srt
    if pmax == 1:
        tsign = _sign(1., csr) * _sign(1., csl) * _sign(1., f)
    elif pmax == 2:
        tsign = _sign(1., snr) * _sign(1., csl) * _sign(1., g)
    else:
        tsign = _sign(1., snr) * _sign(1., snl) * _sign(1., h)
    ssmax = _sign(ssmax, tsign)
    ssmin = _sign(ssmin, tsign * _sign(1., f) * _sign(1., h))
    return ssmin, ssmax, snr, csr, snl, csl


def svd2x2_lapack(A):
    """(u, s, vt) of one real 2x2 matrix the way numpy.linalg.svd (LAPACK dgesdd, path M >= N) builds
    them: Householder reflector H1 zeroing A[1, 0] (dgebrd/dlarfg: beta = -sign(a00) |column|, so
    det U = -1 whenever A[1, 0] != 0), dbdsdc's max-norm scaling, dbdsqr's deflation of a negligible
    superdiagonal, dlasv2 for the 2x2 block, negative singular values flipped onto the rows of VT,
    descending sort by swapping.  Verified against numpy.linalg.svd in tests/test_props_oracle.py."""
    a, b, c, d = (float(A[0][0]), float(A[0][1]), float(A[1][0]), float(A[1][1]))
    if c == 0:
        Q = np.eye(2)
        d1, e, d2 = a, b, d
    else:
        beta = -_sign(np.hypot(a, c), a)
        tau = (beta - a) / beta
        v = c / (a - beta)
        w = b + v * d
        e = b - tau * w
        d2 = d - tau * v * w
        d1 = beta
        Q = np.eye(2) - tau * np.array([[1., v], [v, v * v]])
    nrm = max(abs(d1), abs(d2), abs(e))
    U, VT = np.eye(2), np.eye(2)
    if nrm == 0:
        return Q, np.zeros(2), VT
    d1s, d2s, es = d1 / nrm, d2 / nrm, e / nrm
    tol = max(10., min(100., _EPS ** (-0.125))) * _EPS
    smin = abs(d1s)
    if smin != 0:
        smin = min(smin, abs(d2s) * (smin / (smin + abs(es))))
    thresh = max(tol * smin / np.sqrt(2.), 6 * 2 * 2 * _UNFL)
    if abs(es) <= thresh:
        sv = [d1s, d2s]
    else:
        ssmin, ssmax, snr, csr, snl, csl = _dlasv2(d1s, es, d2s)
        sv = [ssmax, ssmin]
        VT = np.array([[csr, snr], [-snr, csr]])
        U = np.array([[csl, -snl], [snl, csl]])
    for i in range(2):
        if np.signbit(sv[i]):
            sv[i] = -sv[i]
            VT[i] = -VT[i]
    if sv[0] < sv[1]:
        sv, U, VT = sv[::-1], U[:, ::-1], VT[::-1]
    return Q @ U, np.array(sv) * nrm, VT
