"""The per-pixel code of the property-extraction kernels (pygpa_b200/csrc/props_device.cuh,
lsq_device.cuh), compiled for the HOST with g++ and checked on the CPU: LAPACK's sign conventions
for the 2x2 SVD against numpy.linalg.svd, props_from_Jac against the reference-generated fixture,
the weighted least-squares solve against the oracle.  The -m gpu tests run the same source as CUDA."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import ROOT, load_golden

_pd = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("harness") / "props_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                           os.path.join(ROOT, "tests", "cpu_harness", "props_host.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    lib.host_svd2x2.argtypes = [_pd, ctypes.c_long, _pd, _pd, _pd]
    lib.host_props.argtypes = [_pd, ctypes.c_long, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, _pd]
    lib.host_lsq2.argtypes = [_pd, _pd, ctypes.c_long, ctypes.c_int, _pd]
    return lib


def _p(a):
    return a.ctypes.data_as(_pd)


def _svd(host, mats):
    mats = np.ascontiguousarray(mats, dtype=np.float64)
    n = mats.shape[0]
    u, s, vt = np.empty((n, 2, 2)), np.empty((n, 2)), np.empty((n, 2, 2))
    host.host_svd2x2(_p(mats), n, _p(u), _p(s), _p(vt))
    return u, s, vt


def _props(host, jac, refangle=0., refscale=1., diff=False, add_identity=False):
    jac = np.ascontiguousarray(jac, dtype=np.float64).reshape(-1, 2, 2)
    out = np.empty((4, jac.shape[0]))
    host.host_props(_p(jac), jac.shape[0], refangle, refscale, int(diff), int(add_identity), _p(out))
    return out


def _rot(t):
    return np.array([[np.cos(t), -np.sin(t)], [np.sin(t), np.cos(t)]])


def test_svd_conventions_match_numpy(host):
    rng = np.random.default_rng(1)
    mats = [np.eye(2) + 0.3 * rng.normal(size=(2, 2)) for _ in range(5000)]
    mats += [rng.normal(size=(2, 2)) * 10 ** rng.uniform(-3, 3) for _ in range(5000)]
    for _ in range(500):           # upper triangular, diagonal, negligible superdiagonal, rotated anisotropy
        a, b, d = rng.normal(size=3)
        mats += [np.array([[a, b], [0, d]]), np.array([[a, 0], [0, d]]), np.array([[a, 1e-17 * b], [0, d]]),
                 _rot(rng.uniform(-3, 3)) @ _rot(b).T @ np.diag([1 + abs(a), 1]) @ _rot(b)]
    mats += [np.eye(2), -np.eye(2), np.zeros((2, 2)), np.array([[0, 1.], [1, 0]]), np.array([[1., 0], [0, -1]])]
    mats = np.array(mats)
    u, s, vt = _svd(host, mats)
    ru, rs, rvt = np.linalg.svd(mats)
    assert np.abs(s - rs).max() <= 1e-13 * rs.max()
    assert np.abs(u - ru).max() < 1e-9 and np.abs(vt - rvt).max() < 1e-9
    # and the python restatement in the oracle is the same algorithm
    for m in mats[::97]:
        ou, os_, ovt = oracle.svd2x2_lapack(m)
        i = np.flatnonzero((mats == m).all(axis=(1, 2)))[0]
        assert np.allclose(ou, u[i], atol=1e-12) and np.allclose(ovt, vt[i], atol=1e-12) and np.allclose(os_, s[i])


def test_lapack_gives_det_u_minus_one_and_props_depend_on_it(host):
    """Why the kernel cannot use a textbook SVD: flipping one singular pair changes 'angle'."""
    rng = np.random.default_rng(3)
    mats = np.eye(2) + 0.2 * rng.normal(size=(200, 2, 2))
    u, s, vt = _svd(host, mats)
    assert np.all(np.linalg.det(u) < 0)
    assert np.allclose(u * s[:, None, :] @ vt, mats)


def test_props_match_reference_fixture(host):
    g = load_golden("props_64x48.npz")
    jac = g["out_Jac"]
    shape = jac.shape[:2]
    got = _props(host, jac).reshape((4,) + shape)
    assert np.allclose(got, g["out_props"], rtol=1e-9, atol=1e-7)
    got = _props(host, jac, 3.0, 2.0, True).reshape((4,) + shape)
    assert np.allclose(got, g["out_props_diff"], rtol=1e-9, atol=1e-7)
    got = _props(host, g["out_J_rankdef"], add_identity=True).reshape((4,) + shape)
    assert np.allclose(got, g["out_props_rankdef"], rtol=1e-9, atol=1e-7)
    assert np.array_equal(got[:, 0, 0], [0., 0., 1., 1.])          # zero weights: Jac = identity exactly
    got = _props(host, g["out_J_iso"], -1.5, 0.7, add_identity=True).reshape((4,) + shape)
    assert np.allclose(got, g["out_props_from_J"], rtol=1e-9, atol=1e-7)


def test_props_random_jacobians_against_oracle(host):
    rng = np.random.default_rng(5)
    jac = np.stack([_rot(np.deg2rad(t)) @ _rot(p).T @ np.diag([k, 1.0]) @ _rot(p) * a
                    for t, p, k, a in zip(rng.uniform(-20, 20, 3000), rng.uniform(-3, 3, 3000),
                                          1 + rng.uniform(1e-3, 0.5, 3000), rng.uniform(0.8, 1.2, 3000))])
    for kw in (dict(), dict(diff=True, refangle=10.0, refscale=3.0)):
        ref = oracle.props_from_Jac(jac, **kw)
        got = _props(host, jac, kw.get("refangle", 0.), kw.get("refscale", 1.), kw.get("diff", False))
        d = np.abs(got - ref)
        d[1] = np.minimum(d[1], 180 - d[1])        # aniangle lives on a circle of 180 degrees
        assert d.max() < 1e-8


def test_weighted_lstsq_pair_against_oracle(host):
    rng = np.random.default_rng(7)
    n, d = 4000, 3
    K = 2 * np.pi * rng.normal(size=(d, 2)) * 0.1
    w = rng.uniform(0, 1, size=(n, d))
    w[:50] = 0.0                       # rank 0
    w[50:100, 1:] = 0.0                # rank 1
    b = rng.normal(size=(n, 2, d))
    a = np.ascontiguousarray(w[:, :, None] * K[None])
    y = np.ascontiguousarray(w[:, None, :] * b)
    x = np.empty((n, 2, 2))
    host.host_lsq2(_p(a), _p(y), n, d, _p(x))
    for rhs in range(2):
        ref = oracle.weighted_lstsq(b[:, rhs, :].T.reshape(d, n, 1), K, w.T.reshape(d, n, 1))[:, :, 0].T
        assert np.allclose(x[:, rhs, :], ref, rtol=1e-9, atol=1e-11)
    assert np.all(x[:50] == 0.0)
