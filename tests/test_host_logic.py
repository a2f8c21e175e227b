"""Host-side logic that needs no GPU: taps, candidate axes, argument handling, hygiene."""
import os
import re

import numpy as np
import pytest
import scipy.ndimage as ndi
import torch

import oracle
from conftest import ROOT
from pygpa_b200 import _lib, _taps, cuGPA, engine, synth


@pytest.mark.parametrize("n,m,sigma", [(64, 48, 4.0), (200, 256, 10), (33, 40, 2.5)])
def test_taps_reproduce_scipy_fourier_gaussian(n, m, sigma):
    tx, rx = _taps.axis_taps(n, sigma)
    ty, ry = _taps.axis_taps(m, sigma)
    kernel = np.fft.ifft2(ndi.fourier_gaussian(np.ones((n, m)), sigma)).real
    sub = kernel[np.ix_(np.arange(-rx, rx + 1) % n, np.arange(-ry, ry + 1) % m)]
    assert np.abs(np.outer(tx, ty) - sub).max() < 1e-7 * kernel.max() + 1e-9
    assert rx == min(int(np.ceil(4.5 * sigma)), (n - 1) // 2)


def test_taps_clamped_to_frame_and_limited():
    t, r = _taps.axis_taps(21, 10)
    assert r == 10 and len(t) == 21 and abs(t.sum() - 1) < 1e-6      # whole circle: exact filter
    with pytest.raises(ValueError):
        _taps.axis_taps(4096, 60)


def test_candidate_axes_match_oracle_expression():
    ks = synth.primary_ks(0.1, 7.0, 3)
    kw = np.linalg.norm(ks, axis=1).mean() / 2.5
    for k in ks:
        for kstep in (kw / 3, 2 * kw / 21, 2 * kw / 20.5):
            a = engine.grid_axes(k[0], k[1], kw, kstep)
            b = oracle.candidate_axes(k[0], k[1], kw, kstep)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    for n in (7, 21, 41):
        kw, kstep = synth.sweep_params(ks, n)
        wx, wy = engine.grid_axes(ks[0][0], ks[0][1], kw, kstep)
        assert len(wx) == n and len(wy) == n


def test_grad_argument_handling():
    assert cuGPA._grad_mode(None) == engine.GRAD_CENTRAL
    assert cuGPA._grad_mode('diff') == engine.GRAD_FORWARD
    assert cuGPA._grad_mode(lambda p: np.gradient(p)) is None      # a callable runs unfused (cuGPA._grad_by_winner)
    with pytest.raises(ValueError):
        cuGPA._grad_mode('central')


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    with pytest.raises(_lib.GpaError):
        cuGPA.wfr2_grad_opt(np.zeros((32, 32)), 3, 0.1, 0.0, 0.02, 0.01)


def test_synth_is_deterministic():
    a = synth.make_config('C2', size=64, n_grid=5)
    b = synth.make_config('C2', size=64, n_grid=5)
    assert np.array_equal(a['image'], b['image']) and abs(a['image'].mean()) < 1e-12


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|/root/reference|_refimport", re.M)
    for dirpath, _dirs, files in os.walk(os.path.join(ROOT, "pygpa_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), f"{f} references the oracle / reference"


@pytest.mark.parametrize("n,sigma", [(64, 4.0), (500, 10.0), (2048, 8.98), (2048, 4.4), (33, 2.5)])
def test_c_side_taps_and_multirate_plan_agree_with_python(n, sigma):
    """The C ABI's host helpers (for non-Python hosts) reproduce pygpa_b200/_taps.py."""
    import ctypes
    lib = _lib.load()
    ref, r = _taps.axis_taps(n, sigma)
    assert lib.gpa_default_radius(n, sigma, 4.5) == r
    got = np.zeros(2 * r + 1, dtype=np.float32)
    assert lib.gpa_gaussian_taps(n, sigma, r, _lib.as_pf(got)) == 0
    assert np.abs(got - ref).max() <= 2e-7 * ref.max()
    for size, sg in ((2048, 10.0), (1024, 10.0), (256, 5.0), (2048, 22.0), (64, 4.0), (500, 9.0), (250, 10.0),
                     (2048, 29.0), (2048, 30.0), (2048, 35.0), (2048, 49.0)):
        sa, sb = ctypes.c_double(0), ctypes.c_double(0)
        ra, rb = ctypes.c_int(0), ctypes.c_int(0)
        s_c = lib.gpa_multirate_plan(size, size, sg, ctypes.byref(sa), ctypes.byref(sb), ctypes.byref(ra), ctypes.byref(rb))
        mr = _taps.multirate_taps(size, size, sg)
        assert s_c == (mr["S"] if mr else 0)
        if mr:
            assert (ra.value, rb.value) == (mr["Ra_x"], mr["Rb"]) and abs(sa.value - mr["sigma_a"]) < 1e-12


def test_split_pass2_plan():
    """Planner of the split pass 2 (shared anchor stage + coarse-rate stage per candidate): narrow
    grids get a factorisation whose worst-case transfer-function error stays below the truncation
    error of the single-stage filter; wide grids and short candidate axes get none."""
    mr = _taps.multirate_taps(256, 256, 10.0)
    ks = synth.primary_ks(0.05, 7.0, 3)
    kw, kstep = synth.sweep_params(ks, 21)
    wx = np.arange(ks[0][0] - kw, ks[0][0] + kw, kstep)
    sp = _taps.split_taps(256, mr, wx)
    assert sp is not None and 2 * sp["H"] + 1 in (13, 15, 17, 19, 21, 23) and sp["err"] <= _taps.SPLIT_TOL
    assert abs(sp["sigma_1"] ** 2 + sp["sigma_2"] ** 2 - mr["sigma_a"] ** 2) < 1e-9
    assert sp["taps_1"].size == 2 * sp["R1"] + 1 and sp["taps_2"].size == 2 * sp["H"] + 1
    assert abs(sp["taps_1"].sum() - 1) < 1e-5 and abs(sp["taps_2"].sum() - 1) < 1e-4
    assert -(-(2 * sp["R1"] + 1) // mr["S"]) % 2 == 0          # even taps per phase: statically scheduled stage A
    # independent 1-D check of the whole pipeline against the exact circular Gaussian, border rows included
    n, S, H, R1 = 256, mr["S"], sp["H"], sp["R1"]
    rng = np.random.default_rng(3)
    p1 = rng.normal(size=n) + 1j * rng.normal(size=n)
    f = np.fft.fftfreq(n)
    wx0 = wx[wx.size // 2]
    nd, nde, rtot = n // S, n // S + 2 * H, R1 + S * H
    xu = np.arange(S * nde + 2 * R1 + 1) - rtot
    a = p1[xu % n] * np.exp(2j * np.pi * wx0 * (xu % n))
    body = (xu >= 0) & (xu < n)
    e = np.arange(nde)
    A_body = sum(sp["taps_1"][i].astype(float) * np.where(body[S * e + i], a[S * e + i], 0) for i in range(2 * R1 + 1))
    A_edge = sum(sp["taps_1"][i].astype(float) * np.where(body[S * e + i], 0, a[S * e + i]) for i in range(2 * R1 + 1))
    for w in (wx[0], wx[5], wx[-1]):
        exact = np.fft.ifft(np.fft.fft(p1 * np.exp(2j * np.pi * w * np.arange(n))) * np.exp(-2 * np.pi ** 2 * mr["sigma_a"] ** 2 * f ** 2))[::S]
        dw = w - wx0
        delta = dw * mr["sigma_a"] ** 2 / sp["sigma_2"] ** 2
        c = np.exp(2 * np.pi ** 2 * dw ** 2 * mr["sigma_a"] ** 2 * sp["sigma_1"] ** 2 / sp["sigma_2"] ** 2)
        J = np.where(e - H < nd // 2, np.exp(2j * np.pi * dw * n), np.exp(-2j * np.pi * dw * n))
        smp = np.exp(2j * np.pi * delta * S * (e - H)) * (A_body + J * A_edge)
        mx = np.arange(nd)
        got = c * np.exp(2j * np.pi * (dw - delta) * S * mx) * sum(sp["taps_2"][j].astype(float) * smp[mx + j] for j in range(2 * H + 1))
        err = np.abs(got - exact).max() / np.abs(p1).max()
        print("split 1-D error relative to the input amplitude", err)
        assert err < 3e-6
    # a grid twice as wide (r_k = 0.1) cannot share one anchor; a 5-point axis is not worth it
    ks2 = synth.primary_ks(0.1, 7.0, 3)
    kw2, kstep2 = synth.sweep_params(ks2, 21)
    assert _taps.split_taps(256, mr, np.arange(ks2[0][0] - kw2, ks2[0][0] + kw2, kstep2)) is None
    assert _taps.split_taps(256, mr, wx[:5]) is None


@pytest.mark.parametrize("size,sigma,r_k,n_grid", [(2048, 10.0, 0.05, 41), (256, 10.0, 0.05, 21), (256, 10.0, 0.1, 21),
                                                    (160, 5.0, 0.1, 9), (320, 22.0, 0.0227, 9), (256, 10.0, 0.05, 5)])
def test_c_side_split_plan_agrees_with_python(size, sigma, r_k, n_grid):
    """gpa_split_plan (for non-Python hosts) reproduces pygpa_b200/_taps.split_taps."""
    import ctypes
    lib = _lib.load()
    mr = _taps.multirate_taps(size, size, sigma)
    assert mr is not None
    ks = synth.primary_ks(r_k, 7.0, 3)
    kw, kstep = synth.sweep_params(ks, n_grid)
    wx = np.ascontiguousarray(np.arange(ks[1][0] - kw, ks[1][0] + kw, kstep))
    sp = _taps.split_taps(size, mr, wx)
    r1, h, s1 = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_double(0)
    t1, t2 = np.zeros(446, dtype=np.float32), np.zeros(23, dtype=np.float32)
    rc = lib.gpa_split_plan(size, mr["S"], mr["sigma_a"], _lib.as_pd(wx), wx.size, ctypes.byref(r1), ctypes.byref(h),
                            ctypes.byref(s1), _lib.as_pf(t1), _lib.as_pf(t2))
    assert rc == (1 if sp else 0)
    if sp:
        assert (r1.value, h.value) == (sp["R1"], sp["H"]) and abs(s1.value - sp["sigma_1"]) < 1e-9
        assert np.abs(t1[:2 * r1.value + 1] - sp["taps_1"]).max() < 1e-7
        assert np.abs(t2[:2 * h.value + 1] - sp["taps_2"]).max() < 1e-6


def test_multirate_plans_fit_the_pass2_tile_in_shared_memory():
    """Large sigma: the stride-8 plan's pass-2 tile (S (W2 kP + J + 3) fine rows) would exceed 227 KB of shared
    memory from sigma ~ 29.5 on; the planner must then pick a smaller stride instead of a plan that fails at launch."""
    for sigma in (5.0, 10.0, 17.0, 22.0, 28.0, 29.5, 30.0, 33.0, 40.0, 48.0):
        mr = _taps.multirate_taps(2048, 2048, sigma)
        assert mr is not None
        s = mr["S"]
        j = -(-(2 * mr["Ra_x"] + 1) // s)
        w2 = 4 if s == 8 else 8
        assert s * (w2 * 16 + j + 3) * 36 * 8 <= 227 * 1024, (sigma, s, j)
        assert s * (8 * 16 + j + 3) * (33 * 4 + 16) + 4 <= 227 * 1024          # pass-1 tile
    assert _taps.multirate_taps(2048, 2048, 28.0)["S"] == 8 and _taps.multirate_taps(2048, 2048, 30.0)["S"] == 4


def test_split_plans_are_finite_and_bounded():
    """Wide grids at large sigma used to overflow the planner's gain exp(2 pi^2 dw^2 ...) into NaN, which compared as
    'within tolerance'.  Every accepted plan must have a finite error below SPLIT_TOL and a re-amplification <= 8
    (fp32 rounding noise of the anchor stage comes back multiplied by it)."""
    rng = np.random.default_rng(4)
    seen = 0
    for _ in range(60):
        n = int(rng.choice([96, 128, 256, 512, 1024, 2048]))
        sigma = float(rng.choice([4.5, 5, 9, 10, 17.9, 22, 29, 35, 49]))
        mr = _taps.multirate_taps(n, n, sigma)
        if mr is None:
            continue
        ks = synth.primary_ks(float(rng.choice([0.5 / sigma, 0.05, 0.1, 0.02])), 7.0, 3)
        kw, kstep = synth.sweep_params(ks, int(rng.choice([9, 21, 41])))
        sp = _taps.split_taps(n, mr, np.arange(ks[0][0] - kw, ks[0][0] + kw, kstep))
        if sp is not None:
            seen += 1
            assert np.isfinite(sp["err"]) and sp["err"] <= _taps.SPLIT_TOL and sp["c_max"] <= 8.0 * 1.001
    assert seen >= 10


@pytest.mark.parametrize("shape,sigma,r_k,n_grid,peak,expect", [
    ((2048, 2048), 10, 0.05, 41, 0, (4, 43, 43, 7, 6.8557)),        # BASELINE config 3 (bench.py)
    ((1024, 1024), 10, 0.05, 21, 0, (4, 43, 43, 7, 6.8557)),        # config 2
    ((160, 128), 5, 0.1, 9, 1, (2, 21, 21, 7, 3.4278)),
    ((256, 320), 22, 0.5 / 22, 9, 1, (8, 95, 103, 7, 16.4924)),
])
def test_plans_validated_on_hardware_stay_put(shape, sigma, r_k, n_grid, peak, expect):
    """The multirate / split plans of the configurations whose GPU parity runs are recorded in profiles/ (round 1):
    a planner change that alters them needs a new run of the -m gpu tests, so it must show up here first."""
    mr = _taps.multirate_taps(shape[0], shape[1], float(sigma))
    ks = synth.primary_ks(r_k, 7.0, 3)
    kw, kstep = synth.sweep_params(ks, n_grid)
    k = ks[peak]
    sp = _taps.split_taps(shape[0], mr, np.arange(k[0] - kw, k[0] + kw, kstep))
    assert sp is not None
    assert (mr["S"], mr["Ra_x"], sp["R1"], sp["H"]) == expect[:4] and abs(sp["sigma_1"] - expect[4]) < 1e-4
