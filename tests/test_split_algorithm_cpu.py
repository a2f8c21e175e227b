"""Algorithm-level parity of the multirate sweep WITH the split pass 2, on the CPU.

A float64 NumPy emulation of what the CUDA kernels compute (k_mr_pass1 -> anchor stage ->
k_mr_pass2b -> k_mr_interp, csrc/lockin.cu), driven by the REAL host planners
(pygpa_b200/_taps.multirate_taps, split_taps), is compared with the oracle's arg-max: the
selected candidate may differ only at near-ties (tests/parity.py).  This pins the mathematics of
the factorisations (G = G_a * G_b, G_a = G_1 * G_2 with the anchor / frequency-shift identity and
the wrapped-row correction) independently of the GPU; the kernels themselves are checked against
the oracle by the -m gpu tests.
"""
import numpy as np
import pytest

import oracle
from parity import NEAR_TIE
from pygpa_b200 import _taps, synth


def _dec_filter(a, taps, r, s, axis):
    """out[m] = sum_d taps[d + r] a[(s m + d) mod n] along axis."""
    n = a.shape[axis]
    base = np.arange(0, n, s)
    out = 0
    for d in range(-r, r + 1):
        out = out + taps[d + r] * np.take(a, (base + d) % n, axis=axis)
    return out


def _interp(c, taps, r, s, axis, n):
    """out[x] = s sum_m taps[x - s m + r] c[m mod n/s], |x - s m| <= r."""
    nc = c.shape[axis]
    x = np.arange(n)
    out = 0
    for j in range(-(r // s) - 1, r // s + 2):
        m = x // s + j
        d = x - s * m
        wgt = np.where(np.abs(d) <= r, taps[np.clip(d + r, 0, 2 * r)], 0.0) * s
        shape = [1] * c.ndim
        shape[axis] = n
        out = out + wgt.reshape(shape) * np.take(c, m % nc, axis=axis)
    return out


def _split_pass1(img, wys, mr, sp):
    """All planes P1[iy] (N, Md) from ONE anchor plane: the split applied along axis 1 (k_mr_pass1<ANCHOR> +
    k_mr_pass1b).  The rows are independent, so this is _split_pass2 on the transposed problem with a real input."""
    return [p.T for p in _split_pass2(img.T.astype(complex), wys, mr, sp)]


def _split_pass2(p1, wxs, mr, sp):
    """All candidates' coarse grids P2[ix] (Nd, Md) from one anchor stage."""
    n = p1.shape[0]
    S, H, R1 = mr["S"], sp["H"], sp["R1"]
    sa2, s12 = mr["sigma_a"] ** 2, sp["sigma_1"] ** 2
    s22 = sa2 - s12
    nd, nde, rtot = n // S, n // S + 2 * H, R1 + S * H
    wx0 = wxs[len(wxs) // 2]
    xu = np.arange(S * nde + 2 * R1 + 1) - rtot            # unwrapped frame row of padded row r
    t = xu % n
    a = p1[t] * np.exp(2j * np.pi * wx0 * t)[:, None]
    body = ((xu >= 0) & (xu < n))[:, None]
    e = np.arange(nde)
    g1 = sp["taps_1"].astype(np.float64)
    A_body = sum(g1[i] * np.where(body[S * e + i], a[S * e + i], 0) for i in range(2 * R1 + 1))
    A_edge = sum(g1[i] * np.where(body[S * e + i], 0, a[S * e + i]) for i in range(2 * R1 + 1))
    # the kernel only adds A_edge within H + ceil(R1/S) + 1 rows of either end: it must vanish elsewhere
    eb = -(-R1 // S) + 1
    assert not np.any(A_edge[H + eb:nd + H - eb])
    h2 = sp["taps_2"].astype(np.float64)
    mx = np.arange(nd)
    out = []
    for w in wxs:
        dw = w - wx0
        delta = dw * sa2 / s22
        c = np.exp(2 * np.pi ** 2 * dw ** 2 * sa2 * s12 / s22)
        J = np.where(e < H + nd // 2, np.exp(2j * np.pi * dw * n), np.exp(-2j * np.pi * dw * n))
        smp = np.exp(2j * np.pi * delta * S * (e - H))[:, None] * (A_body + J[:, None] * A_edge)
        acc = sum(h2[j] * smp[mx + j] for j in range(2 * H + 1))
        out.append(c * np.exp(2j * np.pi * (dw - delta) * S * mx)[:, None] * acc)
    return out


@pytest.mark.parametrize("shape,sigma,r_k,n_grid", [((96, 128), 10, 0.05, 9), ((64, 48), 5, 0.1, 9)])
def test_multirate_split_argmax_matches_oracle(shape, sigma, r_k, n_grid):
    n, m = shape
    ks = synth.primary_ks(r_k, 7.0, 3)
    u = synth.smooth_random_field(shape, 0.2, seed=21)
    img = synth.lattice_image(shape, ks, u, noise=0.3, seed=22)
    img -= img.mean()
    kw, kstep = synth.sweep_params(ks, n_grid)
    k = ks[0]
    ref = oracle.wfr_sweep(img, sigma, k[0], k[1], kw, kstep, return_diag=True, want_grad=False)
    wxs, wys = ref["wxs"], ref["wys"]
    mr = _taps.multirate_taps(n, m, float(sigma))
    assert mr is not None
    sp = _taps.split_taps(n, mr, wxs)
    spy = _taps.split_taps(m, mr, wys)
    assert sp is not None and spy is not None, "this configuration is meant to exercise both split passes"
    S, ra, rb = mr["S"], mr["Ra_x"], mr["Rb"]
    best = np.zeros(shape)
    bidx = np.full(shape, -1)
    y = np.arange(m)
    worst_p2 = 0.0
    p1s = _split_pass1(img, wys, mr, spy)
    worst_p1 = 0.0
    for iy, wy in enumerate(wys):
        p1 = p1s[iy]                                                                                             # (n, m/S)
        if iy in (0, len(wys) // 2, len(wys) - 1):
            single = _dec_filter(img * np.exp(2j * np.pi * wy * y)[None, :], mr["taps_ay"].astype(np.float64), ra, S, 1)
            worst_p1 = max(worst_p1, np.abs(p1 - single).max() / np.abs(single).max())
        p2s = _split_pass2(p1, wxs, mr, sp)
        for ix, wx in enumerate(wxs):
            if iy == len(wys) // 2 and ix in (0, len(wxs) - 1):     # against the single-stage pass 2, border rows included
                single = _dec_filter(p1 * np.exp(2j * np.pi * wx * np.arange(n))[:, None], mr["taps_ax"].astype(np.float64), ra, S, 0)
                worst_p2 = max(worst_p2, np.abs(p2s[ix] - single).max() / np.abs(single).max())
            sf = _interp(_interp(p2s[ix], mr["taps_bx"].astype(np.float64), rb, S, 0, n), mr["taps_by"].astype(np.float64), rb, S, 1, m)
            a2 = sf.real ** 2 + sf.imag ** 2
            idx = ix * len(wys) + iy
            take = (a2 > best) | ((a2 == best) & (idx < bidx) & (a2 > 0))
            best[take] = a2[take]
            bidx[take] = idx
    assert worst_p1 < 1e-5, f"split pass 1 differs from the single-stage pass 1 by {worst_p1:.2e}"
    assert worst_p2 < 1e-5, f"split pass 2 differs from the single-stage pass 2 by {worst_p2:.2e}"
    same = bidx == ref["kidx"]
    gap = (ref["amp1"] - ref["amp2"]) / np.maximum(ref["amp1"], 1e-300)
    assert np.all(gap[~same] < NEAR_TIE), f"{(~same & ~(gap < NEAR_TIE)).sum()} pixels differ away from a near-tie"
    assert (~same).mean() < 2e-3
    amp_err = np.abs(np.sqrt(best) - ref["amp1"]).max() / ref["amp1"].max()
    assert amp_err < 2e-5, f"winning amplitude differs by {amp_err:.2e}"
