"""NumPy prototype of the SPLIT decimating pass 2 (shared anchor stage + per-candidate coarse stage).

Single stage (k_mr_pass2):  P2[wx](mx) = sum_x' G_a(S mx - x') e^{2 pi i wx t(x')} P1(t(x')),  t = x' mod N
Split:   G_a = G_1 * G_2 (sigma_a^2 = sigma_1^2 + sigma_2^2)
  stage A (per plane, anchor wx0):  A_body / A_edge (e) = sum_x' G_1(S (e - H) - x') [mask] P1(t) e^{2 pi i wx0 t}
  stage B (per candidate, coarse rate):
      P2(mx) = c e^{2 pi i (dw - delta) S mx} sum_j h[j] e^{2 pi i delta S (mx + j - H)} (A_body + J A_edge)(mx + j)
      delta = dw sigma_a^2 / sigma_2^2,  c = exp(2 pi^2 dw^2 sigma_a^2 sigma_1^2 / sigma_2^2),  h[j] = S G_2(S (j - H))
      J = e^{+2 pi i dw N} for the low halo, e^{-2 pi i dw N} for the high halo
Checks P2 against the single-stage result in float64 and float32, including the frame border.
"""
import sys
import numpy as np


def gauss_taps(n, sigma, r):
    f = np.fft.fftfreq(n)
    kern = np.fft.ifft(np.exp(-2 * np.pi ** 2 * sigma ** 2 * f ** 2)).real
    return kern[np.arange(-r, r + 1) % n]


def single_stage(p1, wx, ga, ra, S, cdtype):
    n = p1.shape[0]
    x = np.arange(n)
    b = (p1 * np.exp(2j * np.pi * ((wx * x) % 1.0))[:, None]).astype(cdtype)
    base = np.arange(0, n, S)
    out = np.zeros((n // S, p1.shape[1]), cdtype)
    rd = np.float32 if cdtype == np.complex64 else np.float64
    for d in range(-ra, ra + 1):
        out += rd(ga[d + ra]) * b[(base + d) % n]
    return out


def split_stage(p1, wxs, wx0, sigma_a, sigma_1, S, trunc1, trunc2, cdtype):
    n, md = p1.shape
    nd = n // S
    rd = np.float32 if cdtype == np.complex64 else np.float64
    sigma_2 = np.sqrt(sigma_a ** 2 - sigma_1 ** 2)
    r1 = int(np.ceil(trunc1 * sigma_1))
    H = int(np.ceil(trunc2 * sigma_2 / S))
    g1 = gauss_taps(n, sigma_1, r1)
    u = S * (np.arange(2 * H + 1) - H)
    h = S * np.exp(-u ** 2 / (2 * sigma_2 ** 2)) / (sigma_2 * np.sqrt(2 * np.pi))
    rtot = r1 + S * H
    nde = nd + 2 * H
    # padded linear domain: row r <-> unwrapped frame row r - rtot
    r = np.arange(S * nde + 2 * r1 + 1)
    xu = r - rtot
    t = xu % n
    a = (p1[t] * np.exp(2j * np.pi * ((wx0 * t) % 1.0))[:, None]).astype(cdtype)
    body = ((xu >= 0) & (xu < n))[:, None]
    A_body = np.zeros((nde, md), cdtype)
    A_edge = np.zeros((nde, md), cdtype)
    e = np.arange(nde)
    for i in range(2 * r1 + 1):           # out(e) = sum_i g1[i] sample(S e + i)
        A_body += rd(g1[i]) * np.where(body[S * e + i], a[S * e + i], 0)
        A_edge += rd(g1[i]) * np.where(body[S * e + i], 0, a[S * e + i])
    outs = []
    for wx in wxs:
        dw = wx - wx0
        delta = dw * sigma_a ** 2 / sigma_2 ** 2
        c = np.exp(2 * np.pi ** 2 * dw ** 2 * sigma_a ** 2 * sigma_1 ** 2 / sigma_2 ** 2)
        car = np.exp(2j * np.pi * ((delta * S * (e - H)) % 1.0)).astype(cdtype)
        J = np.where(e - H < nd // 2, np.exp(2j * np.pi * ((dw * n) % 1.0)), np.exp(-2j * np.pi * ((dw * n) % 1.0))).astype(cdtype)
        smp = (car[:, None] * (A_body + J[:, None] * A_edge)).astype(cdtype)
        mx = np.arange(nd)
        acc = np.zeros((nd, md), cdtype)
        for j in range(2 * H + 1):
            acc += rd(h[j]) * smp[mx + j]
        derot = (c * np.exp(2j * np.pi * (((dw - delta) * S * mx) % 1.0))).astype(cdtype)
        outs.append((derot[:, None] * acc).astype(cdtype))
    return outs, dict(r1=r1, H=H, sigma_2=sigma_2, cmax=np.exp(2 * np.pi ** 2 * np.max(np.abs(np.asarray(wxs) - wx0)) ** 2 * sigma_a ** 2 * sigma_1 ** 2 / sigma_2 ** 2))


if __name__ == '__main__':
    sys.path.insert(0, '.')
    from pygpa_b200 import synth
    from pygpa_b200._taps import multirate_taps
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ng = int(sys.argv[2]) if len(sys.argv) > 2 else 21
    cfg = synth.make_config('C3', size=size, n_grid=ng)
    img = cfg['image']
    k = cfg['ks'][0]
    n, m = img.shape
    mr = multirate_taps(n, m, cfg['sigma'])
    S, sigma_a, ra = mr['S'], mr['sigma_a'], mr['Ra_x']
    wxs = np.arange(k[0] - cfg['kw'], k[0] + cfg['kw'], cfg['kstep'])
    wys = np.arange(k[1] - cfg['kw'], k[1] + cfg['kw'], cfg['kstep'])
    wx0 = 0.5 * (wxs[0] + wxs[-1])
    # P1 for the central plane (float64, exact circular)
    y = np.arange(m)
    wy = wys[len(wys) // 2]
    by = img * np.exp(2j * np.pi * ((wy * y) % 1.0))[None, :]
    gay = gauss_taps(m, sigma_a, ra)
    p1 = np.zeros((n, m // S), complex)
    base = np.arange(0, m, S)
    for d in range(-ra, ra + 1):
        p1 += gay[d + ra] * by[:, (base + d) % m]
    gax = gauss_taps(n, sigma_a, ra)
    sel = [0, len(wxs) // 4, len(wxs) // 2, len(wxs) - 1]
    ref64 = [single_stage(p1, wxs[i], gax, ra, S, np.complex128) for i in sel]
    ref32 = [single_stage(p1.astype(np.complex64), wxs[i], gax, ra, S, np.complex64) for i in sel]
    scale = max(np.abs(r).max() for r in ref64)
    print(f"N={n} S={S} sigma_a={sigma_a:.3f} Ra={ra} kw={cfg['kw']:.4f}; single-stage f32 vs f64: "
          f"{max(np.abs(a - b).max() for a, b in zip(ref32, ref64)) / scale:.2e}")
    for sigma_1 in (5.0, 6.0, 6.5, 7.0):
        for tr1, tr2 in ((4.5, 4.5), (4.5, 5.0), (5.0, 5.0)):
            o64, info = split_stage(p1, [wxs[i] for i in sel], wx0, sigma_a, sigma_1, S, tr1, tr2, np.complex128)
            o32, _ = split_stage(p1.astype(np.complex64), [wxs[i] for i in sel], wx0, sigma_a, sigma_1, S, tr1, tr2, np.complex64)
            e64 = [np.abs(a - b).max() / scale for a, b in zip(o64, ref64)]
            e64_edge = [max(np.abs(a - b)[:12].max(), np.abs(a - b)[-12:].max()) / scale for a, b in zip(o64, ref64)]
            e32 = [np.abs(a - b).max() / scale for a, b in zip(o32, ref64)]
            print(f"sigma_1={sigma_1} trunc=({tr1},{tr2}) R1={info['r1']} H={info['H']} ({2 * info['H'] + 1} taps) c_max={info['cmax']:.2f}: "
                  f"f64 err {max(e64):.2e} (border rows {max(e64_edge):.2e})  f32 err {max(e32):.2e}")
    # errors against the UNTRUNCATED circular Gaussian (what the reference applies)
    rfull = (n - 1) // 2
    gfull = gauss_taps(n, sigma_a, rfull)
    exact = [single_stage(p1, wxs[i], gfull, rfull, S, np.complex128) for i in sel]
    print("vs untruncated: single-stage(4.5 sigma) err", f"{max(np.abs(a - b).max() for a, b in zip(ref64, exact)) / scale:.2e}")
    for sigma_1, tr1, tr2 in ((6.5, 4.5, 4.5), (6.5, 4.5, 5.0), (6.5, 5.0, 5.0), (6.0, 4.5, 4.5), (5.0, 4.5, 4.5)):
        o32, info = split_stage(p1.astype(np.complex64), [wxs[i] for i in sel], wx0, sigma_a, sigma_1, S, tr1, tr2, np.complex64)
        print(f"  split sigma_1={sigma_1} trunc=({tr1},{tr2}) taps {2 * info['H'] + 1}: f32 err vs untruncated "
              f"{max(np.abs(a - b).max() for a, b in zip(o32, exact)) / scale:.2e}")
