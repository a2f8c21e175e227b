"""NumPy prototype of the multirate (decimate/interpolate) Gaussian lock-in: accuracy vs the oracle."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import oracle
from pygpa_b200 import synth

def taps(n, sigma, r):
    f = np.fft.fftfreq(n)
    kern = np.fft.ifft(np.exp(-2*np.pi**2*sigma**2*f**2)).real
    d = np.arange(-r, r+1)
    return kern[d % n]

def circ_dec_filter(a, g, r, s, axis, dtype):
    """out[m] = sum_d g[d+r] a[(s m + d) mod n] along axis (decimating)."""
    n = a.shape[axis]
    out = 0
    base = np.arange(0, n, s)
    for d in range(-r, r+1):
        out = out + dtype(g[d+r]) * np.take(a, (base + d) % n, axis=axis)
    return out.astype(a.dtype)

def circ_interp_filter(c, g, r, s, axis, n, dtype):
    """out[x] = s * sum_m g[(x - s m) + r] c[m mod n/s]  for |x - s m| <= r."""
    nc = c.shape[axis]
    x = np.arange(n)
    out = 0
    jmax = r // s + 1
    m0 = x // s
    for j in range(-jmax, jmax+1):
        m = m0 + j
        d = x - s*m
        ok = np.abs(d) <= r
        w = np.where(ok, g[np.clip(d + r, 0, 2*r)], 0.0) * s
        shape = [1]*c.ndim; shape[axis] = n
        out = out + w.astype(dtype).reshape(shape) * np.take(c, m % nc, axis=axis)
    return out.astype(c.dtype)

def sweep_multirate(img, sigma, wxs, wys, s, sigma_b, trunc, cdtype=np.complex128):
    rdtype = np.float64 if cdtype == np.complex128 else np.float32
    n, m = img.shape
    sigma_a = np.sqrt(sigma**2 - sigma_b**2)
    ra, rb = int(np.ceil(trunc*sigma_a)), int(np.ceil(trunc*sigma_b))
    gax, gay = taps(n, sigma_a, ra), taps(m, sigma_a, ra)
    gbx, gby = taps(n, sigma_b, rb), taps(m, sigma_b, rb)
    x = np.arange(n); y = np.arange(m)
    best = np.zeros((n, m), rdtype); bidx = np.full((n, m), -1)
    ny = len(wys)
    for iy, wy in enumerate(wys):
        ph_y = np.exp(2j*np.pi*((wy*y) % 1.0)).astype(cdtype)
        p1 = circ_dec_filter(img.astype(rdtype)[:, :] * ph_y[None, :], gay, ra, s, 1, rdtype)     # (n, m/s)
        for ix, wx in enumerate(wxs):
            ph_x = np.exp(2j*np.pi*((wx*x) % 1.0)).astype(cdtype)
            p2 = circ_dec_filter(p1 * ph_x[:, None], gax, ra, s, 0, rdtype)                        # (n/s, m/s)
            p3 = circ_interp_filter(p2, gbx, rb, s, 0, n, rdtype)                                   # (n, m/s)
            sf = circ_interp_filter(p3, gby, rb, s, 1, m, rdtype)                                   # (n, m)
            a2 = (sf.real**2 + sf.imag**2).astype(rdtype)
            idx = ix*ny + iy
            t = (a2 > best) | ((a2 == best) & (idx < bidx) & (a2 > 0))
            best[t] = a2[t]; bidx[t] = idx
    return bidx, best, (ra, rb)

if __name__ == '__main__':
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    ng = int(sys.argv[2]) if len(sys.argv) > 2 else 11
    cfg = synth.make_config('C2', size=size, n_grid=ng)
    k = cfg['ks'][0]
    ref = oracle.wfr_sweep(cfg['image'], cfg['sigma'], k[0], k[1], cfg['kw'], cfg['kstep'], return_diag=True, want_grad=False)
    gap = (ref['amp1'] - ref['amp2'])/ref['amp1']
    wxs, wys = ref['wxs'], ref['wys']
    for (s, sb, tr, dt) in ((4, 5.0, 4.5, np.complex128), (4, 5.0, 5.0, np.complex128), (4, 4.5, 4.5, np.complex128), (4, 5.0, 4.5, np.complex64), (4, 5.0, 5.0, np.complex64), (1, 5.0, 4.5, np.complex64)):
        t = time.time()
        bidx, best, rr = sweep_multirate(cfg['image'], cfg['sigma'], wxs, wys, s, sb, tr, dt)
        same = bidx == ref['kidx']
        amp_err = np.abs(np.sqrt(best.astype(np.float64)) - ref['amp1'])/ref['amp1'].max()
        print(f"s={s} sigma_b={sb} trunc={tr} {dt.__name__}: R=({rr[0]},{rr[1]}) mismatches {np.sum(~same)}/{same.size} max gap at mismatch {gap[~same].max() if (~same).any() else 0:.3g}  max amp err {amp_err.max():.3g}  ({time.time()-t:.1f}s)")
