"""Estimate how many candidates a tile-level bound |sf| <= max|P2| prunes on the C3 synthetic frame."""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'scratch')
from pygpa_b200 import synth
from proto_multirate import taps, circ_dec_filter, circ_interp_filter
cfg = synth.make_config('C3')
img_full = cfg['image']; k = cfg['ks'][0]
n = 384; x0, y0 = 700, 900
img = img_full[x0:x0+n, y0:y0+n].copy()
sigma, S = 10.0, 4
sb = 4.4; sa = np.sqrt(sigma**2 - sb**2); ra, rb = int(np.ceil(4.5*sa)), int(np.ceil(4.5*sb))
ga, gb = taps(n, sa, ra), taps(n, sb, rb)
kw, kstep = cfg['kw'], cfg['kstep']
wxs = np.arange(k[0]-kw, k[0]+kw, kstep); wys = np.arange(k[1]-kw, k[1]+kw, kstep)
xg = np.arange(x0, x0+n); yg = np.arange(y0, y0+n)
amp = np.zeros((len(wxs), len(wys), n, n), np.float32)      # |sf|
p2a = np.zeros((len(wxs), len(wys), n//S, n//S), np.float32)  # |P2|
for iy, wy in enumerate(wys):
    p1 = circ_dec_filter(img * np.exp(2j*np.pi*wy*yg)[None, :], ga, ra, S, 1, np.float64)
    for ix, wx in enumerate(wxs):
        p2 = circ_dec_filter(p1 * np.exp(2j*np.pi*wx*xg)[:, None], ga, ra, S, 0, np.float64)
        p3 = circ_interp_filter(p2, gb, rb, S, 0, n, np.float64)
        sf = circ_interp_filter(p3, gb, rb, S, 1, n, np.float64)
        amp[ix, iy] = np.abs(sf); p2a[ix, iy] = np.abs(p2)
best = amp.reshape(-1, n, n).max(axis=0)
print('amp range of best', best.min(), best.max())
TX, TY, H = 64, 128, 5
tot = surv_final = surv_center = 0
center = amp[:, len(wys)//2].max(axis=0)      # best after the centre plane only
for tx in range(1, n//TX - 1):
    for ty in range(1, n//TY - 1):
        sl = (slice(tx*TX, (tx+1)*TX), slice(ty*TY, (ty+1)*TY))
        cs = (slice(tx*TX//S - H, (tx+1)*TX//S + H), slice(ty*TY//S - H, (ty+1)*TY//S + H))
        m = p2a[:, :, cs[0], cs[1]].max(axis=(2, 3))
        surv_final += (m * 1.0001 >= best[sl].min()).sum()
        surv_center += (m * 1.0001 >= center[sl].min()).sum()
        tot += m.size
print(f'tiles {tot // m.size}: survivors with final thresholds {surv_final / tot:.3f}, with centre-plane thresholds {surv_center / tot:.3f}')
