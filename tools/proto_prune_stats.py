"""How tight could the spatial branch-and-bound of k_mr_interp get?  CPU emulation on a C3-like frame.

For every candidate: |sf|^2 on the fine grid (exact Gaussian, FFT) and |P2|^2 on the coarse grid
(G_a-filtered, decimated by S).  With the FINAL winners as thresholds (best case for any processing order)
count the (block, candidate) pairs that survive the bound  max_{window} |P2|^2 * 1.0002 >= min_{block} winner
for several block sizes / windows:
  8x8 cells, 3x3 neighbouring blocks (what the kernel does now)
  8x8 cells, exact halo of +-6 cells
  4x4 cells, exact halo
  2x2 cells, exact halo
and how many candidates actually win a pixel in a block (the floor for any bound).
"""
import sys, time
import numpy as np
from scipy.ndimage import maximum_filter
sys.path.insert(0, '.')
from pygpa_b200 import synth
from pygpa_b200._taps import multirate_taps

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ng = int(sys.argv[2]) if len(sys.argv) > 2 else 41
cfg = synth.make_config('C3', size=size, n_grid=ng)
img = cfg['image']; n = size
mr = multirate_taps(n, n, cfg['sigma']); S = mr['S']; sa = mr['sigma_a']
k = cfg['ks'][0]
wxs = np.arange(k[0] - cfg['kw'], k[0] + cfg['kw'], cfg['kstep'])
wys = np.arange(k[1] - cfg['kw'], k[1] + cfg['kw'], cfg['kstep'])
f = np.fft.fftfreq(n)
T = np.exp(-2 * np.pi ** 2 * cfg['sigma'] ** 2 * (f[:, None] ** 2 + f[None, :] ** 2))
Ta = np.exp(-2 * np.pi ** 2 * sa ** 2 * (f[:, None] ** 2 + f[None, :] ** 2))
x = np.arange(n)[:, None]; y = np.arange(n)[None, :]
nd = n // S
best = np.zeros((n, n)); widx = np.full((n, n), -1)
schemes = {'8x8 cells, 3x3 blocks (now)': (8, None), '8x8 cells, halo 6': (8, 6), '4x4 cells, halo 6': (4, 6), '2x2 cells, halo 6': (2, 6)}
bounds = {name: [] for name in schemes}
t0 = time.time()
for ix, wx in enumerate(wxs):
    for iy, wy in enumerate(wys):
        F = np.fft.fft2(img * np.exp(2j * np.pi * (wx * x + wy * y)))
        a2 = np.abs(np.fft.ifft2(F * T)) ** 2
        p2 = np.abs(np.fft.ifft2(F * Ta)[::S, ::S]) ** 2
        idx = ix * len(wys) + iy
        take = a2 > best
        best[take] = a2[take]; widx[take] = idx
        for name, (b, halo) in schemes.items():
            if halo is None:
                bm = p2.reshape(nd // b, b, nd // b, b).max(axis=(1, 3))
                bounds[name].append(maximum_filter(bm, size=3, mode='wrap'))
            else:
                # window = block + halo cells each side, evaluated at the block's first cell
                w = b + 2 * halo
                mf = maximum_filter(p2, size=w, mode='wrap', origin=0)
                # maximum_filter is centred: centre of the window for block starting at c0 is c0 + (b-1)/2
                c = np.arange(0, nd, b) + (b - 1) // 2
                bounds[name].append(mf[np.ix_(c, c)] if (w % 2 == 1 or True) else None)
print(f"{len(wxs) * len(wys)} candidates at {n}^2 in {time.time() - t0:.0f} s")
ncand = len(wxs) * len(wys)
for name, (b, halo) in schemes.items():
    px = b * S
    thr = best.reshape(n // px, px, n // px, px).min(axis=(1, 3))
    B = np.stack(bounds[name])                 # (cand, blocks, blocks)
    alive = B * 1.0002 >= thr[None]
    # candidates that really win somewhere in the block
    wb = widx.reshape(n // px, px, n // px, px).transpose(0, 2, 1, 3).reshape(n // px, n // px, -1)
    nwin = np.array([[len(np.unique(wb[i, j])) for j in range(n // px)] for i in range(n // px)])
    print(f"{name:30s}: block {px:3d} px: survivors {alive.mean() * 100:5.1f} % of (block, candidate) pairs "
          f"= {alive.sum(0).mean():6.1f} per block; true winners per block {nwin.mean():5.1f}")
    if b == 8 and halo is None:
        # tile = 2 x 4 blocks (64 x 128 px): survivors per tile = union over its blocks
        a = alive.reshape(ncand, n // px // 2, 2, n // px // 4, 4).any(axis=(2, 4))
        print(f"{'':30s}  per 64x128 tile (any block): {a.mean() * 100:5.1f} %")
