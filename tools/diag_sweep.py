import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import oracle
from pygpa_b200 import synth, cuGPA, engine
np.set_printoptions(linewidth=200)

def compare(name, img, sigma, k, kw, kstep):
    o = oracle.wfr_sweep(img, sigma, k[0], k[1], kw, kstep, return_diag=True)
    t = time.time(); g = cuGPA.wfr2_grad_opt(img, sigma, k[0], k[1], kw, kstep); dt = time.time() - t
    nx, ny = len(o['wxs']), len(o['wys'])
    same = (g['w'][0] == o['w'][0]) & (g['w'][1] == o['w'][1])
    gap = (o['amp1'] - o['amp2']) / np.maximum(o['amp1'], 1e-300)
    print(f"{name}: shape {img.shape} grid {nx}x{ny} t={dt*1e3:.1f}ms  k mismatches {np.sum(~same)} / {same.size}; "
          f"max gap at mismatch {gap[~same].max() if (~same).any() else 0:.3g}")
    amax = np.abs(o['lockin']).max()
    d = np.abs(g['lockin'] - o['lockin'])[same]
    ph = np.abs(np.angle(g['lockin'] * np.conj(o['lockin'])))
    for thr in (0.0, 0.01, 0.05, 0.2):
        m = same & (np.abs(o['lockin']) > thr * amax)
        print(f"   amp>{thr:4.2f}max: max|dlockin|/amax {np.abs(g['lockin']-o['lockin'])[m].max()/amax:.3g}  max phase err {ph[m].max():.3g} rad")
    dg = np.abs(oracle.wrap_to_pi(2 * (g['grad'] - o['grad'])) / 2)
    for thr in (0.0, 0.05, 0.2):
        m = same & (np.abs(o['lockin']) > thr * amax)
        print(f"   grad err (mod pi) amp>{thr}: {dg[m].max():.3g}")
    return g, o

ks = synth.primary_ks(0.12, 7.0, 3)
gold = dict(np.load('tests/golden/sweep_64x48.npz'))
compare('golden64x48', gold['in_image'], int(gold['in_sigma']), gold['in_ks'][0], float(gold['in_kw']), float(gold['in_kstep']))
shape = (160, 128)
ks = synth.primary_ks(0.1, 7.0, 3)
u = synth.smooth_random_field(shape, 0.1, 5)
img = synth.lattice_image(shape, ks, u, noise=0.3, seed=6); img -= img.mean()
kw, kstep = synth.sweep_params(ks, 7)
compare('160x128 s5', img, 5, ks[1], kw, kstep)
cfg = synth.make_config('C2', size=256, n_grid=21)
compare('C2-256', cfg['image'], cfg['sigma'], cfg['ks'][0], cfg['kw'], cfg['kstep'])
compare('C2-256 k2', cfg['image'], cfg['sigma'], cfg['ks'][2], cfg['kw'], cfg['kstep'])
# fixed lock-in
r = cuGPA.cuGPA(cfg['image'], cfg['ks'][1], 10)
o = oracle.lockin_fixed(cfg['image'], cfg['ks'][1], 10)
print('fixed lockin max abs err / max', np.abs(r - o).max() / np.abs(o).max())
