import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import oracle
from pygpa_b200 import solvers, synth
from pygpa_b200 import geometric_phase_analysis as GPA
from pygpa_b200 import phase_unwrap as PU
g = dict(np.load('tests/golden/unwrap.npz'))
pr, wr = g["in_psi_r"], g["in_w_r"]
for kmax in (5, 20, 50, 100):
    ref, kr = oracle.phase_unwrap(pr, wr, kmax=kmax, return_iters=True)
    got = PU.phase_unwrap(pr, wr, kmax=kmax)
    print('golden r kmax', kmax, 'iters', kr, 'max diff', np.abs(got-ref).max(), 'range', np.ptp(ref))
rng = np.random.default_rng(64+128+100)
n, m = 64, 128
x, y = np.meshgrid(np.arange(n), np.arange(m), indexing='ij')
truth = 0.002 * (x - n / 3) ** 2 + 0.15 * y + 3 * np.sin(x / 9.0) * np.cos(y / 13.0)
psi = oracle.wrap_to_pi(truth + 0.1 * rng.normal(size=(n,m)))
w = rng.uniform(0.01, 1.0, size=(n,m)); w[5:9, 7:20] = 1e-6
for kmax in (10, 30, 60, 100):
    ref, kr = oracle.phase_unwrap(psi, w, kmax=kmax, return_iters=True)
    dev = solvers.require_cuda()
    got, kg = solvers.unwrap(psi=solvers.to_device_f64(psi, dev), weight=solvers.to_device_f64(w, dev), kmax=kmax, return_iters=True)
    print('64x128 kmax', kmax, 'iters', kr, kg, 'max diff', np.abs(got.cpu().numpy()-ref).max())
# lstsq
rng = np.random.default_rng(8)
d, n, m = 3, 37, 53
ks = synth.primary_ks(0.08, 11.0, d); K = 2*np.pi*ks
b = rng.normal(size=(d, n, m)); w = rng.uniform(1e-6, 1, size=(d, n+2, m+1))**3
ref = oracle.weighted_lstsq(b, K, w); got = GPA.myweighed_lstsq(b, K, w)
err = np.abs(got-ref).max(axis=0)
wl = w[:, :n, :m]
A = wl.reshape(d,-1).T[:,:,None]*K[None]
sv = np.linalg.svd(A, compute_uv=False); cond = (sv[:,0]/sv[:,1]).reshape(n,m)
i = np.unravel_index(err.argmax(), err.shape)
print('lstsq max err', err.max(), 'at cond', cond[i], 'rel', err.max()/np.abs(ref[:, i[0], i[1]]).max())
for c in (1e2, 1e4, 1e6, 1e8, 1e12):
    mk = cond < c
    print('  cond <', c, 'n', mk.sum(), 'max rel err', (err[mk]/np.abs(ref).max(axis=0)[mk]).max())
