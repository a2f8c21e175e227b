import sys, numpy as np, torch, scipy.ndimage as ndi
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import oracle
from pygpa_b200 import synth, cuGPA
from pygpa_b200 import geometric_phase_analysis as GPA
s = 256; shape = (s, s)
ks = synth.primary_ks(0.1, 7.0, 3)
bump = 0.5 * synth.gaussian_bump(shape)
deformed = synth.lattice_image(shape, ks, bump, second_order=0.3)
noise = ndi.gaussian_filter(2 * np.random.default_rng(0).normal(size=shape), sigma=0.5)
img = deformed + noise
u_ref, gs_ref = oracle.extract_displacement_field(img, ks, return_gs=True, sweep=lambda im, s_, kx, ky, kw, kstep: oracle.wfr_sweep(im, s_, kx, ky, kw, kstep, want_grad=False, return_diag=True))
u, gs = GPA.extract_displacement_field(img, ks, return_gs=True)
err = np.abs(u - u_ref).max(axis=0)
print('max err', err.max(), 'n>1e-3', (err > 1e-3).sum(), 'n>1e-4', (err>1e-4).sum(), 'interior max', err[20:-20,20:-20].max())
flip = np.zeros(shape, bool)
for g, r in zip(gs, gs_ref):
    same = np.all(g['w'] == r['w'], axis=0)
    gap = (r['amp1']-r['amp2'])/r['amp1']
    print(' mismatches', (~same).sum(), 'max gap', gap[~same].max() if (~same).any() else 0, 'max dphase at mismatch', np.abs(np.angle(g['lockin']*np.conj(r['lockin'])))[~same].max() if (~same).any() else 0)
    flip |= ~same
dist = ndi.distance_transform_edt(~flip)
for thr in (1e-3, 1e-4):
    m = err > thr
    print(thr, 'pixels', m.sum(), 'max distance to a flip', dist[m].max() if m.any() else None)
# feed oracle phases into GPU tail to isolate
phases = np.stack([np.angle(r['lockin']) for r in gs_ref]); 
mask = np.zeros(shape); sig = int(np.ceil(1/np.linalg.norm(ks,axis=1).min())); mask[2*sig:-2*sig, 2*sig:-2*sig] = 1
weights = np.stack([np.abs(r['lockin']) for r in gs_ref])*(mask+1e-6)
u_tail = GPA.reconstruct_u_inv_from_phases(ks, phases, weights)
print('tail only err', np.abs(u_tail - u_ref).max())
