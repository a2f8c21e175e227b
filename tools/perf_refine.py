import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from pygpa_b200 import engine, solvers, synth
from pygpa_b200 import geometric_phase_analysis as GPA
dev = engine.require_cuda()
n = 2048
xx, yy = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
for name, img in (("bump", 8*np.exp(-((xx-900)**2+(yy-1200)**2)/2e5)+0.01*xx), ("waves", 30*np.sin(xx/300.0)*np.cos(yy/400.0)+0.2*yy)):
    d = solvers.to_device_f64(img, dev)
    solvers.fit_plane_huber(d)
    torch.cuda.synchronize(); t = time.perf_counter()
    th, it = solvers.fit_plane_huber(d, return_iters=True)
    torch.cuda.synchronize(); print(name, th, it, f"{(time.perf_counter()-t)*1e3:.2f} ms")
cfg = synth.make_config("C2")
t = time.perf_counter(); prs, w, corr = GPA.iterate_GPA(cfg["image"], cfg["ks"] * 1.01, cfg["sigma"]); print("iterate_GPA 1024^2 first call", time.perf_counter()-t)
t = time.perf_counter(); prs, w, corr = GPA.iterate_GPA(cfg["image"], cfg["ks"] * 1.01, cfg["sigma"]); print("iterate_GPA 1024^2", time.perf_counter()-t, corr)
