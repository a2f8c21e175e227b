import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from pygpa_b200 import synth, engine, _lib
dev = engine.require_cuda(); lib = _lib.load()
size, ng = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 41)
method = sys.argv[3] if len(sys.argv) > 3 else "auto"
ks = synth.primary_ks(0.05, 7.0, 3); kw, kstep = synth.sweep_params(ks, ng)
if len(sys.argv) > 4 and sys.argv[4] == "noise":
    img = torch.from_numpy(np.random.default_rng(0).normal(size=(size, size)).astype(np.float32)).to(dev)
else:
    img = engine.image_to_device(synth.make_config('C3', size=size, n_grid=ng)['image'], dev)
k = ks[0]; wxs, wys = engine.grid_axes(k[0], k[1], kw, kstep)
plan = engine.SweepPlan(img.shape, wxs, wys, 10, device=dev, method=method)
for _ in range(3): plan.run(img, k)
torch.cuda.synchronize(); lib.gpa_profile_enable(1)
reps = 5
for _ in range(reps): plan.run(img, k)
torch.cuda.synchronize(); lib.gpa_profile_enable(0)
tot, n = ctypes.c_double(0), ctypes.c_int(0); total = 0
for name in ("k_mr_pass1", "k_mr_pass1a", "k_mr_pass1b", "k_mr_pass2", "k_mr_pass2a", "k_mr_pass2b", "k_mr_order", "k_mr_interp", "k_pass1",
             "k_pass2_argmax", "k_finalize", "k_mr_finalize"):
    lib.gpa_profile_read(name.encode(), ctypes.byref(tot), ctypes.byref(n), 0)
    if n.value: print(f"{name:16s} {tot.value / reps:8.3f} ms/peak ({n.value // reps} launches)"); total += tot.value / reps
print("sum", round(total, 3), "ms/peak; units/peak", size * size * ng * ng / 1e9, "G")
