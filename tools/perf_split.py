"""Split vs single-stage pass 2 on the C3 frame: ms per step and per-kernel event times."""
import ctypes, json, sys, numpy as np, torch
sys.path.insert(0, '.')
from pygpa_b200 import synth, engine, _lib
dev = engine.require_cuda()
lib = _lib.load()
size, ng = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 41)
cfg = synth.make_config('C3', size=size, n_grid=ng)
img = engine.image_to_device(cfg['image'], dev)
ks = cfg['ks']
names = ("k_mr_pass1", "k_mr_pass1a", "k_mr_pass1b", "k_mr_pass2", "k_mr_pass2a", "k_mr_pass2b", "k_mr_order", "k_mr_interp", "k_mr_finalize")
res = {}
keys = {}
for method in ("multirate", "multirate-single", "multirate"):
    plans = []
    for k in ks:
        wxs, wys = engine.grid_axes(k[0], k[1], cfg['kw'], cfg['kstep'])
        plans.append(engine.SweepPlan(img.shape, wxs, wys, cfg['sigma'], device=dev, method=method))
    outs = None
    def step():
        global outs
        outs = [p.run(img, k) for p, k in zip(plans, ks)]
    for _ in range(3): step()
    torch.cuda.synchronize()
    keys[method] = [o["kidx"].clone() for o in outs]
    ts = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(7):
        e0.record(); step(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    lib.gpa_profile_enable(1)
    for _ in range(3): step()
    torch.cuda.synchronize()
    lib.gpa_profile_enable(0)
    tot, n = ctypes.c_double(0), ctypes.c_int(0)
    kern = {}
    for nm in names:
        lib.gpa_profile_read(nm.encode(), ctypes.byref(tot), ctypes.byref(n), 0)
        if n.value: kern[nm] = round(tot.value / 3, 3)
    lib.gpa_profile_read(b"k_mr_pass1", ctypes.byref(tot), ctypes.byref(n), 1)
    res[method] = dict(ms_per_step=float(np.median(ts)), split=plans[0].split is not None and {k_: v for k_, v in plans[0].split.items() if not k_.startswith('taps')},
                       kernels_ms_per_step=kern)
    print(method, json.dumps(res[method]), flush=True)
diff = sum(int((a != b).sum().item()) for a, b in zip(keys["multirate"], keys["multirate-single"]))
print("k-index differences split vs single:", diff, "of", 3 * size * size)
json.dump(res, open("gpurun_out/perf_split.json", "w"), indent=1)
