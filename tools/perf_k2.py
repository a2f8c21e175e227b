"""K2: pipelined DCT kernels vs the one-CTA-per-pair kernels: bit-identity and timings (library event profiler).

usage: python tools/perf_k2.py [n ...]      (default 2048 1024)
"""
import sys, ctypes, json, numpy as np, torch
sys.path.insert(0, '.')
from pygpa_b200 import engine, solvers, _lib
dev = engine.require_cuda(); lib = _lib.load()
sizes = [int(a) for a in sys.argv[1:]] or [2048, 1024]
out = {}


def timed(fn, names, reps=3):
    fn(); torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(0), ctypes.c_int(0)
    lib.gpa_profile_read(b"none", ctypes.byref(tot), ctypes.byref(cnt), 1)      # reset clears every record
    lib.gpa_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); lib.gpa_profile_enable(0)
    res = {"wall_ms": e0.elapsed_time(e1) / reps}
    for nme in names:
        lib.gpa_profile_read(nme.encode(), ctypes.byref(tot), ctypes.byref(cnt), 0)
        res[nme] = tot.value / reps
    lib.gpa_profile_read(b"none", ctypes.byref(tot), ctypes.byref(cnt), 1)
    return res


for shape in [(n, n) for n in sizes] + [(512, 2048), (1023, 1024), (300, 256)]:
    n, m = shape
    rng = np.random.default_rng(0)
    x, y = np.meshgrid(np.arange(n), np.arange(m), indexing='ij')
    psi = torch.from_numpy(((0.01 * x + 0.02 * y + rng.normal(size=(n, m)) * 0.1 + np.pi) % (2 * np.pi)) - np.pi).to(dev)
    w = torch.from_numpy(rng.uniform(0.1, 1, size=(n, m))).to(dev)
    rec = {}
    res = {}
    for mode in (0, 1):
        lib.gpa_set_dct_pipeline(mode)
        f = solvers.dctn(psi); b = solvers.dctn(f, inverse=True)
        phi, it = solvers.unwrap(psi=psi, weight=w, kmax=10, return_iters=True)
        res[mode] = (f.clone(), b.clone(), phi.clone(), it)
        if shape[0] == shape[1] and n >= 1024:
            rec["pipe" if mode else "old"] = timed(lambda: solvers.unwrap(psi=psi, weight=w, kmax=10),
                                                   ("uw_setup", "uw_poisson_solve", "uw_vector_ops"))
    rec["dctn_equal"] = bool(torch.equal(res[0][0], res[1][0])); rec["idctn_equal"] = bool(torch.equal(res[0][1], res[1][1]))
    rec["unwrap_equal"] = bool(torch.equal(res[0][2], res[1][2])); rec["iters"] = [res[0][3], res[1][3]]
    rec["dctn_maxdiff"] = float((res[0][0] - res[1][0]).abs().max()); rec["unwrap_maxdiff"] = float((res[0][2] - res[1][2]).abs().max())
    rec["roundtrip_err"] = float((res[1][1] - psi).abs().max())
    out[f"{n}x{m}"] = rec
    print(f"{n}x{m}", json.dumps(rec), flush=True)
lib.gpa_set_dct_pipeline(1)
json.dump(out, open("gpurun_out/perf_k2.json", "w"), indent=1)
