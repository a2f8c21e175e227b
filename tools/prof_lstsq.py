"""The two per-pixel least-squares solves of reconstruct_u_inv_from_phases at 2048^2 (for ncu / timing)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from pygpa_b200 import engine, solvers, synth
dev = engine.require_cuda()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rng = np.random.default_rng(0)
ks = synth.primary_ks(0.05, 7.0, 3)
ph = torch.from_numpy(rng.uniform(-np.pi, np.pi, size=(3, n, n))).to(dev)
w = torch.from_numpy(rng.uniform(0.1, 1, size=(3, n, n))).to(dev)
for _ in range(3):
    a = solvers.lstsq(ph, solvers.SRC_DIFF1, ks, w); b = solvers.lstsq(ph, solvers.SRC_DIFF0, ks, w)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    a = solvers.lstsq(ph, solvers.SRC_DIFF1, ks, w); b = solvers.lstsq(ph, solvers.SRC_DIFF0, ks, w)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 40
print(f"k_lstsq {ms*1e3:.1f} us per solve, {64*n*n/ms/1e6:.0f} GB/s on the 64 B/pixel basis; checksum {float(a.sum()+b.sum()):.12e}")
