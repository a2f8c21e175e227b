"""K5 streaming kernels at 2048^2: phasegradient2J (104 B/pixel) and props_from_Jac (64 B/pixel), back-to-back timing."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from pygpa_b200 import engine, solvers, synth
dev = engine.require_cuda()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rng = np.random.default_rng(0)
ks = synth.primary_ks(0.05, 7.0, 3)
g = torch.from_numpy(rng.uniform(-0.3, 0.3, size=(3, n, n, 2))).to(dev)
w = torch.from_numpy(rng.uniform(0.1, 1, size=(3, n, n))).to(dev)
K = 2 * np.pi * ks
for wrap in (False, True):
    for _ in range(3): J = solvers.phasegradient_to_J(g, w, K, do_wrap=wrap)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): J = solvers.phasegradient_to_J(g, w, K, do_wrap=wrap)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"k_grad2J wrap={wrap}: {ms*1e3:.1f} us, {104*n*n/ms/1e6:.0f} GB/s on the 104 B/pixel basis; checksum {float(J.sum()):.12e}")
for _ in range(3): P = solvers.props_from_jac(J, add_identity=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): P = solvers.props_from_jac(J, add_identity=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"k_props_from_jac: {ms*1e3:.1f} us, {64*n*n/ms/1e6:.0f} GB/s on the 64 B/pixel basis; checksum {float(torch.nan_to_num(P).sum()):.12e}")
