"""One weighted PCG unwrap (kmax=2) per DCT mode for an ncu launch list: old kernels first, then the pipelined ones."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from pygpa_b200 import engine, solvers, _lib
dev = engine.require_cuda(); lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rng = np.random.default_rng(0)
x, y = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
psi = torch.from_numpy(((0.01*x + 0.02*y + rng.normal(size=(n, n))*0.1 + np.pi) % (2*np.pi)) - np.pi).to(dev)
w = torch.from_numpy(rng.uniform(0.1, 1, size=(n, n))).to(dev)
for mode in (0, 1):
    lib.gpa_set_dct_pipeline(mode)
    solvers.unwrap(psi=psi, weight=w, kmax=3)
    torch.cuda.synchronize()
