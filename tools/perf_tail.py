import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from pygpa_b200 import engine, solvers, _lib
dev = engine.require_cuda(); lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rng = np.random.default_rng(0)
x, y = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
psi = torch.from_numpy(((0.01*x + 0.02*y + rng.normal(size=(n, n))*0.1 + np.pi) % (2*np.pi)) - np.pi).to(dev)
w = torch.from_numpy(rng.uniform(0.1, 1, size=(n, n))).to(dev)
u = torch.from_numpy(np.stack([3*np.sin(x/200.0), 2*np.cos(y/150.0)])).to(dev)
for _ in range(2): solvers.unwrap(psi=psi, weight=w, kmax=10); solvers.invert_u(u)
torch.cuda.synchronize(); lib.gpa_profile_enable(1)
solvers.unwrap(psi=psi, weight=w, kmax=10); solvers.invert_u(u); torch.cuda.synchronize(); lib.gpa_profile_enable(0)
tot, cnt = ctypes.c_double(0), ctypes.c_int(0)
for name in ("uw_setup", "uw_poisson_solve", "uw_vector_ops", "lf_prefilter", "k_invert_u"):
    lib.gpa_profile_read(name.encode(), ctypes.byref(tot), ctypes.byref(cnt), 0)
    print(f"{name:18s} {tot.value:8.3f} ms total ({cnt.value} timed regions)")
