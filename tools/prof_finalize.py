"""ncu target: one k_mr_finalize (whole frame, N = 1 layout) and one k_mr_finalize_sharded launch (rank 0 of 8: planes 0, 8, ..)
on the C3 frame, peak 0, after a warm-up of both."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygpa_b200 import _lib, engine, synth      # noqa: E402

dev = engine.require_cuda()
lib = _lib.load()
cfg = synth.make_config("C3")
img = engine.image_to_device(cfg["image"], dev)
k = cfg["ks"][0]
wxs, wys = engine.grid_axes(k[0], k[1], cfg["kw"], cfg["kstep"])
plan = engine.SweepPlan(img.shape, wxs, wys, cfg["sigma"], device=dev, private_ws=True)
key = torch.zeros((plan.n, plan.m), dtype=torch.int64, device=dev)
plan.argmax(img, key)
lock = torch.empty((plan.n, plan.m), dtype=torch.complex64, device=dev)
grad = torch.empty((plan.n, plan.m, 2), dtype=torch.float32, device=dev)
mr = plan.mr
ws = plan._workspace()
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for rep in range(2):
    plan.finalize(img, key, k, planes_valid=True)
    ld = (ctypes.c_void_p * 1)(lock.data_ptr())
    gd = (ctypes.c_void_p * 1)(grad.data_ptr())
    _lib.check(lib.gpa_sweep_finalize_mr_sharded(
        *plan._geom(), 0, plan.wy.size, world, mr["S"], mr["Ra_x"], mr["Ra_y"], _lib.as_pf(mr["taps_bx"]), _lib.as_pf(mr["taps_by"]),
        mr["Rb"], *plan._split_geom(), engine._ptr(key), k[0], k[1], 0, 0, ld, gd, 1, plan.n, 1,
        engine._ptr(ws), ws.numel(), engine._stream()))
    torch.cuda.synchronize()
