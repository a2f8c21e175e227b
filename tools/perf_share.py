"""One GPU playing rank r of an N-rank k-grid sharded C3 sweep: per-kernel CUDA-event times of the rank's share
(interleaved planes r, r+N, ...) on ONE stream, and of the owner-writes finalize of that share with local
destinations — isolates what does not scale from the NVLink part.  python tools/perf_share.py [N ...]"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygpa_b200 import _lib, engine, synth      # noqa: E402
from pygpa_b200 import dist as gdist            # noqa: E402

dev = engine.require_cuda()
lib = _lib.load()
cfg = synth.make_config("C3")
img = engine.image_to_device(cfg["image"], dev)
ks = cfg["ks"]
plans = []
for k in ks:
    wxs, wys = engine.grid_axes(k[0], k[1], cfg["kw"], cfg["kstep"])
    plans.append(engine.SweepPlan(img.shape, wxs, wys, cfg["sigma"], device=dev, private_ws=True))
names = ("k_mr_pass1a", "k_mr_pass1b", "k_mr_pass2a", "k_mr_pass2b", "k_mr_order", "k_mr_interp", "k_mr_finalize")


def read():
    tot, n = ctypes.c_double(0), ctypes.c_int(0)
    out = {}
    for nm in names:
        lib.gpa_profile_read(nm.encode(), ctypes.byref(tot), ctypes.byref(n), 0)
        out[nm] = round(tot.value, 4)
    lib.gpa_profile_read(b"k_mr_pass1", ctypes.byref(tot), ctypes.byref(n), 1)
    return out


full_keys = []
for p, plan in enumerate(plans):
    key = torch.zeros((plan.n, plan.m), dtype=torch.int64, device=dev)
    plan.argmax(img, key)
    full_keys.append(key)
torch.cuda.synchronize()
for world in [int(a) for a in (sys.argv[1:] or ["1", "2", "4", "8"])]:
    for rank in sorted({0, world // 2, world - 1}):
        ranges = gdist.shard_units_interleaved(3, 41, world, rank)
        lock = torch.empty((plans[0].n, plans[0].m), dtype=torch.complex64, device=dev)
        grad = torch.empty((plans[0].n, plans[0].m, 2), dtype=torch.float32, device=dev)
        res = {}
        for rep in range(2):
            if rep == 1:
                lib.gpa_profile_enable(1)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            keys = []
            for p, plan in enumerate(plans):
                lo, hi, st = ranges[p]
                # PERFECT=1: start from the final single-GPU keys = the best thresholds any exchange could ever provide
                key = full_keys[p].clone() if os.environ.get("PERFECT") == "1" else torch.zeros((plan.n, plan.m), dtype=torch.int64, device=dev)
                if hi > lo:
                    plan.argmax(img, key, lo, hi, st)
                keys.append(key)
            e[1].record()
            for p, plan in enumerate(plans):
                lo, hi, st = ranges[p]
                mr = plan.mr
                ws = plan._workspace()
                ld = (ctypes.c_void_p * 1)(lock.data_ptr())
                gd = (ctypes.c_void_p * 1)(grad.data_ptr())
                _lib.check(lib.gpa_sweep_finalize_mr_sharded(
                    *plan._geom(), lo, hi, st, mr["S"], mr["Ra_x"], mr["Ra_y"], _lib.as_pf(mr["taps_bx"]), _lib.as_pf(mr["taps_by"]),
                    mr["Rb"], *plan._split_geom(), engine._ptr(full_keys[p]), ks[p][0], ks[p][1], 0, 0, ld, gd, 1, plan.n, 1,
                    engine._ptr(ws), ws.numel(), engine._stream()))
            e[2].record()
            torch.cuda.synchronize()
        lib.gpa_profile_enable(0)
        res = read()
        res.update(world=world, rank=rank, units=sum(-(-(hi - lo) // st) for lo, hi, st in ranges),
                   argmax_ms=round(e[0].elapsed_time(e[1]), 4), finalize_ms=round(e[1].elapsed_time(e[2]), 4))
        print(json.dumps(res), flush=True)
