"""One C3 sweep (real synthetic frame, 1 peak after a 1-peak warm-up) for ncu captures."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from pygpa_b200 import synth, engine
dev = engine.require_cuda()
cfg = synth.make_config("C3")
img = engine.image_to_device(cfg["image"], dev)
k = cfg["ks"][0]
wxs, wys = engine.grid_axes(k[0], k[1], cfg["kw"], cfg["kstep"])
plan = engine.SweepPlan(img.shape, wxs, wys, cfg["sigma"], device=dev)
plan.run(img, k)
torch.cuda.synchronize()
plan.run(img, k)
torch.cuda.synchronize()
