import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygpa_b200 import engine, synth
cfg = synth.make_config('C2', size=512, n_grid=5)
dev = engine.require_cuda()
img = engine.image_to_device(cfg["image"], dev)
k = cfg["ks"][1]
wxs, wys = engine.grid_axes(k[0], k[1], cfg["kw"], cfg["kstep"])
plan = engine.SweepPlan(img.shape, wxs, wys, cfg["sigma"], device=dev, method="multirate")
a = plan.run(img, k)
torch.cuda.synchronize()
print("ok", int(a["key"].sum().item()))
