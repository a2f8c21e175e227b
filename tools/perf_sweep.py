import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from pygpa_b200 import synth, engine
dev = engine.require_cuda()
for name, size, ng in (('C2', 1024, 21), ('C3', 2048, 41)):
    ks = synth.primary_ks(0.05, 7.0, 3)
    kw, kstep = synth.sweep_params(ks, ng)
    rng = np.random.default_rng(0)
    img = torch.from_numpy(rng.normal(size=(size, size)).astype(np.float32)).to(dev)
    plans = []
    for k in ks:
        wxs, wys = engine.grid_axes(k[0], k[1], kw, kstep)
        assert len(wxs) == ng and len(wys) == ng
        plans.append(engine.SweepPlan(img.shape, wxs, wys, 10, device=dev))
    def step():
        for p, k in zip(plans, ks):
            p.run(img, k)
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        e0.record(); step(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = np.median(ts)
    work = size * size * 3 * ng * ng
    falg = 4 * 91 * (1 + 1 / ng) + 8 + 2 / ng
    print(f"{name}: {t:.2f} ms/step  {work / t / 1e3:.1f} Mpx.kvec/s  alg {work * falg / t / 1e9:.1f} TFLOP/s  in_flight={plans[0].in_flight}")
