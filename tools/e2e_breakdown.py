import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from pygpa_b200 import synth, engine, cuGPA
cfg = synth.make_config('C3')
img_host = torch.from_numpy(cfg['image']).pin_memory()
ks = cfg['ks']
def once():
    t = {}
    t0 = time.perf_counter()
    dev = engine.require_cuda()
    k = ks[0]
    img = engine.image_to_device(img_host.numpy(), dev); torch.cuda.synchronize(); t['h2d+cast'] = time.perf_counter() - t0; t0 = time.perf_counter()
    wxs, wys = engine.grid_axes(k[0], k[1], cfg['kw'], cfg['kstep'])
    plan = engine.SweepPlan(img.shape, wxs, wys, cfg['sigma'], engine.CAND_GRID, device=dev); t['plan'] = time.perf_counter() - t0; t0 = time.perf_counter()
    res = plan.run(img, k, engine.GRAD_CENTRAL, out_f64=True, want_w=True, want_kidx=False); torch.cuda.synchronize(); t['run'] = time.perf_counter() - t0; t0 = time.perf_counter()
    host = {kk: cuGPA._to_host(res[kk]) for kk in ('lockin', 'w', 'grad')}; t['to_host_enqueue'] = time.perf_counter() - t0; t0 = time.perf_counter()
    torch.cuda.current_stream().synchronize(); t['d2h_wait'] = time.perf_counter() - t0; t0 = time.perf_counter()
    out = {kk: v.numpy() for kk, v in host.items()}; t['numpy'] = time.perf_counter() - t0
    return t, out
keep = []
for i in range(6):
    t, out = once(); keep.append(out); keep = keep[-1:]
    print(i, {k: round(v * 1e3, 2) for k, v in t.items()}, 'total', round(sum(t.values()) * 1e3, 1))
print('--- bench-like loop: 3 peaks per step, outs held')
def e2e_step():
    return [cuGPA.wfr2_grad_opt(img_host.numpy(), cfg['sigma'], k[0], k[1], cfg['kw'], cfg['kstep']) for k in ks]
for i in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter(); outs = e2e_step(); torch.cuda.synchronize(); print(i, round((time.perf_counter() - t0) * 1e3, 1), 'ms/step')
