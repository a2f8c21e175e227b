import os, sys, time, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
from pygpa_b200 import synth, engine
from pygpa_b200 import dist as gdist
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK'])); dev = engine.require_cuda()
dist.init_process_group('nccl', device_id=dev)
ks = synth.primary_ks(0.05, 7.0, 3); kw, kstep = synth.sweep_params(ks, 41)
img = torch.from_numpy(np.random.default_rng(0).normal(size=(2048, 2048)).astype(np.float32)).to(dev)
plans = []
for k in ks:
    wxs, wys = engine.grid_axes(k[0], k[1], kw, kstep)
    plans.append(engine.SweepPlan(img.shape, wxs, wys, 10, device=dev, private_ws=True))
ranges = gdist.shard_units(3, 41, world, rank)
def timeit(fn, name, n=10):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    if rank == 0: print(f"{name:34s} gpu {e0.elapsed_time(e1)/n:7.2f} ms   wall {(t1-t0)/n*1e3:7.2f} ms", flush=True)
def argmax_only():
    ks_ = []
    for plan, (lo, hi) in zip(plans, ranges):
        key = torch.zeros((plan.n, plan.m), dtype=torch.int64, device=dev)
        if hi > lo: plan.argmax(img, key, lo, hi)
        ks_.append(key)
    return ks_
def argmax_and_keys():
    ks_ = argmax_only()
    w = [dist.all_reduce(k, op=dist.ReduceOp.MAX, async_op=True) for k in ks_]
    for x in w: x.wait()
    return ks_
def plus_finalize():
    ks_ = argmax_and_keys()
    outs = [plan.finalize(img, key, kref, 0, plane_begin=lo, plane_end=hi, want_kidx=False, planes_valid=True)
            for plan, key, kref, (lo, hi) in zip(plans, ks_, ks, ranges)]
    return ks_, outs
def full():
    return gdist.sharded_sweep(img, plans, ks, dst=0)
def full_allreduce():
    return gdist.sharded_sweep(img, plans, ks)
timeit(argmax_only, 'argmax only (local share)')
timeit(argmax_and_keys, '+ key MAX all-reduce')
timeit(plus_finalize, '+ finalize')
timeit(full, 'sharded_sweep (reduce to 0)')
timeit(full_allreduce, 'sharded_sweep (all-reduce payload)')

lock = [torch.zeros((2048, 2048), dtype=torch.complex64, device=dev) for _ in range(3)]
grad = [torch.zeros((2048, 2048, 2), dtype=torch.float32, device=dev) for _ in range(3)]
keys = [torch.zeros((2048, 2048), dtype=torch.int64, device=dev) for _ in range(3)]
def payload_only():
    w = []
    for a, b in zip(lock, grad):
        w.append(dist.all_reduce(torch.view_as_real(a), async_op=True)); w.append(dist.all_reduce(b, async_op=True))
    for x in w: x.wait()
def payload_one_call():
    flat = torch.cat([torch.view_as_real(a).reshape(-1) for a in lock] + [b.reshape(-1) for b in grad])
    dist.all_reduce(flat)
def keys_only():
    w = [dist.all_reduce(k, op=dist.ReduceOp.MAX, async_op=True) for k in keys]
    for x in w: x.wait()
def unpack_only():
    for k in keys: gdist.unpack_key(k)[1].to(torch.int32)
timeit(payload_only, '6 x all_reduce SUM f32 (33.5 MB each)')
timeit(payload_one_call, 'cat + 1 all_reduce (201 MB)')
timeit(keys_only, '3 x all_reduce MAX i64 (33.5 MB each)')
timeit(unpack_only, 'unpack_key x3')
dist.destroy_process_group()
