"""Per-phase timing of the k-grid sharded sweep on the C3 frame (torchrun, N ranks):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/perf_dist.py [transport ...]
Prints, per transport, the step time (CUDA events, max over ranks) and the per-peak phase table of every rank."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygpa_b200 import dist as gdist      # noqa: E402
from pygpa_b200 import engine, synth      # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = engine.require_cuda()
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = synth.make_config("C3")
img = engine.image_to_device(cfg["image"], dev)
ks = cfg["ks"]
plans = []
for k in ks:
    wxs, wys = engine.grid_axes(k[0], k[1], cfg["kw"], cfg["kstep"])
    plans.append(engine.SweepPlan(img.shape, wxs, wys, cfg["sigma"], device=dev, private_ws=True))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


single = None
if rank == 0:
    single = [{k_: (v.clone() if v is not None else None) for k_, v in p.run(img, k).items()} for p, k in zip(plans, ks)]
for spec in (sys.argv[1:] or ["peer:0", "peer:rows", "collective:0"]):
    transport, dst = spec.split(":")
    dst = int(dst) if dst.isdigit() else (None if dst == "none" else dst)
    sw = gdist.ShardedSweep(plans, ks, dst=dst, transport=transport)
    for _ in range(3):
        outs = sw(img)
    barrier()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    e0.record()
    t0 = time.perf_counter()
    for _ in range(n):
        outs = sw(img, join=False)
    sw.join()
    e1.record()
    enqueue_ms = (time.perf_counter() - t0) / n * 1e3
    barrier()
    sw.check()
    ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    sw.record = True
    outs = sw(img)
    tm = sw.timings() if transport == "peer" else []
    sw.record = False
    ok = None
    if rank == 0 and outs[0]["rows"][1] > 0:
        r0, r1 = outs[0]["rows"]
        ok = all(torch.equal(a["key"], b["key"]) and torch.equal(torch.view_as_real(a["lockin"])[r0:r1], torch.view_as_real(b["lockin"])[r0:r1])
                 and torch.equal(a["grad"][r0:r1], b["grad"][r0:r1]) for a, b in zip(outs, single))
    rows = [None] * world
    if world > 1:
        dist.all_gather_object(rows, tm)
    else:
        rows = [tm]
    if rank == 0:
        print(json.dumps({"transport": transport, "dst": dst, "world": world, "ms_per_step": float(ms.item()), "host_enqueue_ms_per_step": enqueue_ms,
                          "bit_identical_to_single_gpu": ok, "phases_per_rank": rows}), flush=True)
    sw.close()
if world > 1:
    dist.destroy_process_group()
