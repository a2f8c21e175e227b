"""k-grid sharded sweep over 2 GPUs must be bit-identical to the single-GPU sweep, with both transports
(peer memory kernels over NVLink; torch.distributed collectives).  The multi-GPU cases skip on boxes
with fewer than 2 GPUs; the single-rank cases of the same code path always run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pygpa_b200 import synth

pytestmark = pytest.mark.gpu


def _equal(a, b, rows=None):
    r0, r1 = rows if rows is not None else (0, a["key"].shape[0])
    ok = {}
    ok["key"] = torch.equal(a["key"], b["key"])            # merged keys are complete on every rank
    ok["kidx"] = torch.equal(a["kidx"][r0:r1], b["kidx"][r0:r1])
    ok["lockin"] = torch.equal(torch.view_as_real(a["lockin"])[r0:r1], torch.view_as_real(b["lockin"])[r0:r1])
    ok["grad"] = torch.equal(a["grad"][r0:r1], b["grad"][r0:r1])
    if a.get("w") is not None and b.get("w") is not None:
        ok["w"] = torch.equal(a["w"][:, r0:r1], b["w"][:, r0:r1])
    return ok


def _worker(rank, world, port, img, ks, kw, kstep, sigma, out_dir):
    from pygpa_b200 import dist as gdist
    from pygpa_b200 import engine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = engine.require_cuda()
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    report = []
    try:
        d_img = engine.image_to_device(img, dev)
        plans = []
        for i, k in enumerate(ks):
            wxs, wys = engine.grid_axes(k[0], k[1], kw, kstep)
            if i == 1:
                wys = wys[:-1]          # unequal plane counts per peak (np.arange lengths are rounding dependent)
            plans.append(engine.SweepPlan(d_img.shape, wxs, wys, sigma, device=dev, private_ws=True))
        single = [p.run(d_img, k, want_w=True) for p, k in zip(plans, ks)]
        single = [{k_: (v.clone() if v is not None else None) for k_, v in o.items()} for o in single]
        for transport, dst in (("peer", 0), ("peer", "rows"), ("peer2", 0), ("collective", None), ("collective", 0)):
            # "peer2": the two-phase variant of the threshold exchange (gpa_sweep_arm_two_phase)
            sw = gdist.ShardedSweep(plans, ks, dst=dst, want_w=True, transport=transport.rstrip("2"), timeout_s=5.0,
                                    two_phase=transport == "peer2")
            for rep in range(3):        # epochs: the flags are never reset
                outs = sw(d_img, join=rep != 1)     # a frame stream may skip the join between frames
            torch.cuda.synchronize()
            sw.check()
            rows = outs[0]["rows"]
            if rows[1] > rows[0]:
                for p, (a, b) in enumerate(zip(outs, single)):
                    res = _equal(a, b, rows)
                    report.append((transport, str(dst), p, rank, rows, res))
            sw.close()
        open(os.path.join(out_dir, f"report{rank}"), "w").write(repr(report))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_sweep_bit_identical(tmp_path):
    cfg = synth.make_config('C2', size=192, n_grid=9)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, cfg["image"], cfg["ks"], cfg["kw"], cfg["kstep"], cfg["sigma"], str(tmp_path)),
             nprocs=2, join=True)
    seen = set()
    for rank in range(2):
        for transport, dst, p, r, rows, res in eval(open(tmp_path / f"report{rank}").read()):   # noqa: S307 - our own file
            assert all(res.values()), f"{transport} dst={dst} peak {p} rank {r} rows {rows}: {res}"
            seen.add((transport, dst, r))
    # rank 0 checked every configuration, rank 1 its row slice
    assert {("peer", "0", 0), ("peer", "rows", 0), ("peer", "rows", 1), ("peer2", "0", 0), ("collective", "None", 0),
            ("collective", "None", 1), ("collective", "0", 0)} <= seen


def test_single_rank_peer_path_matches_plan_run():
    """World size 1 takes the same code path as a rank of a sharded job (arena, owner-writes finalize with the CTA
    compaction, k-index / w decoded from the keys) minus the exchange: must equal SweepPlan.run bit for bit."""
    from pygpa_b200 import dist as gdist
    from pygpa_b200 import engine
    cfg = synth.make_config('C2', size=160, n_grid=7)
    dev = engine.require_cuda()
    d_img = engine.image_to_device(cfg["image"], dev)
    plans = []
    for k in cfg["ks"]:
        wxs, wys = engine.grid_axes(k[0], k[1], cfg["kw"], cfg["kstep"])
        plans.append(engine.SweepPlan(d_img.shape, wxs, wys, cfg["sigma"], device=dev, private_ws=True))
    for out_f64 in (False, True):
        single = [p.run(d_img, k, out_f64=out_f64, want_w=True) for p, k in zip(plans, cfg["ks"])]
        single = [{k_: (v.clone() if v is not None else None) for k_, v in o.items()} for o in single]
        sw = gdist.ShardedSweep(plans, cfg["ks"], dst=0, out_f64=out_f64, want_w=True, transport="peer")
        outs = sw(d_img)
        torch.cuda.synchronize()
        sw.check()
        for a, b in zip(outs, single):
            res = _equal(a, b)
            assert all(res.values()), res
        # grad_mode none: no gradient written, lock-in unchanged
        outs = sw(d_img, grad_mode=engine.GRAD_NONE)
        torch.cuda.synchronize()
        for a, b in zip(outs, single):
            assert torch.equal(torch.view_as_real(a["lockin"]), torch.view_as_real(b["lockin"]))
        sw.close()
