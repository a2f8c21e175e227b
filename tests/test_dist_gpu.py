"""k-grid sharded sweep over 2 GPUs (NCCL) must be bit-identical to the single-GPU sweep.
Skipped on boxes with fewer than 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pygpa_b200 import synth

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, img, ks, kw, kstep, sigma, out_dir):
    from pygpa_b200 import dist as gdist
    from pygpa_b200 import engine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = engine.require_cuda()
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        d_img = engine.image_to_device(img, dev)
        plans = []
        for k in ks:
            wxs, wys = engine.grid_axes(k[0], k[1], kw, kstep)
            plans.append(engine.SweepPlan(d_img.shape, wxs, wys, sigma, device=dev, private_ws=True))
        outs = gdist.sharded_sweep(d_img, plans, ks)
        if rank == 0:
            single = [p.run(d_img, k) for p, k in zip(plans, ks)]
            ok = all(torch.equal(a["key"], b["key"]) and torch.equal(a["kidx"], b["kidx"])
                     and torch.equal(torch.view_as_real(a["lockin"]), torch.view_as_real(b["lockin"]))
                     and torch.equal(a["grad"], b["grad"]) for a, b in zip(outs, single))
            open(os.path.join(out_dir, "ok"), "w").write("1" if ok else "0")
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
    assert open(tmp_path / "ok").read() == "1"
