"""Multi-GPU sharding of the adaptive sweep: one process per GPU, torch.distributed (NCCL).

The candidates are independent until the per-pixel arg-max, so the k-grid shards with ONE
exchange step: every rank holds the whole frame, runs pass 1 / pass 2 for its share of the
first-pass planes (sharding over wy duplicates no work), and the packed keys
(|sf|^2 bits << 32 | ~flat_index) are combined with an integer MAX all-reduce — largest
amplitude wins, exact ties go to the lowest flat index, which is the reference's strict-'>'
first-wins rule (geometric_phase_analysis.py:806).  Each rank then finalises the pixels whose
winner it owns; the payload is combined with a SUM all-reduce (exactly one non-zero
contributor per pixel, so the sum is exact and the N-GPU result is bit-identical to 1 GPU).

The helpers work on CPU tensors with the gloo backend too (tests/test_dist_gloo.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["shard_units", "merge_keys", "merge_payload", "pack_key", "unpack_key", "sharded_sweep"]


def shard_units(n_peaks, n_planes, world, rank):
    """Split the flattened (peak, plane) list into `world` contiguous, balanced shares.
    Returns [(plane_begin, plane_end)] * n_peaks for `rank` (empty ranges have begin == end)."""
    total = n_peaks * n_planes
    lo = (total * rank) // world
    hi = (total * (rank + 1)) // world
    out = []
    for p in range(n_peaks):
        a, b = max(lo, p * n_planes), min(hi, (p + 1) * n_planes)
        out.append((a - p * n_planes, b - p * n_planes) if b > a else (0, 0))
    return out


def pack_key(amp2, flat_index):
    """(float32 |sf|^2 >= 0, int index) -> int64 key, same encoding as k_pass2 (lockin.cu)."""
    bits = amp2.to(torch.float32).contiguous().view(torch.int32).to(torch.int64)
    return (bits << 32) | (0xFFFFFFFF - flat_index.to(torch.int64))


def unpack_key(key):
    """key -> (amp2 float32, flat index int64; -1 where no candidate won)."""
    amp2 = (key >> 32).to(torch.int32).view(torch.float32)
    idx = 0xFFFFFFFF - (key & 0xFFFFFFFF)
    return amp2, torch.where((key >> 32) == 0, torch.full_like(idx, -1), idx)


def merge_keys(key, group=None):
    """In-place MAX all-reduce of packed keys.  Keys are non-negative as int64 (the sign bit of
    |sf|^2 is clear), so the signed MAX torch offers equals the unsigned one the kernel uses."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(key, op=dist.ReduceOp.MAX, group=group)
    return key


def merge_payload(tensors, group=None):
    """SUM all-reduce of the finalised payload (each pixel is non-zero on exactly one rank)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in tensors:
            if t is not None:
                dist.all_reduce(torch.view_as_real(t) if t.is_complex() else t, op=dist.ReduceOp.SUM, group=group)
    return tensors


def sharded_sweep(img_dev, plans, krefs, grad_mode=0, group=None, dst=None):
    """All peaks of one frame, k-grid sharded over the ranks of `group`.
    plans: one engine.SweepPlan per peak (identical on every rank; give them private workspaces so
    the finalize reuses what the arg-max left and stays bit-identical to one GPU).
    Returns [dict(lockin, grad, kidx, key)] per peak — on every rank, or with dst=<rank> the payload
    (lockin, grad) is only reduced to that rank (half the NVLink traffic of an all-reduce).

    The collectives are asynchronous and pipelined per peak: the MAX all-reduce of peak p's keys
    runs on NCCL's stream while the arg-max kernels of peak p+1 execute, and the payload reduction
    of peak p overlaps the finalize of peak p+1."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return [plan.run(img_dev, kref, grad_mode) for plan, kref in zip(plans, krefs)]
    ranges = shard_units(len(plans), plans[0].wy.size, world, rank)
    keys, key_work = [], []
    for plan, (lo, hi) in zip(plans, ranges):
        key = torch.zeros((plan.n, plan.m), dtype=torch.int64, device=img_dev.device)
        if hi > lo:
            plan.argmax(img_dev, key, lo, hi)
        keys.append(key)
        key_work.append(dist.all_reduce(key, op=dist.ReduceOp.MAX, group=group, async_op=True))
    outs, pay_work = [], []
    for plan, key, work, kref, (lo, hi) in zip(plans, keys, key_work, krefs, ranges):
        work.wait()
        out = plan.finalize(img_dev, key, kref, grad_mode, plane_begin=lo, plane_end=hi, want_kidx=False,
                            planes_valid=plan._private)
        out["key"] = key
        outs.append(out)
        for t in (out["lockin"], out["grad"]):
            if t is None:
                continue
            flat = torch.view_as_real(t) if t.is_complex() else t
            if dst is None:
                pay_work.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True))
            else:
                pay_work.append(dist.reduce(flat, dst=dst, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for w in pay_work:
        w.wait()
    for o in outs:
        o["kidx"] = unpack_key(o["key"])[1].to(torch.int32)
    return outs
