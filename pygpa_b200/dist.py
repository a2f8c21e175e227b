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

__all__ = ["shard_units", "shard_units_interleaved", "merge_keys", "merge_payload", "pack_key", "unpack_key", "sharded_sweep"]


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


def shard_units_interleaved(n_peaks, n_planes, world, rank):
    """Round-robin share of the flattened (peak, plane) list: unit u = peak * n_planes + plane goes
    to rank u % world.  Returns [(plane_begin, plane_end, plane_step)] * n_peaks.  Every rank gets
    planes spread over the whole grid (in particular some near its centre, where the winners
    usually are), which keeps the exact pruning of the multirate arg-max effective on every rank."""
    out = []
    for p in range(n_peaks):
        begin = (rank - p * n_planes) % world
        out.append((begin, n_planes, world) if begin < n_planes else (0, 0, 1))
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
    the finalize reuses what the arg-max left and stays bit-identical to one GPU).  A single rank with
    private multirate plans takes the same route (no collectives): its peaks overlap on their streams.
    Returns [dict(lockin, grad, kidx, key)] per peak — on every rank, or with dst=<rank> the payload
    (lockin, grad) is only reduced to that rank.

    Exactly two collectives per frame, each issued when every rank has finished its local work:
    one MAX all-reduce of the stacked keys (8 B/pixel/peak) and one SUM reduction of the stacked
    payload (16 B/pixel/peak).  They are deliberately NOT overlapped with the arg-max kernels: the
    plane shares of the ranks start at different peaks, so an early collective would make NCCL's
    CTAs spin on the slower peer while occupying SMs the sweep kernels need (measured: +4 ms per
    frame on 2 GPUs, against 0.75 ms for the two collectives issued at the end)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_peaks, n, m, dev = len(plans), plans[0].n, plans[0].m, img_dev.device
    private = all(p._private and p.mr is not None and p.mr_in_flight == p.wy.size for p in plans)
    if world == 1 and not (private and dev.type == "cuda"):
        return [plan.run(img_dev, kref, grad_mode) for plan, kref in zip(plans, krefs)]
    # multirate plans with their own workspace take an interleaved share (good pruning thresholds on
    # every rank); otherwise contiguous plane ranges
    interleave = all(p.mr is not None and p._private and p.mr_in_flight >= -(-p.wy.size // world) for p in plans)
    if interleave:
        ranges = shard_units_interleaved(n_peaks, plans[0].wy.size, world, rank)
    else:
        ranges = [(lo, hi, 1) for lo, hi in shard_units(n_peaks, plans[0].wy.size, world, rank)]
    # The peaks are independent until the collectives.  With private workspaces their kernels go to one
    # stream per peak: a rank's share of one peak is only a few planes (5 of 41 on 8 GPUs = 320 pass-2
    # CTAs = 2.2 waves of the 148 SMs, which run as 3), so the tail of one peak is filled by the next.
    side = _peak_streams(dev, n_peaks) if dev.type == "cuda" and all(p._private for p in plans) else None
    keys = torch.zeros((n_peaks, n, m), dtype=torch.int64, device=dev)

    def per_peak(fn):
        if side is None:
            for p in range(n_peaks):
                fn(p)
            return
        main = torch.cuda.current_stream(dev)
        start = torch.cuda.Event()
        start.record(main)
        for p in range(n_peaks):
            side[p].wait_event(start)
            with torch.cuda.stream(side[p]):
                fn(p)
            done = torch.cuda.Event()
            done.record(side[p])
            main.wait_event(done)

    def argmax_peak(p):
        lo, hi, step = ranges[p]
        if hi > lo:
            plans[p].argmax(img_dev, keys[p], lo, hi, step)
    per_peak(argmax_peak)
    if world > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=group)
    want_grad = grad_mode != 2
    # payload buffer: [lockin (P,N,M,2) | grad (P,N,M,2)] float32, zero where this rank owns nothing
    payload = torch.zeros((2 if want_grad else 1, n_peaks, n, m, 2), dtype=torch.float32, device=dev)
    lockin = torch.view_as_complex(payload[0])
    outs = [{"lockin": lockin[p], "grad": payload[1, p] if want_grad else None, "w": None, "kidx": None}
            for p in range(n_peaks)]

    def finalize_peak(p):
        lo, hi, step = ranges[p]
        if hi > lo:
            plans[p].finalize(img_dev, keys[p], krefs[p], grad_mode, plane_begin=lo, plane_end=hi, want_kidx=False,
                              planes_valid=plans[p]._private, out=outs[p], plane_step=step)
    per_peak(finalize_peak)
    for p in range(n_peaks):
        outs[p]["key"] = keys[p]
    if world == 1:
        pass
    elif dst is None:
        dist.all_reduce(payload, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(payload, dst=dst, op=dist.ReduceOp.SUM, group=group)
    kidx = torch.empty((n_peaks, n, m), dtype=torch.int32, device=dev)
    if dev.type == "cuda":
        from . import _lib, engine
        _lib.check(_lib.load().gpa_key_to_kidx(engine._ptr(keys), engine._ptr(kidx), keys.numel(), engine._stream()))
    else:
        kidx.copy_(unpack_key(keys)[1])
    for p, o in enumerate(outs):
        o["kidx"] = kidx[p]
    return outs


_side_streams = {}


def _peak_streams(dev, n):
    """One side stream per peak and device, created once."""
    pool = _side_streams.setdefault(str(dev), [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:n]
