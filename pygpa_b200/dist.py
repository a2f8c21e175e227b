"""Multi-GPU sharding of the adaptive sweep: one process per GPU of one NVSwitch box.

The candidates are independent until the per-pixel arg-max, so the k-grid shards with ONE exchange
step: every rank holds the whole frame, runs the arg-max kernels for its
share of the flattened (peak, first-pass plane) list — sharding over wy duplicates no work — and the
packed keys (|sf|^2 bits << 32 | ~flat_index) are combined with an integer MAX: largest amplitude
wins, exact ties go to the lowest flat index, which is the reference's strict-'>' first-wins rule
(geometric_phase_analysis.py:806).  Each rank then finalises the pixels whose winner it owns.

Two transports:

* ``peer`` (default on CUDA, world <= 8): the exchange runs inside our own kernels over NVLink peer
  memory (csrc/peer.cu, peer.py).  Per peak: local arg-max -> flag signal/wait -> ``k_key_merge``
  (in-place reduce-scatter + all-gather of the keys, every rank reduces 1/W of the pixels and stores
  the result into all ranks) -> signal/wait -> ``k_mr_finalize_sharded``, which writes the lock-in /
  gradient of the pixels a rank owns STRAIGHT INTO THE DESTINATION RANK'S output arrays (owner-writes:
  exactly one rank owns a pixel, so there is no payload reduction at all) -> signal to the
  destination.  The peaks run on priority-ordered streams, so the exchange and finalize of peak p
  overlap the arg-max of peak p+1; only the last peak's exchange is exposed.
* ``collective`` (fallback: direct-form plans, shared workspaces, CPU tensors with gloo in the
  tests): one MAX all-reduce of the stacked keys and one SUM reduction of the zero-padded payload
  through torch.distributed, as in round 1.

Both give results bit-identical to one GPU (tests/test_dist_gpu.py, bench.py's multi_gpu_check).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_units", "shard_units_interleaved", "merge_keys", "merge_payload", "pack_key", "unpack_key",
           "sharded_sweep", "ShardedSweep", "HostSweep"]


def _counts(n_peaks, n_planes):
    """Per-peak plane counts: an int applies to every peak, a sequence is taken as is.  np.arange grids are
    rounding dependent (6 or 7 planes for ksteps = 3), so the peaks of one lattice may differ."""
    if np.isscalar(n_planes):
        return [int(n_planes)] * int(n_peaks)
    counts = [int(c) for c in n_planes]
    if len(counts) != n_peaks:
        raise ValueError(f"{n_peaks} peaks but {len(counts)} plane counts")
    return counts


def shard_units(n_peaks, n_planes, world, rank):
    """Split the flattened (peak, plane) list into `world` contiguous, balanced shares.
    n_planes: planes per peak (int, or one count per peak).
    Returns [(plane_begin, plane_end)] * n_peaks for `rank` (empty ranges have begin == end)."""
    counts = _counts(n_peaks, n_planes)
    total = sum(counts)
    lo = (total * rank) // world
    hi = (total * (rank + 1)) // world
    out, off = [], 0
    for c in counts:
        a, b = max(lo, off), min(hi, off + c)
        out.append((a - off, b - off) if b > a else (0, 0))
        off += c
    return out


def shard_units_interleaved(n_peaks, n_planes, world, rank):
    """Round-robin share of the flattened (peak, plane) list: unit u = offset(peak) + plane goes to rank
    u % world (offset = planes of the earlier peaks).  Returns [(plane_begin, plane_end, plane_step)] *
    n_peaks.  Every rank gets planes spread over the whole grid (in particular some near its centre,
    where the winners usually are), which keeps the exact pruning of the multirate arg-max effective on
    every rank."""
    out, off = [], 0
    for c in _counts(n_peaks, n_planes):
        begin = (rank - off) % world
        out.append((begin, c, world) if begin < c else (0, 0, 1))
        off += c
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


# ----------------------------------------------------------------------------------------------
# the sharded sweep
# ----------------------------------------------------------------------------------------------
_PH_ARGMAX, _PH_MERGED, _PH_DELIVERED, _PH_BOUNDS, _PH_WAVE0 = 0, 1, 2, 3, 4
_FLAG_BYTES = 4096


class ShardedSweep:
    """All peaks of one frame, k-grid sharded over the ranks of `group`; reusable for every frame of that
    shape (collective constructor: every rank builds it with identical arguments).

    plans   one engine.SweepPlan per peak, identical on every rank.  The peer transport needs multirate
            plans with private workspaces that keep a rank's whole share resident (then the finalize
            interpolates from the coarse grids the arg-max left, exactly as on one GPU).
    dst     where the winner payload lands: a rank number (default 0: whole arrays on that rank),
            'rows' (row slice r of every array on rank r: N/W rows each, what the host-facing path uses —
            every GPU then copies its slice out over its own PCIe link), or None with the collective
            transport (every rank gets everything).
    A call returns, per peak, a dict of tensors that live in this object's buffers (valid until the next
    call): key, lockin, grad, kidx, w (if want_w), plus 'rows' = the (begin, end) rows that are valid on
    this rank ((0, 0) on a rank that is not a destination)."""

    def __init__(self, plans, krefs, group=None, dst=0, out_f64=False, want_w=False, transport="auto", timeout_s=20.0,
                 gossip=True, two_phase=False):
        from . import _lib
        self.plans, self.krefs, self.group = list(plans), [tuple(map(float, k)) for k in krefs], group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        p0 = self.plans[0]
        self.n, self.m, self.dev = p0.n, p0.m, p0.device
        if any((p.n, p.m) != (self.n, self.m) for p in self.plans):
            raise ValueError("all plans must share one frame shape")
        self.n_peaks = len(self.plans)
        self.out_f64, self.want_w = bool(out_f64), bool(want_w)
        self.timeout_s = float(timeout_s)
        counts = [p.wy.size for p in self.plans]
        world, rank = self.world, self.rank
        private_mr = all(p._private and p.mr is not None for p in self.plans)
        # multirate plans with their own workspace take an interleaved share (good pruning thresholds on
        # every rank) if it fits their resident planes; otherwise contiguous plane ranges
        inter = shard_units_interleaved(self.n_peaks, counts, world, rank)
        fits = all(-(-(hi - lo) // st) <= p.mr_in_flight for p, (lo, hi, st) in zip(self.plans, inter)) if private_mr else False
        self.ranges = inter if fits else [(lo, hi, 1) for lo, hi in shard_units(self.n_peaks, counts, world, rank)]
        can_peer = self.dev.type == "cuda" and private_mr and fits and world <= 8
        if transport == "auto":
            transport = "peer" if can_peer else "collective"
        if transport == "peer" and not can_peer:
            raise ValueError("the peer transport needs multirate plans with private workspaces holding a rank's share, "
                             "a CUDA device and at most 8 ranks")
        if transport not in ("peer", "collective"):
            raise ValueError("transport must be 'auto', 'peer' or 'collective'")
        self.transport = transport
        if dst == "rows" and transport != "peer":
            raise ValueError("dst='rows' needs the peer transport")
        if dst is None and transport == "peer":
            dst = "rows"
        self.dst = dst
        self.lib = _lib.load() if self.dev.type == "cuda" else None
        self.epoch = 0
        self._pending, self._join = [], True
        self.record = False          # True: bracket the phases of every peak with CUDA events (bench.py)
        self.after_peak = None       # optional callable(p), run on peak p's stream once its rows are delivered
        self._events = None
        self._streams = _peak_streams(self.dev, self.n_peaks) if self.dev.type == "cuda" and all(p._private for p in self.plans) else None
        n, m, P = self.n, self.m, self.n_peaks
        cb, rb = (16, 8) if self.out_f64 else (8, 4)
        if transport == "peer":
            from .peer import PeerArena
            if 5 * P * 64 > 2048:
                raise ValueError("too many peaks for the flag block")
            off_key = _FLAG_BYTES
            off_lock = off_key + P * n * m * 8
            off_grad = off_lock + P * n * m * cb
            off_hint = off_grad + P * n * m * 2 * rb
            # threshold gossip (gpa_sweep_arm_gossip): one uint64 per bound block (8 coarse cells = 8 S pixels) and peak
            blk = 8 * self.plans[0].mr["S"]
            self._hint_n = (n // blk) * (m // blk) if all(p.mr["S"] * 8 == blk for p in self.plans) and n % blk == 0 and m % blk == 0 else 0
            self.gossip = bool(gossip) and world > 1 and self._hint_n > 0
            off_best = off_hint + (-(-P * self._hint_n * 8 // 256) * 256 if self.gossip else 0)
            # two-phase sweep (gpa_sweep_arm_two_phase; OFF by default: measured on 8 GPUs it gains nothing over the plain
            # gossip — 3.144 vs 3.140 ms per C3 step — because block-minimum bounds of single planes stay ~0.1-0.5 % below
            # the key-based thresholds of one GPU and the two extra flag barriers per peak cost what the better start saves):
            # per peak a float table [world][tiles of k_mr_interp (64 x 128 pixels)]
            self._tiles = (-(-n // 64)) * (-(-m // 128))
            # every rank must take part in the two flag barriers of every peak: only when no share is empty
            everybody = all(hi > lo for r in range(world) for lo, hi, _st in shard_units_interleaved(self.n_peaks, counts, world, r))
            self.two_phase = self.gossip and bool(two_phase) and everybody and self.ranges is inter
            total = off_best + (-(-P * world * self._tiles * 4 // 256) * 256 if self.two_phase else 0)
            self.arena = PeerArena(total, group=group, device=self.dev)
            self._off = {"key": off_key, "lockin": off_lock, "grad": off_grad, "hint": off_hint, "best": off_best}
            self.keys = self.arena.tensor(off_key, (P, n, m), torch.int64)
            self.lockin = self.arena.tensor(off_lock, (P, n, m), torch.complex128 if self.out_f64 else torch.complex64)
            self.grad = self.arena.tensor(off_grad, (P, n, m, 2), torch.float64 if self.out_f64 else torch.float32)
            self._status = self.arena.tensor(2048, (4,), torch.int32)
            if dst == "rows":
                self.dst_ranks, self.dst_rows = list(range(world)), -(-n // world)
            else:
                self.dst_ranks, self.dst_rows = [int(dst)], n
            r0 = self.dst_ranks.index(rank) * self.dst_rows if rank in self.dst_ranks else 0
            self.rows = (min(r0, n), min(r0 + self.dst_rows, n)) if rank in self.dst_ranks else (0, 0)
        else:
            self.arena = None
            self.keys = torch.zeros((P, n, m), dtype=torch.int64, device=self.dev)
            self.rows = (0, n) if (dst is None or dst == rank or world == 1) else (0, 0)
        self.kidx = torch.empty((P, n, m), dtype=torch.int32, device=self.dev)
        self.w = torch.empty((P, 2, n, m), dtype=torch.float64 if self.out_f64 else torch.float32, device=self.dev) if want_w else None
        if self.dev.type == "cuda":
            self._axes = [(torch.from_numpy(p.wx).to(self.dev), torch.from_numpy(p.wy).to(self.dev)) for p in self.plans]

    # ---- helpers --------------------------------------------------------------------------------
    def _flag_off(self, phase, peak, src=0):
        return ((phase * self.n_peaks + peak) * 8 + src) * 8

    def _signal(self, phase, peak, targets):
        from . import _lib, engine
        slots = (ctypes.c_void_p * len(targets))(*[self.arena.addr(t, self._flag_off(phase, peak, self.rank)) for t in targets])
        _lib.check(self.lib.gpa_peer_signal(slots, len(targets), self.epoch, engine._stream()))
        engine._count(1)

    def _wait(self, phase, peak):
        from . import _lib, engine
        _lib.check(self.lib.gpa_peer_wait(ctypes.c_void_p(self.arena.addr(self.rank, self._flag_off(phase, peak))), self.world,
                                          self.epoch, self.timeout_s, engine._ptr(self._status), engine._stream()))
        engine._count(1)

    def _mark(self, peak, name):
        if self.record:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._events[peak][name] = ev

    def check(self):
        """Raise if a flag wait timed out since the last check (host-synchronising)."""
        if self.transport == "peer":
            st = int(self._status[0].item())
            if st:
                self._status.zero_()
                from ._lib import GpaError
                raise GpaError(f"rank {self.rank}: peer flag wait timed out (missing source rank {st - 1})")

    def timings(self):
        """Per-peak phase durations (ms) of the last recorded call; synchronises."""
        torch.cuda.synchronize(self.dev)
        out = []
        order = ["start", "argmax", "peers_ready", "merged", "finalized", "delivered"]
        for evs in self._events or []:
            row = {}
            prev = "start"
            for name in order[1:]:
                if name in evs:
                    row[name + "_ms"] = evs[prev].elapsed_time(evs[name])
                    prev = name
            row["total_ms"] = evs["start"].elapsed_time(evs[prev])
            if getattr(self, "_base", None) is not None:
                row["start_offset_ms"] = self._base.elapsed_time(evs["start"])      # from the caller's stream entering the call
                row["frame_ms"] = self._base.elapsed_time(self._tail)               # ... to the caller's stream past the join
            out.append(row)
        return out

    def close(self):
        if self.arena is not None:
            self.arena.close()
            self.arena = None

    # ---- one frame ------------------------------------------------------------------------------
    def __call__(self, img_dev, grad_mode=0, join=True):
        """join=False (peer transport, frame streams): the caller's stream does NOT wait for the peaks' streams, so the
        exchange / finalize tail of this frame overlaps the arg-max of the next call; call join() before using the
        results on the caller's stream (they are reused by the next call on the same per-peak streams, in order)."""
        self._join = bool(join) or self.transport != "peer"
        if self.transport == "peer":
            return self._run_peer(img_dev, grad_mode)
        return self._run_collective(img_dev, grad_mode)

    def join(self):
        """Make the current stream wait for everything the last call(s) enqueued on the per-peak streams."""
        main = torch.cuda.current_stream(self.dev)
        for ev in self._pending:
            main.wait_event(ev)
        self._pending = []

    def _per_peak(self, fn):
        side = self._streams
        if side is None:
            for p in range(self.n_peaks):
                fn(p)
            return
        main = torch.cuda.current_stream(self.dev)
        if self.record:
            self._base = torch.cuda.Event(enable_timing=True)
            self._base.record(main)
        start = torch.cuda.Event()
        start.record(main)
        self._pending = []
        for p in range(self.n_peaks):
            side[p].wait_event(start)
            with torch.cuda.stream(side[p]):
                fn(p)
            done = torch.cuda.Event()
            done.record(side[p])
            self._pending.append(done)
        if getattr(self, "_join", True):
            self.join()
        if self.record:
            self._tail = torch.cuda.Event(enable_timing=True)
            self._tail.record(main)

    def _outs(self):
        return [{"key": self.keys[p], "lockin": self.lockin[p], "grad": self.grad[p], "kidx": self.kidx[p],
                 "w": self.w[p] if self.w is not None else None, "rows": self.rows} for p in range(self.n_peaks)]

    def _run_peer(self, img_dev, grad_mode):
        from . import _lib, engine
        lib, world, rank, n, m = self.lib, self.world, self.rank, self.n, self.m
        self.epoch += 1
        everyone = list(range(world))
        want_grad = grad_mode != engine.GRAD_NONE
        cb, rb = (16, 8) if self.out_f64 else (8, 4)
        if self.record:
            self._events = [dict() for _ in range(self.n_peaks)]

        def peak(p):
            plan = self.plans[p]
            lo, hi, step = self.ranges[p]
            self._mark(p, "start")
            self.keys[p].zero_()
            if hi > lo:
                if self.gossip:      # own array first, then the peers'
                    order = [rank] + [r for r in everyone if r != rank]
                    hp = (ctypes.c_void_p * world)(*[self.arena.addr(r, self._off["hint"] + p * self._hint_n * 8) for r in order])
                    _lib.check(lib.gpa_sweep_arm_gossip(hp, world, self.epoch & 0xFFFFFFFF))
                    if self.two_phase:
                        vp = ctypes.c_void_p * world
                        bp = vp(*[self.arena.addr(r, self._off["best"] + p * world * self._tiles * 4) for r in order])
                        fa = vp(*[self.arena.addr(r, self._flag_off(_PH_BOUNDS, p, rank)) for r in order])
                        fb = vp(*[self.arena.addr(r, self._flag_off(_PH_WAVE0, p, rank)) for r in order])
                        _lib.check(lib.gpa_sweep_arm_two_phase(
                            bp, fa, fb, ctypes.c_void_p(self.arena.addr(rank, self._flag_off(_PH_BOUNDS, p))),
                            ctypes.c_void_p(self.arena.addr(rank, self._flag_off(_PH_WAVE0, p))), rank, self.epoch,
                            self.timeout_s, engine._ptr(self._status)))
                plan.argmax(img_dev, self.keys[p], lo, hi, step)
            self._mark(p, "argmax")
            if world > 1:
                self._signal(_PH_ARGMAX, p, everyone)
                self._wait(_PH_ARGMAX, p)
                self._mark(p, "peers_ready")
                ptrs = (ctypes.c_void_p * world)(*[self.arena.addr(r, self._off["key"] + p * n * m * 8) for r in everyone])
                _lib.check(lib.gpa_key_merge(ptrs, world, rank, n * m, engine._stream()))
                engine._count(1)
                self._signal(_PH_MERGED, p, everyone)
                self._wait(_PH_MERGED, p)
                self._mark(p, "merged")
            lock_dst = (ctypes.c_void_p * len(self.dst_ranks))(*[self.arena.addr(r, self._off["lockin"] + p * n * m * cb) for r in self.dst_ranks])
            grad_dst = (ctypes.c_void_p * len(self.dst_ranks))(*[self.arena.addr(r, self._off["grad"] + p * n * m * 2 * rb) for r in self.dst_ranks])
            mr = plan.mr
            ws = plan._workspace()
            _lib.check(lib.gpa_sweep_finalize_mr_sharded(
                *plan._geom(), lo, hi, step, mr["S"], mr["Ra_x"], mr["Ra_y"], _lib.as_pf(mr["taps_bx"]), _lib.as_pf(mr["taps_by"]),
                mr["Rb"], *plan._split_geom(), engine._ptr(self.keys[p]), self.krefs[p][0], self.krefs[p][1], grad_mode,
                int(self.out_f64), lock_dst, grad_dst if want_grad else None, len(self.dst_ranks), self.dst_rows,
                int(rank == 0), engine._ptr(ws), ws.numel(), engine._stream()))
            engine._count(1)
            self._mark(p, "finalized")
            if world > 1:
                self._signal(_PH_DELIVERED, p, self.dst_ranks)
                if rank in self.dst_ranks:
                    self._wait(_PH_DELIVERED, p)
            r0, r1 = self.rows
            if r1 > r0:       # k-index and w of this rank's rows, decoded from the merged keys
                kk = self.keys[p, r0:r1]
                _lib.check(lib.gpa_key_to_kidx(engine._ptr(kk), engine._ptr(self.kidx[p, r0:r1]), kk.numel(), engine._stream()))
                engine._count(1)
                if self.w is not None:
                    wx_d, wy_d = self._axes[p]
                    _lib.check(lib.gpa_key_to_w(engine._ptr(kk), kk.numel(), n * m, engine._ptr(wx_d), engine._ptr(wy_d), plan.wy.size,
                                                int(plan.cand_mode == engine.CAND_LIST), int(self.out_f64),
                                                ctypes.c_void_p(self.w[p].data_ptr() + r0 * m * rb), engine._stream()))
                    engine._count(1)
            self._mark(p, "delivered")
            if self.after_peak is not None:
                self.after_peak(p)
        self._per_peak(peak)
        return self._outs()

    def _run_collective(self, img_dev, grad_mode):
        """Round-1 transport: one MAX all-reduce of the stacked keys, one SUM reduction of the zero-padded payload.
        The collectives are issued when every rank has finished its local work (an early collective makes NCCL's
        CTAs spin on the slower peer while holding SMs the sweep kernels need)."""
        from . import _lib, engine
        world, rank, n, m, dev, P = self.world, self.rank, self.n, self.m, self.dev, self.n_peaks
        if world == 1 and not (self._streams is not None and all(p.mr is not None and p.mr_in_flight == p.wy.size for p in self.plans)):
            outs = [plan.run(img_dev, kref, grad_mode, out_f64=self.out_f64, want_w=self.want_w)
                    for plan, kref in zip(self.plans, self.krefs)]
            for o in outs:
                o["rows"] = (0, n)
            return outs
        keys = self.keys
        keys.zero_()

        def argmax_peak(p):
            lo, hi, step = self.ranges[p]
            if hi > lo:
                self.plans[p].argmax(img_dev, keys[p], lo, hi, step)
        self._per_peak(argmax_peak)
        if world > 1:
            dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=self.group)
        want_grad = grad_mode != engine.GRAD_NONE
        real = torch.float64 if self.out_f64 else torch.float32
        # payload buffer: [lockin (P,N,M,2) | grad (P,N,M,2)], zero where this rank owns nothing
        payload = torch.zeros((2 if want_grad else 1, P, n, m, 2), dtype=real, device=dev)
        lockin = torch.view_as_complex(payload[0])
        outs = [{"lockin": lockin[p], "grad": payload[1, p] if want_grad else None, "w": None, "kidx": None} for p in range(P)]

        def finalize_peak(p):
            lo, hi, step = self.ranges[p]
            if hi > lo:
                self.plans[p].finalize(img_dev, keys[p], self.krefs[p], grad_mode, out_f64=self.out_f64, plane_begin=lo,
                                       plane_end=hi, want_kidx=False, planes_valid=self.plans[p]._private, out=outs[p],
                                       plane_step=step)
        self._per_peak(finalize_peak)
        if world == 1:
            pass
        elif self.dst is None:
            dist.all_reduce(payload, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.reduce(payload, dst=self.dst, op=dist.ReduceOp.SUM, group=self.group)
        lib = self.lib
        _lib.check(lib.gpa_key_to_kidx(engine._ptr(keys), engine._ptr(self.kidx), keys.numel(), engine._stream()))
        for p, o in enumerate(outs):
            o["key"], o["kidx"], o["rows"] = keys[p], self.kidx[p], self.rows
            if self.w is not None:
                wx_d, wy_d = self._axes[p]
                _lib.check(lib.gpa_key_to_w(engine._ptr(keys[p]), n * m, n * m, engine._ptr(wx_d), engine._ptr(wy_d),
                                            self.plans[p].wy.size, int(self.plans[p].cand_mode == engine.CAND_LIST),
                                            int(self.out_f64), engine._ptr(self.w[p]), engine._stream()))
                o["w"] = self.w[p]
        return outs


class HostSweep:
    """cuGPA.wfr2_grad_opt for every peak of a frame on all GPUs of the box: NumPy image in on rank 0, the reference's
    NumPy arrays out on rank 0 ('lockin' (N,M) c16, 'w' (2,N,M) f8, 'grad' (N,M,2) f8 per peak; cuGPA.py:41-87).
    SPMD: every rank constructs it and calls it once per frame; only rank 0 passes the image and gets the results.

    Data path per frame: rank 0 copies the frame to its GPU and casts it to float32 into its peer arena; the other
    ranks pull it over NVLink.  The sweep runs k-grid sharded with dst='rows' (ShardedSweep), so rank r ends up with
    rows [r N/W, (r+1) N/W) of every output array, already widened to float64 / complex128, and copies them over
    ITS OWN PCIe link into one page-locked shared-memory segment that all ranks map — the device-to-host traffic of
    the reference's result arrays (48 B per pixel and peak) is spread over W links, and each peak's copy overlaps the
    arg-max of the next peak.  The arrays rank 0 returns are views of that segment, valid until the next call."""

    def __init__(self, shape, sigma, kvecs, kw, kstep, group=None, grad=None, planes_in_flight=None):
        from multiprocessing import shared_memory
        from . import _lib, cuGPA, engine
        from .peer import PeerArena
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dev = engine.require_cuda()
        self.lib = _lib.load()
        self.n, self.m = int(shape[0]), int(shape[1])
        self.kvecs = [tuple(map(float, k)) for k in kvecs]
        self.grad_mode = cuGPA._grad_mode(grad)
        P, n, m = len(self.kvecs), self.n, self.m
        plans = []
        for k in self.kvecs:
            wxs, wys = engine.grid_axes(k[0], k[1], kw, kstep)
            plans.append(engine.SweepPlan((n, m), wxs, wys, sigma, device=self.dev, private_ws=True,
                                          planes_in_flight=planes_in_flight))
        self.sweep = ShardedSweep(plans, self.kvecs, group=group, dst="rows", out_f64=True, want_w=True, transport="peer")
        self.sweep.after_peak = self._copy_out
        self.img_arena = PeerArena(_FLAG_BYTES + n * m * 4, group=group, device=self.dev)
        self.img = self.img_arena.tensor(_FLAG_BYTES, (n, m), torch.float32)
        self._stage = torch.empty((n, m), dtype=torch.float64, device=self.dev) if self.rank == 0 else None
        # shared, page-locked result segment: [host flags 4096 B | lockin P c16 | w P 2 f8 | grad P .. 2 f8]
        self._seg_bytes = 4096 + P * n * m * 48
        name = [None]
        if self.rank == 0:
            self._shm = shared_memory.SharedMemory(create=True, size=self._seg_bytes)
            name[0] = self._shm.name
            if self.world > 1:
                _touch_interleaved(self._shm.buf, self._seg_bytes)
        if self.world > 1:
            dist.broadcast_object_list(name, src=0, group=group)
        if self.rank != 0:
            self._shm = shared_memory.SharedMemory(name=name[0])
            try:        # the creator unlinks; attaching processes must not (Python < 3.13 registers them with the tracker)
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:
                pass
        buf = np.frombuffer(self._shm.buf, dtype=np.uint8, count=self._seg_bytes)
        self._buf_addr = buf.ctypes.data
        _lib.check(self.lib.gpa_host_register(ctypes.c_void_p(self._buf_addr), self._seg_bytes))
        self._hflags = np.frombuffer(self._shm.buf, dtype=np.int64, count=64)
        if self.rank == 0:
            self._hflags[:] = 0
        off = 4096
        self._h_lock = np.frombuffer(self._shm.buf, dtype=np.complex128, count=P * n * m, offset=off).reshape(P, n, m)
        off += P * n * m * 16
        self._h_w = np.frombuffer(self._shm.buf, dtype=np.float64, count=P * 2 * n * m, offset=off).reshape(P, 2, n, m)
        off += P * n * m * 16
        self._h_grad = np.frombuffer(self._shm.buf, dtype=np.float64, count=P * n * m * 2, offset=off).reshape(P, n, m, 2)
        self.calls = 0
        if self.world > 1:
            dist.barrier(group=group)

    def _host_ptr(self, arr, *index):
        view = arr[index]
        return ctypes.c_void_p(view.ctypes.data)

    def _copy_out(self, p):
        """Rows of peak p that live on this rank -> the shared host segment (on peak p's stream)."""
        from . import _lib, engine
        r0, r1 = self.sweep.rows
        if r1 <= r0:
            return
        sw, m, st = self.sweep, self.m, engine._stream()
        rows = r1 - r0
        _lib.check(self.lib.gpa_peer_copy(self._host_ptr(self._h_lock, p, r0), engine._ptr(sw.lockin[p, r0:r1]), rows * m * 16, st))
        if self.grad_mode != 2:
            _lib.check(self.lib.gpa_peer_copy(self._host_ptr(self._h_grad, p, r0), engine._ptr(sw.grad[p, r0:r1]), rows * m * 16, st))
        for c in range(2):
            _lib.check(self.lib.gpa_peer_copy(self._host_ptr(self._h_w, p, c, r0), engine._ptr(sw.w[p, c, r0:r1]), rows * m * 8, st))

    def __call__(self, image=None):
        from . import _lib, engine
        self.calls += 1
        ep, world, rank = self.calls, self.world, self.rank
        st = engine._stream()
        if rank == 0:
            arr = np.ascontiguousarray(image, dtype=np.float64)
            if arr.shape != (self.n, self.m):
                raise ValueError(f"image must have shape {(self.n, self.m)}")
            self._hflags[8] = ep                       # go: the workers may enqueue this frame
            self._stage.copy_(torch.from_numpy(arr), non_blocking=True)
            _lib.check(self.lib.gpa_cast_f64_to_f32(engine._ptr(self._stage), engine._ptr(self.img), arr.size, st))
            if world > 1:
                slots = (ctypes.c_void_p * world)(*[self.img_arena.addr(r, 0) for r in range(world)])
                _lib.check(self.lib.gpa_peer_signal(slots, world, ep, st))
        else:
            while self._hflags[8] < ep:                # host-side: no device spin while rank 0's caller is busy elsewhere
                pass
            _lib.check(self.lib.gpa_peer_wait(ctypes.c_void_p(self.img_arena.addr(rank, 0)), 1, ep, self.sweep.timeout_s,
                                              engine._ptr(self.sweep._status), st))
            _lib.check(self.lib.gpa_peer_copy(engine._ptr(self.img), ctypes.c_void_p(self.img_arena.addr(0, _FLAG_BYTES)),
                                              self.n * self.m * 4, st))
        self.sweep(self.img, self.grad_mode)
        torch.cuda.current_stream(self.dev).synchronize()
        self._hflags[16 + rank] = ep                   # this rank's rows are in the segment
        if rank != 0:
            return None
        for r in range(1, world):
            while self._hflags[16 + r] < ep:
                pass
        self.sweep.check()
        return [{"lockin": self._h_lock[p], "w": self._h_w[p], "grad": self._h_grad[p]} if self.grad_mode != 2 else
                {"lockin": self._h_lock[p], "w": self._h_w[p]} for p in range(len(self.kvecs))]

    @property
    def d2h_bytes_per_rank(self):
        r0, r1 = self.sweep.rows
        return (r1 - r0) * self.m * 48 * len(self.kvecs)

    def close(self):
        if getattr(self, "_shm", None) is None:
            return
        torch.cuda.synchronize(self.dev)
        self.sweep.close()
        self.img_arena.close()
        self.lib.gpa_host_unregister(ctypes.c_void_p(self._buf_addr))
        self._hflags = self._h_lock = self._h_w = self._h_grad = None
        try:
            self._shm.close()
        except BufferError:
            pass
        if self.rank == 0:
            self._shm.unlink()
        self._shm = None


def _touch_interleaved(buf, nbytes):
    """First-touch the pages of the shared result segment with the NUMA interleave policy, so that the GPUs of both
    sockets write their rows into local memory half of the time instead of all funnelling into rank 0's node
    (8 GPUs x PCIe gen5 otherwise queue behind one socket's memory controllers and the inter-socket link).
    Best effort: without NUMA (or without the syscall) the pages are simply touched."""
    import platform
    arr = np.frombuffer(buf, dtype=np.uint8, count=nbytes)
    libc = None
    try:
        if platform.machine() == "x86_64":
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(0xFFFF)                    # nodes 0..15; absent nodes are ignored by the kernel
            if libc.syscall(238, 3, ctypes.byref(mask), 17) != 0:     # set_mempolicy(MPOL_INTERLEAVE, ...)
                libc = None
    except Exception:
        libc = None
    arr[::4096] = 0
    if libc is not None:
        libc.syscall(238, 0, None, 0)                        # back to MPOL_DEFAULT


_sweeps = {}


def sharded_sweep(img_dev, plans, krefs, grad_mode=0, group=None, dst=None, transport="auto"):
    """All peaks of one frame, k-grid sharded over the ranks of `group` (function form of ShardedSweep; the
    executor — plans, peer arena, streams — is cached per plan set, so every rank must call this with the same
    plans in the same order).  dst=None: every rank gets the whole result with the collective transport; the peer
    transport delivers rows [r N/W, (r+1) N/W) to rank r.  dst=<rank>: everything lands on that rank.
    Returns [dict(key, lockin, grad, kidx, rows)] per peak; the tensors are reused by the next call."""
    key = (tuple(id(p) for p in plans), id(group), dst, transport)
    sw = _sweeps.get(key)
    if sw is None:
        if len(_sweeps) >= 4:
            for old in _sweeps.values():
                old.close()
            _sweeps.clear()
        if dst is None and transport == "auto":
            transport = "collective"
        sw = ShardedSweep(plans, krefs, group=group, dst=dst, transport=transport)
        _sweeps[key] = sw
    sw.krefs = [tuple(map(float, k)) for k in krefs]
    return sw(img_dev, grad_mode)


def release():
    """Close every cached executor (collective: unmaps the peer arenas)."""
    for sw in _sweeps.values():
        sw.close()
    _sweeps.clear()


_side_streams = {}


def _peak_streams(dev, n):
    """One side stream per peak and device, created once, in DESCENDING priority: peak 0's kernels are scheduled
    first, so its key exchange and finalize overlap the arg-max of the later peaks."""
    pool = _side_streams.setdefault(str(dev), [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=dev, priority=-(n - 1 - len(pool))))
    return pool[:n]
