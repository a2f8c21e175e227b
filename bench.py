#!/usr/bin/env python
"""Benchmark of the adaptive-GPA hot path (BASELINE.json metric: Mpixel*kvec/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE config 3 — synthetic twisted-bilayer moire 2048x2048,
3 primary k-vectors, adaptive sweep 41x41 candidates per peak, sigma = 10 px.  One "step" is
the full sweep of one frame (all peaks): arg-max over the candidate grid + winner lock-in,
k-index and phase gradient.  N > 1 shards the k-grid over the GPUs (strong scaling): the packed
keys are max-reduced by our own kernel over NVLink peer memory and every rank writes the winners
it owns straight into rank 0's arrays (pygpa_b200/dist.py, csrc/peer.cu).

  value  device-resident throughput, CUDA events around exactly K steps, max over ranks
  e2e    same sweep through the public host API (pygpa_b200.cuGPA.wfr2_grad_opt_peaks; the
         per-peak cuGPA.wfr2_grad_opt figure is kept beside it): NumPy image in pinned host
         memory in, float64/complex128 NumPy arrays out, H2D and D2H inside the timed region
  roofline       dominant kernel of the timed region, FP32 FMA pipe, peak measured live
  cpu_baseline   the oracle port of the reference CPU path on a bounded sample (rank 0, N=1)
  multi_gpu      (N > 1) per-phase times of the exchange, per-rank compute spread, bit-identity
                 check against the single-GPU result
  configs        the other BASELINE configs: C2 (1 GPU), C4 frames/s (frames sharded), C5 (k-grid sharded)

--impl reference times the reference CPU algorithm (oracle port, all host cores) instead.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "C3: synthetic TBG moire 2048x2048, 3 peaks x 41x41 k-vectors, sigma=10"
SIZE, NGRID, SIGMA = 2048, 41, 10
METRIC, UNIT = "adaptive_gpa_sweep_throughput", "Mpixel*kvec/s"


def units_per_step(size=SIZE, ngrid=NGRID, peaks=3):
    return size * size * peaks * ngrid * ngrid


def f_alg(taps, nx):
    """Algorithmic flops per pixel*kvec: separable, demodulate then real taps."""
    return 4 * taps * (1 + 1.0 / nx) + 8 + 2.0 / nx


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ----------------------------------------------------------------------------------------------
_W = {}


def _worker_init(image, sigma):
    import oracle
    _W["oracle"], _W["image"], _W["sigma"] = oracle, image, sigma


def _worker_run(args):
    klist, kref = args
    out = _W["oracle"].wfr_sweep_klist(_W["image"], _W["sigma"], klist, kref)
    return float(np.abs(out["lockin"]).sum())


def cpu_sample(image, cfg, n_cand, procs, pool=None):
    """Time the oracle's sweep loop on `n_cand` candidates of peak 0 split over `procs` processes.
    Returns Mpixel*kvec/s."""
    import oracle
    k = cfg["ks"][0]
    wxs, wys = oracle.candidate_axes(k[0], k[1], cfg["kw"], cfg["kstep"])
    klist = np.stack(np.meshgrid(wxs, wys, indexing="ij"), axis=-1).reshape(-1, 2)[:n_cand]
    t0 = time.perf_counter()
    if procs == 1:
        oracle.wfr_sweep_klist(image, cfg["sigma"], klist, k)
    else:
        pool.map(_worker_run, [(part, k) for part in np.array_split(klist, procs)])
    dt = time.perf_counter() - t0
    return image.size * len(klist) / dt / 1e6, dt


def run_reference(args, cfg):
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    per_step = procs            # one candidate per core per step
    ctx = mp.get_context("fork")
    with ctx.Pool(procs, initializer=_worker_init, initargs=(cfg["image"], cfg["sigma"])) as pool:
        for _ in range(args.warmup):
            cpu_sample(cfg["image"], cfg, per_step, procs, pool)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_sample(cfg["image"], cfg, per_step, procs, pool)
        dt = time.perf_counter() - t0
    value = cfg["image"].size * per_step * args.steps / dt / 1e6
    sample = f"{per_step} candidates of peak 0 per step ({procs} processes x 1), full 2048x2048 frame, float64 FFT path"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampling in the background.  The process is started before the warm-up (it needs
    ~0.2 s to produce its first line) and samples every 50 ms with a timestamp; stop() keeps the
    samples that fall inside the timed window [mark(), stop()]."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t0 = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def mark(self):
        import datetime
        self.t0 = datetime.datetime.now()

    def stop(self):
        import datetime
        t1 = datetime.datetime.now()
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [[c.strip() for c in r.split(",")] for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 9]
        os.unlink(self.f.name)

        def stamp(r):
            try:
                return datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f")
            except ValueError:
                return None
        pad = datetime.timedelta(milliseconds=60)
        inside = [r for r in rows if stamp(r) is not None and self.t0 is not None and self.t0 - pad <= stamp(r) <= t1 + pad]
        used, where = (inside, "timed region") if inside else (rows[-4:], "last samples before the end of the timed region")
        if not used:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[2]) for r in used]
        reasons = set()
        for r in used:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[6:10]):
                if v.lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(used[0][3]), "reasons": sorted(reasons),
                "samples": len(used), "window": where, "power_w_max": max(float(r[4]) for r in used)}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def _profile_read(lib, names, reset_with=None):
    from pygpa_b200 import _lib
    tot, n = ctypes.c_double(0), ctypes.c_int(0)
    out = {}
    for name in names:
        _lib.check(lib.gpa_profile_read(name.encode(), ctypes.byref(tot), ctypes.byref(n), 0))
        out[name] = (tot.value, n.value)
    if reset_with:
        _lib.check(lib.gpa_profile_read(reset_with.encode(), ctypes.byref(tot), ctypes.byref(n), 1))
    return out


def _hbm_peak():
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        return json.load(open(pk_path)).get("hbm_gbs", 6551.0), "hbm_gbs of MEASURED_PEAKS.json"
    return 6551.0, "fallback 6551 GB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def _timed(torch, fn, reps=1):
    """(result, ms per call) after one warm-up call, CUDA events on the current stream."""
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return r, e0.elapsed_time(e1) / reps


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    from pygpa_b200 import _lib, cuGPA, engine
    from pygpa_b200 import dist as gdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = engine.require_cuda()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ks = cfg["ks"]
    img_host = torch.from_numpy(cfg["image"]).pin_memory()            # float64, pinned
    img = engine.image_to_device(img_host.numpy(), dev)
    plans = []
    for k in ks:
        wxs, wys = engine.grid_axes(k[0], k[1], cfg["kw"], cfg["kstep"])
        assert len(wxs) == NGRID and len(wys) == NGRID
        # one GPU: the peaks run back to back on one stream, so the per-kernel CUDA-event times of the roofline are
        # those of kernels that have the GPU to themselves; N > 1: private workspaces, one priority stream per peak
        plans.append(engine.SweepPlan(img.shape, wxs, wys, cfg["sigma"], device=dev, private_ws=world > 1))
    taps = 2 * plans[0].rx + 1
    transport = os.environ.get("GPA_TRANSPORT", "auto")
    sweep = gdist.ShardedSweep(plans, ks, dst=0, transport=transport, gossip=os.environ.get("GPA_GOSSIP", "1") != "0",
                                two_phase=os.environ.get("GPA_TWO_PHASE", "0") == "1") if world > 1 else None
    # N > 1, frame stream: the caller's stream does not join the per-peak streams between frames (exactly like the single
    # stream of N = 1, nothing synchronises between the K steps), so the exchange tail of frame t overlaps frame t + 1;
    # multi_gpu.frame_latency_ms is the time of ONE isolated frame
    pipelined = sweep is not None and sweep.transport == "peer" and os.environ.get("GPA_PIPELINE", "1") != "0"

    def step():
        if sweep is not None:
            return sweep(img, join=not pipelined)
        return [plan.run(img, k) for plan, k in zip(plans, ks)]

    # measured FP32 peak (register-only FFMA2 loop) before anything else warms the chip differently
    fp32_measured = None
    if rank == 0:
        ws = engine.workspace(4 << 20, dev)
        tf = ctypes.c_double(0)
        _lib.check(lib.gpa_fp32_peak_tflops(engine._ptr(ws), ws.numel(), ctypes.byref(tf), engine._stream()))
        fp32_measured = tf.value

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    if pipelined:
        sweep.join()
    barrier()
    lib.gpa_profile_enable(1)
    launches0 = engine.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    e0.record()
    for _ in range(args.steps):
        step()
    if pipelined:
        sweep.join()
    e1.record()
    barrier()
    lib.gpa_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    gpu_launches = engine.launch_count - launches0
    names = ("k_mr_pass1", "k_mr_pass1a", "k_mr_pass1b", "k_mr_pass2", "k_mr_pass2a", "k_mr_pass2b", "k_mr_order", "k_mr_interp", "k_mr_finalize",
             "k_key_merge", "k_pass1", "k_pass2_argmax", "k_finalize")
    kernels = _profile_read(lib, names, reset_with="k_pass1")
    if sweep is not None:
        sweep.check()

    # The same kernels with the (exact) pruning switched off: the interpolation kernel then executes its full
    # algorithmic work, which is the duration its pipe utilisation is computed from (pruned launches skip work).
    unpruned, unpruned_ms = {}, None
    if world == 1 and plans[0].mr is not None:
        engine.set_pruning(False)
        try:
            step()
            torch.cuda.synchronize()
            lib.gpa_profile_enable(1)
            eu0, eu1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eu0.record()
            for _ in range(3):
                step()
            eu1.record()
            torch.cuda.synchronize()
            unpruned_ms = eu0.elapsed_time(eu1) / 3
            lib.gpa_profile_enable(0)
            unpruned = _profile_read(lib, ("k_mr_interp",), reset_with="k_pass1")
        finally:
            engine.set_pruning(True)

    units = units_per_step()
    value = units * args.steps / (ms_total / 1e3) / 1e6

    # ---- N > 1: where the time goes, and is the result the single-GPU one? -------------------------------------
    multi = None
    if sweep is not None:
        lat = []
        for _ in range(5):           # isolated frames: barrier, one frame, join
            barrier()
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record()
            sweep(img)
            l1.record()
            torch.cuda.synchronize()
            lat.append(l0.elapsed_time(l1))
        lat_t = torch.tensor([sorted(lat)[len(lat) // 2]], device=dev)
        dist.all_reduce(lat_t, op=dist.ReduceOp.MAX)
        frame_latency_ms = float(lat_t.item())
        sweep.record = True
        outs = sweep(img)
        phases = sweep.timings() if sweep.transport == "peer" else []
        sweep.record = False
        every = [None] * world
        dist.all_gather_object(every, phases)
        check = None
        if rank == 0:
            single = [p.run(img, k) for p, k in zip(plans, ks)]
            torch.cuda.synchronize()
            check = {
                "keys_equal": all(torch.equal(a["key"], b["key"]) for a, b in zip(outs, single)),
                "lockin_equal": all(torch.equal(torch.view_as_real(a["lockin"]), torch.view_as_real(b["lockin"])) for a, b in zip(outs, single)),
                "grad_equal": all(torch.equal(a["grad"], b["grad"]) for a, b in zip(outs, single)),
                "kidx_equal": all(torch.equal(a["kidx"], b["kidx"]) for a, b in zip(outs, single)),
                "checksum_keys": int(torch.stack([o["key"] for o in outs]).sum().item()),
                "checksum_keys_single_gpu": int(torch.stack([o["key"] for o in single]).sum().item()),
                "against": "SweepPlan.run of the same plans on rank 0 alone (torch.equal on every output array)",
            }
        if rank == 0 and every and every[0]:
            def agg(name, fn):
                return [fn([r[p].get(name, 0.0) for r in every]) for p in range(len(ks))]
            multi = {
                "transport": "peer memory over NVLink (csrc/peer.cu): k_key_merge + flag signal/wait + owner-writes finalize; no NCCL on the data path",
                "shares": "interleaved (peak, plane) units, unit u -> rank u % N" if sweep.ranges[0][2] > 1 or world == 1 else "contiguous",
                "per_peak_ms": {
                    "argmax_max_over_ranks": agg("argmax_ms", max), "argmax_min_over_ranks": agg("argmax_ms", min),
                    "wait_for_peers_max": agg("peers_ready_ms", max),
                    "key_merge_incl_flags_max": agg("merged_ms", max),
                    "finalize_owner_writes_max": agg("finalized_ms", max),
                    "delivery_wait_max": agg("delivered_ms", max),
                },
                "comm_ms": {
                    "key_exchange": float(sum(agg("merged_ms", max))),
                    "peer_skew_wait": float(sum(agg("peers_ready_ms", max))),
                    "payload_delivery_wait": float(sum(agg("delivered_ms", max))),
                    "what": "summed over the 3 peaks, max over ranks, CUDA events on each peak's stream in one instrumented step; the exchange "
                            "of peaks 0 and 1 overlaps the arg-max of the later peaks, only the last peak's is exposed",
                },
                "per_rank_compute_ms": {
                    "max": max(max(r[p]["argmax_ms"] for p in range(len(ks))) for r in every),
                    "min": min(max(r[p]["argmax_ms"] for p in range(len(ks))) for r in every),
                    "per_rank": [max(r[p]["argmax_ms"] for p in range(len(ks))) for r in every],
                    "step_end_ms_per_rank": [max(r[p]["total_ms"] for p in range(len(ks))) for r in every],
                    "what": "time from the start of the step to the end of the rank's last arg-max kernel (the peaks' streams overlap); "
                            "step_end = its last delivery wait"},
                "frame_latency_ms": frame_latency_ms,
                "pipelined_frames": bool(pipelined),
                "threshold_gossip": bool(sweep.gossip),
                "two_phase_sweep": bool(getattr(sweep, "two_phase", False)),
                "exposed_tail_ms": max(max(r[p]["total_ms"] for p in range(len(ks))) - max(r[p]["argmax_ms"] for p in range(len(ks))) for r in every),
                "k_key_merge_ms_per_step_rank0": kernels["k_key_merge"][0] / args.steps,
                "check": check,
            }
        elif rank == 0:
            multi = {"transport": "torch.distributed collectives (MAX all-reduce of the keys + SUM reduce of the payload)", "check": check,
                     "frame_latency_ms": frame_latency_ms}

    # ---- end to end through the public host API ------------------------------------------------------------------
    n_e2e = max(3, min(args.steps, 5))
    e2e = None
    first_call_ms = None
    per_call = None
    if world == 1:
        cuGPA.clear_plans()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cuGPA.wfr2_grad_opt(img_host.numpy(), cfg["sigma"], ks[0][0], ks[0][1], cfg["kw"], cfg["kstep"])
        first_call_ms = (time.perf_counter() - t0) * 1e3

        def per_peak_calls():
            return [cuGPA.wfr2_grad_opt(img_host.numpy(), cfg["sigma"], k[0], k[1], cfg["kw"], cfg["kstep"]) for k in ks]
        for _ in range(2):
            outs_h = per_peak_calls()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            outs_h = per_peak_calls()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        per_call = {"value": units / dt / 1e6, "ms_per_step": dt * 1e3,
                    "h2d_bytes_per_step": int(img_host.numel() * 8 * len(ks)),
                    "d2h_bytes_per_step": int(sum(v.nbytes for o in outs_h for v in o.values())),
                    "api": "pygpa_b200.cuGPA.wfr2_grad_opt x3, one call per peak as extract_displacement_field issues them "
                           "(every call uploads the frame and returns after its own D2H)"}
        del outs_h
        cuGPA.clear_plans()
    # batched / SPMD entry: the same call on 1 and on N GPUs
    shape = tuple(cfg["image"].shape)

    def e2e_step():
        return cuGPA.wfr2_grad_opt_peaks(img_host.numpy() if rank == 0 else None, cfg["sigma"], ks, cfg["kw"], cfg["kstep"], shape=shape)
    try:
        for _ in range(3):
            host = e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            host = e2e_step()
        dt_local = (time.perf_counter() - t0) / n_e2e          # rank 0 returns when every rank's rows are on the host
        dt = torch.tensor([dt_local], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        if rank == 0:
            d2h = sum(v.nbytes for o in host for v in o.values())
            e2e = {"value": units / dt / 1e6, "unit": UNIT, "ms_per_step": dt * 1e3,
                   "h2d_bytes_per_step": int(img_host.numel() * 8), "d2h_bytes_per_step": int(d2h),
                   "d2h_bytes_per_step_per_gpu": int(d2h // world),
                   "api": "pygpa_b200.cuGPA.wfr2_grad_opt_peaks (all peaks of the frame in one call): pinned NumPy float64 frame in on rank 0 -> "
                          "'lockin' c128, 'w' f64, 'grad' f64 NumPy arrays per peak out on rank 0; each peak's D2H overlaps the next peak's sweep"
                          + ("" if world == 1 else "; every GPU copies its rows of the results into one page-locked shared segment over its own PCIe link"),
                   "per_peak_calls": per_call, "first_call_ms_cold_plan": first_call_ms}
            # sanity: the batched arrays equal the per-call API's (N = 1) / carry the single-GPU checksum
            e2e["lockin_abs_sum"] = float(sum(np.abs(o["lockin"][::16, ::16]).sum() for o in host))
        del host
    except Exception as exc:      # the e2e leg must never take the device-resident line down
        if rank == 0:
            e2e = {"value": per_call["value"] if per_call else None, "unit": UNIT, "error": repr(exc), "per_peak_calls": per_call,
                   "h2d_bytes_per_step": per_call["h2d_bytes_per_step"] if per_call else None,
                   "d2h_bytes_per_step": per_call["d2h_bytes_per_step"] if per_call else None}
    cuGPA.clear_plans()

    # ---- the rest of the adaptive pipeline and the reference-GPU baseline (rank 0, N = 1 only) ----
    pipeline = None
    cugpa = None
    if world == 1 and rank == 0 and not args.no_extras:
        pipeline, cugpa = bench_tail(torch, lib, cfg, img, ks, step, ms_total / args.steps, unpruned_ms)

    # ---- the other BASELINE configs --------------------------------------------------------------------------------
    configs = None
    if not args.no_extras:
        if sweep is not None:
            sweep.close()
            sweep = None
        del plans
        engine.release_workspaces()
        torch.cuda.empty_cache()
        configs = bench_configs(torch, dist, lib, world, rank, dev)

    if rank == 0:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        sm_max = peaks.get("sm_max_mhz", 1965.0)
        fp32_theory = 148 * 128 * 2 * sm_max * 1e6 / 1e12          # TFLOP/s, FFMA at the max SM clock
        fp32_peak = fp32_measured or fp32_theory
        # dominant kernel = the one with the largest summed duration inside the timed region
        dom = max(kernels, key=lambda k_: kernels[k_][0])
        d_ms, d_n = kernels[dom]
        mr = cfg["_mr"]
        split = cfg["_split"]
        if dom == "k_mr_interp":
            # multirate arg-max: per unit (pixel*kvec) the y-interpolation issues W_eff real-tap x
            # complex-sample MACs (4 flop each) + 3 flop for |sf|^2, the x-interpolation 1/S of that.
            S = mr["S"]
            w_eff = (11 + 10 * (S - 1)) / S
            flop_per_unit = 4 * w_eff * (1 + 1.0 / S) + 3
            basis = f"multirate form, stride {S}: 4*{w_eff:.2f}*(1+1/{S})+3 = {flop_per_unit:.1f} flop per pixel*kvec (tile halos not counted)"
        elif dom == "k_mr_pass2b":
            S, jb = mr["S"], 2 * split["H"] + 1
            flop_per_unit = (4 * jb + 12) / (S * S)
            basis = f"split pass 2, coarse stage, stride {S}: (4*{jb} + 12)/S^2 = {flop_per_unit:.2f} flop per pixel*kvec"
        elif dom == "k_mr_pass2":
            S = mr["S"]
            flop_per_unit = (4 * (2 * mr["Ra_x"] + 1) + 8 * S) / (S * S)
            basis = f"decimating pass 2, stride {S}: (4*Ta + 8*S)/S^2 = {flop_per_unit:.1f} flop per pixel*kvec"
        else:
            flop_per_unit = 4 * taps + 8
            basis = f"direct form: 4*T+8 = {flop_per_unit} flop per pixel*kvec"
        d_flops_per_launch = flop_per_unit * units * args.steps / world / max(d_n, 1)
        d_avg_s = d_ms / max(d_n, 1) / 1e3
        achieved = d_flops_per_launch / d_avg_s / 1e12 if d_n else None
        pruning_note = None
        if dom in unpruned and unpruned[dom][1]:
            u_avg_s = unpruned[dom][0] / unpruned[dom][1] / 1e3
            pruned_equiv = achieved
            achieved = d_flops_per_launch / u_avg_s / 1e12
            pruning_note = {"avg_launch_ms_pruning_off": u_avg_s * 1e3, "achieved_algorithmic_over_pruned_duration": pruned_equiv,
                            "frac_algorithmic_over_pruned_duration": pruned_equiv / fp32_peak,
                            "what": "achieved / frac use the duration of the same kernel with the exact pruning OFF (it then executes "
                                    "exactly the algorithmic work, measured live in this run); with pruning ON the launch in the timed "
                                    "region skips most candidates per tile, so algorithmic flops / its duration exceeds the FP32 peak"}
        survey_alg = (4 * taps + 8) * units * args.steps / world / max(d_n, 1) / d_avg_s / 1e12 if d_n else None
        roofline = {
            "bound": "fp32", "kernel": dom, "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": achieved / fp32_peak if achieved else None, "traffic": None,
            "work_basis": basis,
            "peak_source": ("measured in this run: register-only fma.rn.f32x2 loop, 4 CTAs x 256 threads per SM, best of 5 (gpa_fp32_peak_tflops)"
                            if fp32_measured else f"148 SM x 128 FFMA lanes x 2 x {sm_max:.0f} MHz"),
            "peak_theoretical": fp32_theory, "frac_of_theoretical_peak": achieved / fp32_theory if achieved else None,
            "avg_launch_ms": d_avg_s * 1e3, "launches": d_n, "share_of_step": d_ms / ms_total,
            # SURVEY section 8d charges the direct form's 4T+8 flop per unit whatever the kernel really does;
            # the multirate factorisation executes ~5x fewer, so these two exceed 1 by design
            "frac_direct_form_equivalent": survey_alg / fp32_peak if survey_alg else None,
            "step_frac_direct_form_equivalent": f_alg(taps, NGRID) * units * args.steps / (ms_total / 1e3) / 1e12 / fp32_peak / world,
            "hbm_gbs_algorithmic": 28.0 * SIZE * SIZE * 3 * args.steps / (ms_total / 1e3) / 1e9,
            "kernels_ms_per_step": {k_: v[0] / args.steps for k_, v in kernels.items() if v[1]},
        }
        if pruning_note:
            roofline["pruning"] = pruning_note
        if world > 1:
            roofline["note"] = ("N > 1: the three peaks run on three priority streams of every rank, so these per-kernel event times "
                                "overlap and include sharing the SMs with the other peaks' kernels; the N = 1 line has the isolated ones")
        prof = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
        if os.path.exists(prof):
            tr = json.load(open(prof))
            roofline["traffic"] = tr.get(dom)
            if isinstance(tr.get(dom + "_ncu"), dict):
                roofline["ncu_timed_launch"] = tr[dom + "_ncu"]
        cpu = None
        if world == 1 and not args.no_cpu:
            n_cand = 8
            v, dt_cpu = cpu_sample(cfg["image"], cfg, n_cand, 1)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"first {n_cand} candidates of peak 0 on the full 2048x2048 frame ({dt_cpu:.1f} s), "
                             "oracle.wfr_sweep_klist = the reference's single-threaded float64 FFT loop"}
            if not args.no_extras:
                cpu["tail"] = cpu_tail(cfg)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "working set per step (3 x 1.44 GB of first-pass planes) exceeds L2; no flush needed",
                       "parallelism": f"k-grid sharded over {world} GPU(s)", "filter": f"{taps} taps (4.5 sigma)",
                       "argmax_form": (f"multirate, stride {mr['S']}" + (f", split pass 2 ({2 * split['H'] + 1} coarse taps per candidate)" if split else "") if mr else "direct"),
                       "pruning": "exact per-tile branch and bound on (results bit-identical to off; value_unpruned gives the off figure)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(gpu_launches), "roofline": roofline, "cpu_baseline": cpu,
            "value_unpruned": units / (unpruned_ms / 1e3) / 1e6 if unpruned_ms else None,
            "ms_per_step_unpruned": unpruned_ms,
            "multi_gpu": multi, "pipeline": pipeline, "cugpa_equivalent": cugpa, "configs": configs,
            "ms_per_2048_frame": ms_total / args.steps,
        }
        print(json.dumps(line), flush=True)
    if sweep is not None:
        sweep.close()
    if world > 1:
        dist.destroy_process_group()


def bench_tail(torch, lib, cfg, img, ks, step, sweep_ms, unpruned_ms):
    """Rest of the adaptive pipeline on the C3 frame (device resident) + the cuGPA-equivalent baseline."""
    from pygpa_b200 import _lib, engine, solvers
    from pygpa_b200 import property_extract as pe_b200, unit_cell_averaging as uc_b200
    outs = step()
    dr = 2 * cfg["sigma"]
    pw = [solvers.phase_weight(o["lockin"], dr) for o in outs]
    phases, weights = torch.stack([a for a, _ in pw]), torch.stack([b for _, b in pw])
    for _ in range(2):
        u = solvers.displacement_from_phases(ks, phases, weights)
    torch.cuda.synchronize()
    lib.gpa_profile_enable(1)
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0e.record()
    u = solvers.displacement_from_phases(ks, phases, weights)
    t1e.record()
    torch.cuda.synchronize()
    tail_ms = t0e.elapsed_time(t1e)
    img64 = img.double()
    rec = solvers.undistort(img64, u)            # warm-up (function attributes, workspace growth)
    torch.cuda.synchronize()
    t0e.record()
    rec = solvers.undistort(img64, u)
    t1e.record()
    torch.cuda.synchronize()
    lf_ms = t0e.elapsed_time(t1e)
    del rec
    grads64 = torch.stack([o["grad"] for o in outs]).double()

    def timed(fn):
        return _timed(torch, fn)
    jac, j_ms = timed(lambda: pe_b200.phasegradient2J_device(ks, grads64, weights, 1.0, add_identity=True))
    _props, p_ms = timed(lambda: solvers.props_from_jac(jac))
    _dec, d_ms2 = timed(lambda: solvers.gaussian_deconvolve(u, cfg["sigma"], dr))
    _cell, c_ms = timed(lambda: uc_b200.unit_cell_average_device(img64, ks[:2], u, z=2))
    _fit, f_ms = timed(lambda: solvers.fit_plane_huber(u[0]))

    # BASELINE config 3 names "+ weighted phase_unwrap": per peak, weights sqrt(|lockin| / max) as iterate_GPA
    # uses them (geometric_phase_analysis.py:141), kmax = 100
    def unwrap_peaks():
        its = []
        for o in outs:
            ph, amp, amax = solvers.lockin_phase_amp(o["lockin"], 0)
            _phi, it = solvers.unwrap(psi=ph, weight=solvers.weight_sqrt_norm(amp, amax), kmax=100, return_iters=True)
            its.append(it)
        return its
    lib.gpa_profile_enable(0)            # keep these solves out of the uw_* event timers of the tail above
    uw_iters, uw3_ms = timed(unwrap_peaks)
    prof = _profile_read(lib, ("uw_setup", "uw_poisson_solve", "uw_vector_ops", "k_lstsq", "k_norm_axis0", "lf_prefilter",
                               "k_invert_u", "k_resample"), reset_with="k_lstsq")
    hbm, hbm_src = _hbm_peak()
    npx = SIZE * SIZE
    uw_ms = prof["uw_poisson_solve"][0] + prof["uw_vector_ops"][0]      # 2 solves x 10 iterations
    uw_gbs = 168.0 * npx * 20 / (uw_ms / 1e3) / 1e9
    uw3_gbs = 168.0 * npx * sum(uw_iters) / (uw3_ms / 1e3) / 1e9
    ls_gbs = 64.0 * npx * 2 / (prof["k_lstsq"][0] / 1e3) / 1e9
    # K4: prefilter 32 B/pixel once + 24 B/pixel/iteration x 36 + final resample 16 B/pixel
    k4_bytes = (32.0 + 24.0 * 36 + 16.0) * npx
    k4_ms = prof["lf_prefilter"][0] + prof["k_invert_u"][0] + prof["k_resample"][0]
    k4_gbs = k4_bytes / (k4_ms / 1e3) / 1e9 if k4_ms else None
    pipeline = {
        "what": "tail of extract_displacement_field on the C3 frame, device resident: 2 x per-pixel least squares + "
                "2 x PCG unwrap (kmax=10) ; then undistort_image (Lawler-Fujita, 35 iterations)",
        "tail_ms": tail_ms, "lawler_fujita_ms": lf_ms, "sweep_plus_tail_ms_per_2048_frame": sweep_ms + tail_ms,
        "sweep_plus_weighted_unwrap_ms_per_2048_frame": sweep_ms + uw3_ms,
        "sweep_ms_per_step_without_pruning": unpruned_ms,
        "hbm_peak_source": hbm_src,
        "unwrap_pcg": {"ms": uw_ms, "bound": "hbm", "achieved_gbs": uw_gbs, "peak_gbs": hbm, "frac": uw_gbs / hbm,
                       "basis": "168 B/pixel/iteration (SURVEY 8d) x 2 solves x 10 iterations"},
        "lstsq": {"ms": prof["k_lstsq"][0], "bound": "hbm", "achieved_gbs": ls_gbs, "peak_gbs": hbm, "frac": ls_gbs / hbm,
                  "basis": "64 B/pixel (SURVEY 8d) x 2 solves"},
        "lawler_fujita": {"ms": k4_ms, "bound": "hbm/L2", "achieved_gbs": k4_gbs, "peak_gbs": hbm, "frac": k4_gbs / hbm if k4_gbs else None,
                          "basis": "SURVEY 8d: 32 B/pixel prefilter + 24 B/pixel/iteration x 36 + 16 B/pixel resample = 912 B/pixel; the 36 iterations "
                                   "run in registers inside ONE k_invert_u launch, so the executed DRAM traffic is far below this figure"},
        "kernels_ms": {k_: v[0] for k_, v in prof.items()},
        "weighted_phase_unwrap_3_peaks": {"ms": uw3_ms, "pcg_iterations": uw_iters, "kmax": 100, "achieved_gbs": uw3_gbs, "frac": uw3_gbs / hbm,
                                          "what": "phase_unwrap(angle(lockin), sqrt(|lockin| / max)) per peak, device resident; 168 B/pixel/iteration basis"},
        "consumers_ms": {
            "what": "SURVEY 8f rows on the C3 frame, device resident, one call each (float64, HBM-bound streaming kernels)",
            "phasegradient2Jac": j_ms, "phasegradient2Jac_gbs": 104.0 * SIZE * SIZE / (j_ms / 1e3) / 1e9,
            "props_from_Jac": p_ms, "props_from_Jac_gbs": 64.0 * SIZE * SIZE / (p_ms / 1e3) / 1e9,
            "gaussian_deconvolve_2_planes": d_ms2, "unit_cell_average_z2": c_ms, "fit_plane_huber": f_ms,
            "basis": "104 B/pixel (3 x (2 gradients + 1 weight) in, 4 out) and 64 B/pixel (4 in, 4 out) against hbm_gbs of MEASURED_PEAKS.json",
        },
    }
    # reference cuGPA on this GPU: CuPy and pyGPA itself do not exist on the GPU box -> the torch transcription of its algorithm
    cugpa = None
    try:
        from baseline import cugpa_torch
        n_cand = 24
        k0 = ks[0]
        kind = "cuGPA-equivalent: torch.cuda transcription of pyGPA/cuGPA.py:41-87 (complex128, cuFFT, unfused); CuPy is not installed"
        cugpa_torch.wfr2_grad_opt(cfg["image"], cfg["sigma"], k0[0], k0[1], cfg["kw"], cfg["kstep"], max_candidates=4)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cugpa_torch.wfr2_grad_opt(cfg["image"], cfg["sigma"], k0[0], k0[1], cfg["kw"], cfg["kstep"], max_candidates=n_cand)
        torch.cuda.synchronize()
        dt1 = time.perf_counter() - t0
        t0 = time.perf_counter()
        cugpa_torch.wfr2_grad_opt(cfg["image"], cfg["sigma"], k0[0], k0[1], cfg["kw"], cfg["kstep"], max_candidates=4)
        torch.cuda.synchronize()
        dt0 = time.perf_counter() - t0
        per_cand = (dt1 - dt0) / (n_cand - 4)
        step_s = per_cand * 3 * NGRID * NGRID + 3 * (dt0 - 4 * per_cand)
        cugpa = {"kind": kind, "ms_per_candidate": per_cand * 1e3, "ms_per_step_extrapolated": step_s * 1e3,
                 "value": units_per_step() / step_s / 1e6, "unit": UNIT,
                 "sample": f"{n_cand} candidates of peak 0 on the 2048x2048 frame, API level (H2D + .get()), extrapolated linearly to 3 x 1681"}
    except Exception as exc:   # the baseline must never take the benchmark down
        cugpa = {"unavailable": repr(exc)}
    engine.release_workspaces()
    return pipeline, cugpa


def cpu_tail(cfg):
    """CPU reference (oracle port, one core) timings of the pipeline tail, BASELINE.md section 3: bounded sizes."""
    import oracle
    rng = np.random.default_rng(0)
    out = {"kind": "port", "cores": 1}
    n = 1024
    yy, xx = np.meshgrid(np.arange(n), np.arange(n))
    truth = 40 * np.sin(xx / 150.0) + 30 * np.cos(yy / 170.0)
    psi = (truth + np.pi) % (2 * np.pi) - np.pi
    wgt = 0.2 + rng.random((n, n))
    t0 = time.perf_counter()
    oracle.phase_unwrap(psi, wgt, kmax=10)
    out["phase_unwrap_weighted_1024_kmax10_s"] = time.perf_counter() - t0
    ks = cfg["ks"]
    phases = np.stack([((2 * np.pi * (k[0] * xx.T + k[1] * yy.T) * 0.02 + np.pi) % (2 * np.pi)) - np.pi for k in ks])
    weights = 0.5 + rng.random((3, n, n))
    t0 = time.perf_counter()
    u = oracle.reconstruct_u_inv_from_phases(ks, phases, weights)
    out["reconstruct_u_inv_from_phases_1024_s"] = time.perf_counter() - t0
    n2 = 512
    u2 = 2.0 * np.stack([np.sin(np.arange(n2)[:, None] / 60.0) * np.ones((1, n2)), np.cos(np.arange(n2)[None, :] / 70.0) * np.ones((n2, 1))])
    t0 = time.perf_counter()
    oracle.undistort_image(rng.random((n2, n2)), u2)
    out["undistort_image_512_s"] = time.perf_counter() - t0
    out["sample"] = ("oracle port (NumPy/SciPy restatement of the reference, float64, single process) at bounded sizes: weighted phase_unwrap "
                     "kmax=10 and reconstruct_u_inv_from_phases at 1024^2, undistort_image at 512^2; all three scale linearly in the pixel "
                     "count (x4 / x4 / x16 for the 2048^2 frame)")
    return out


def bench_configs(torch, dist, lib, world, rank, dev):
    """The BASELINE configs other than the headline: C2 (1 GPU), C5 (k-grid sharded over the N GPUs, checked against one GPU),
    C4 (frames sharded over the N GPUs, frames/s)."""
    from pygpa_b200 import batch, engine, synth
    from pygpa_b200 import dist as gdist
    out = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def sweep_config(name, steps, check):
        cfg = synth.make_config_device(name, dev)
        img64 = cfg["image"]
        if world > 1:
            dist.broadcast(img64, 0)               # one frame for everybody (generated identically anyway)
        img = img64.float()
        del img64
        ks, ng = cfg["ks"], cfg["n_grid"]
        share = -(-ng // world) + 1 if world > 1 else None
        plans = []
        for k in ks:
            wxs, wys = engine.grid_axes(k[0], k[1], cfg["kw"], cfg["kstep"])
            plans.append(engine.SweepPlan(img.shape, wxs, wys, cfg["sigma"], device=dev, private_ws=world > 1, planes_in_flight=share))
        sw = gdist.ShardedSweep(plans, ks, dst=0) if world > 1 else None

        def step():
            return sw(img) if sw is not None else [p.run(img, k) for p, k in zip(plans, ks)]
        step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            outs = step()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        n = img.shape[0]
        units = n * n * len(ks) * ng * ng
        rec = {"workload": f"{name}: {n}x{n}, 3 peaks x {ng}x{ng} k-vectors, sigma=10", "n_gpus": world, "ms_per_step": float(ms.item()),
               "value": units / (float(ms.item()) / 1e3) / 1e6, "unit": UNIT, "steps": steps,
               "parallelism": f"k-grid sharded over {world} GPU(s)" + (", peer-memory key merge + owner-writes" if world > 1 else "")}
        if check and world > 1:
            keys = torch.stack([o["key"] for o in outs]).clone()
            if rank == 0:
                eq = True
                for p, (plan, k) in enumerate(zip(plans, ks)):
                    key1 = torch.zeros_like(keys[p])
                    plan.argmax(img, key1)          # the whole grid on rank 0 alone, in chunks of its resident planes
                    eq = eq and bool(torch.equal(key1, keys[p]))
                rec["check"] = {"keys_equal_single_gpu": eq, "checksum_keys": int(keys.sum().item())}
        elif check and rank == 0:
            rec["check"] = {"checksum_keys": int(torch.stack([o["key"] for o in outs]).sum().item())}
        if sw is not None:
            sw.check()
            sw.close()
        del plans, outs
        engine.release_workspaces()
        torch.cuda.empty_cache()
        return rec

    if world == 1:
        out["C2"] = sweep_config("C2", 10, False)
    out["C5"] = sweep_config("C5", 2 if world == 1 else 4, True)
    # C4: 512 LEEM-like frames of 1024^2, the whole adaptive pipeline per frame, frames sharded over the ranks
    total = 512
    mine = batch.shard_frames(total, world, rank)
    frames, ks = synth.frame_series_device(len(mine), dev, size=1024, t0=mine.start, total=total)
    c4_streams = 3       # consecutive frames on three streams: the under-filled kernels of a 1024^2 frame overlap (tools/perf_c4.py)
    pipe = batch.FramePipeline(frames.shape[1:], ks, sigma=10, n_grid=21, device=dev, streams=c4_streams)
    for i in range(min(2 * c4_streams, frames.shape[0])):
        pipe.submit(frames[i])
    pipe.join()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    acc = 0.0
    for i in range(frames.shape[0]):
        res = pipe.submit(frames[i])
    pipe.join()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out["C4"] = {"workload": "C4: 512 synthetic LEEM-like frames 1024x1024 (16 bit), per frame: adaptive sweep 3 x 21x21 + displacement "
                             "(2 x least squares + 2 x PCG unwrap) + Lawler-Fujita undistortion, device resident, consecutive frames on "
                             f"{c4_streams} CUDA streams per GPU",
                 "n_gpus": world, "frames": total, "frames_per_rank": len(mine), "seconds": float(ms.item()) / 1e3,
                 "frames_per_s": total / (float(ms.item()) / 1e3), "scaling": "weak (frames sharded, no collective on the data path)",
                 "mean_abs_u_last_frame_px": float(res["u"].abs().mean().item())}
    del frames, pipe, res
    engine.release_workspaces()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline sample")
    ap.add_argument("--no-extras", action="store_true", help="skip the pipeline-tail and cuGPA-equivalent measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference" and int(os.environ.get("RANK", "0")) != 0:
        return
    from pygpa_b200 import synth
    cfg = synth.make_config("C3")
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        from pygpa_b200 import _taps, engine
        mr = _taps.multirate_taps(SIZE, SIZE, float(cfg["sigma"]), _taps.DEFAULT_TRUNC)
        wxs, _wys = engine.grid_axes(cfg["ks"][0][0], cfg["ks"][0][1], cfg["kw"], cfg["kstep"])
        cfg["_mr"] = mr
        cfg["_split"] = _taps.split_taps(SIZE, mr, wxs) if mr is not None else None
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
