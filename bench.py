#!/usr/bin/env python
"""Benchmark of the adaptive-GPA hot path (BASELINE.json metric: Mpixel*kvec/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE config 3 — synthetic twisted-bilayer moire 2048x2048,
3 primary k-vectors, adaptive sweep 41x41 candidates per peak, sigma = 10 px.  One "step" is
the full sweep of one frame (all peaks): arg-max over the candidate grid + winner lock-in,
k-index and phase gradient.  N > 1 shards the k-grid over the GPUs (strong scaling) with one
NCCL MAX all-reduce of the packed keys and one SUM all-reduce of the payload.

  value  device-resident throughput, CUDA events around exactly K steps, max over ranks
  e2e    same sweep through the reference-facing API (pygpa_b200.cuGPA.wfr2_grad_opt per
         peak): NumPy image in pinned host memory in, float64/complex128 NumPy arrays out,
         H2D and D2H inside the timed region
  roofline       dominant kernel k_pass2 (arg-max sweep), FP32 FMA pipe
  cpu_baseline   the oracle port of the reference CPU path on a bounded sample (rank 0, N=1)

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

    ks = cfg["ks"]
    img_host = torch.from_numpy(cfg["image"]).pin_memory()            # float64, pinned
    img = engine.image_to_device(img_host.numpy(), dev)
    plans = []
    for k in ks:
        wxs, wys = engine.grid_axes(k[0], k[1], cfg["kw"], cfg["kstep"])
        assert len(wxs) == NGRID and len(wys) == NGRID
        # one GPU: the peaks run back to back on one stream, so the per-kernel CUDA-event times of the roofline are
        # those of kernels that have the GPU to themselves (private plans would overlap the peaks: 27.0 vs 27.55 ms)
        plans.append(engine.SweepPlan(img.shape, wxs, wys, cfg["sigma"], device=dev, private_ws=world > 1))
    taps = 2 * plans[0].rx + 1

    def step():
        return gdist.sharded_sweep(img, plans, ks, dst=0 if world > 1 else None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    lib.gpa_profile_enable(1)
    launches0 = engine.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    lib.gpa_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    gpu_launches = engine.launch_count - launches0
    tot, n = ctypes.c_double(0), ctypes.c_int(0)
    kernels = {}
    names = ("k_mr_pass1", "k_mr_pass1a", "k_mr_pass1b", "k_mr_pass2", "k_mr_pass2a", "k_mr_pass2b", "k_mr_order", "k_mr_interp", "k_mr_finalize",
             "k_pass1", "k_pass2_argmax", "k_finalize")
    for name in names:
        _lib.check(lib.gpa_profile_read(name.encode(), ctypes.byref(tot), ctypes.byref(n), 0))
        kernels[name] = (tot.value, n.value)
    _lib.check(lib.gpa_profile_read(b"k_pass1", ctypes.byref(tot), ctypes.byref(n), 1))
    # The same kernels with the (exact) pruning switched off: the interpolation kernel then executes its full
    # algorithmic work, which is the duration its pipe utilisation is computed from (pruned launches skip work).
    unpruned = {}
    if world == 1 and plans[0].mr is not None:
        engine.set_pruning(False)
        try:
            step()
            torch.cuda.synchronize()
            lib.gpa_profile_enable(1)
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            lib.gpa_profile_enable(0)
            for name in ("k_mr_interp",):
                _lib.check(lib.gpa_profile_read(name.encode(), ctypes.byref(tot), ctypes.byref(n), 0))
                unpruned[name] = (tot.value, n.value)
            _lib.check(lib.gpa_profile_read(b"k_pass1", ctypes.byref(tot), ctypes.byref(n), 1))
        finally:
            engine.set_pruning(True)

    units = units_per_step()
    value = units * args.steps / (ms_total / 1e3) / 1e6

    # ---- end to end through the public API (rank 0 drives; N>1 shares the work the same way) ----
    e2e = None
    if world == 1:
        def e2e_step():
            outs = [cuGPA.wfr2_grad_opt(img_host.numpy(), cfg["sigma"], k[0], k[1], cfg["kw"], cfg["kstep"]) for k in ks]
            return outs
        for _ in range(2):
            outs = e2e_step()
        torch.cuda.synchronize()
        n_e2e = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            outs = e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        d2h = sum(v.nbytes for o in outs for v in o.values())
        e2e = {"value": units / dt / 1e6, "unit": UNIT, "ms_per_step": dt * 1e3,
               "h2d_bytes_per_step": int(img_host.numel() * 8 * len(ks)), "d2h_bytes_per_step": int(d2h),
               "api": "pygpa_b200.cuGPA.wfr2_grad_opt x3 (NumPy float64 in pinned memory -> NumPy c128/f64 out)"}
    else:
        def e2e_step():
            if rank == 0:
                staged = torch.from_numpy(img_host.numpy()).to(dev, non_blocking=True)
            else:
                staged = torch.empty(img_host.shape, dtype=torch.float64, device=dev)
            dist.broadcast(staged, 0)
            im = torch.empty(staged.shape, dtype=torch.float32, device=dev)
            _lib.check(lib.gpa_cast_f64_to_f32(ctypes.c_void_p(staged.data_ptr()), ctypes.c_void_p(im.data_ptr()),
                                               staged.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
            outs = gdist.sharded_sweep(im, plans, ks, dst=0)
            if rank == 0:
                host = [(cuGPA._to_host(o["lockin"]), cuGPA._to_host(o["grad"]), cuGPA._to_host(o["kidx"])) for o in outs]
                torch.cuda.current_stream().synchronize()
                return host
            return None
        host = None
        for _ in range(3):
            host = e2e_step()       # keep the previous result alive like the timed loop does: both pinned buffer sets get cached
        barrier()
        n_e2e = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            host = e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        if rank == 0:
            d2h = sum(t.numel() * t.element_size() for h in host for t in h)
            e2e = {"value": units / dt / 1e6, "unit": UNIT, "ms_per_step": dt * 1e3,
                   "h2d_bytes_per_step": int(img_host.numel() * 8), "d2h_bytes_per_step": int(d2h),
                   "api": "pinned float64 frame on rank 0 -> NCCL broadcast -> pygpa_b200.dist.sharded_sweep -> c64/f32/i32 to rank-0 host"}

    # ---- the rest of the adaptive pipeline and the reference-GPU baseline (rank 0, N = 1 only) ----
    pipeline = None
    cugpa = None
    if world == 1 and rank == 0 and not args.no_extras:
        from pygpa_b200 import solvers
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
        rec = solvers.undistort(img.double(), u)            # warm-up (function attributes, workspace growth)
        torch.cuda.synchronize()
        t0e.record()
        rec = solvers.undistort(img.double(), u)
        t1e.record()
        torch.cuda.synchronize()
        lf_ms = t0e.elapsed_time(t1e)
        # ---- consumers of the sweep: timed once each after a warm-up call ----
        from pygpa_b200 import property_extract as pe_b200, unit_cell_averaging as uc_b200
        grads64 = torch.stack([o["grad"] for o in outs]).double()
        img64 = img.double()

        def timed(fn):
            fn()
            torch.cuda.synchronize()
            t0e.record()
            r = fn()
            t1e.record()
            torch.cuda.synchronize()
            return r, t0e.elapsed_time(t1e)
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
        lib.gpa_profile_enable(1)
        # transparency: the same sweep with the (exact) branch-and-bound pruning switched off
        engine.set_pruning(False)
        step()
        torch.cuda.synchronize()
        t0e.record()
        for _ in range(3):
            step()
        t1e.record()
        torch.cuda.synchronize()
        unpruned_ms = t0e.elapsed_time(t1e) / 3
        engine.set_pruning(True)
        lib.gpa_profile_enable(0)
        prof = {}
        for name in ("uw_setup", "uw_poisson_solve", "uw_vector_ops", "k_lstsq", "k_norm_axis0", "lf_prefilter", "k_invert_u", "k_resample"):
            _lib.check(lib.gpa_profile_read(name.encode(), ctypes.byref(tot), ctypes.byref(n), 0))
            prof[name] = (tot.value, n.value)
        _lib.check(lib.gpa_profile_read(b"k_lstsq", ctypes.byref(tot), ctypes.byref(n), 1))
        hbm = 6551.0
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk_path):
            hbm = json.load(open(pk_path)).get("hbm_gbs", hbm)
        npx = SIZE * SIZE
        uw_ms = prof["uw_poisson_solve"][0] + prof["uw_vector_ops"][0]      # 2 solves x 10 iterations
        uw_gbs = 168.0 * npx * 20 / (uw_ms / 1e3) / 1e9
        ls_gbs = 64.0 * npx * 2 / (prof["k_lstsq"][0] / 1e3) / 1e9
        pipeline = {
            "what": "tail of extract_displacement_field on the C3 frame, device resident: 2 x per-pixel least squares + "
                    "2 x PCG unwrap (kmax=10) ; then undistort_image (Lawler-Fujita, 35 iterations)",
            "tail_ms": tail_ms, "lawler_fujita_ms": lf_ms, "sweep_plus_tail_ms_per_2048_frame": ms_total / args.steps + tail_ms,
            "sweep_ms_per_step_without_pruning": unpruned_ms,
            "unwrap_pcg": {"ms": uw_ms, "bound": "hbm", "achieved_gbs": uw_gbs, "peak_gbs": hbm, "frac": uw_gbs / hbm,
                           "basis": "168 B/pixel/iteration (SURVEY 8d) x 2 solves x 10 iterations"},
            "lstsq": {"ms": prof["k_lstsq"][0], "bound": "hbm", "achieved_gbs": ls_gbs, "peak_gbs": hbm, "frac": ls_gbs / hbm,
                      "basis": "64 B/pixel (SURVEY 8d) x 2 solves"},
            "kernels_ms": {k_: v[0] for k_, v in prof.items()},
            "weighted_phase_unwrap_3_peaks": {"ms": uw3_ms, "pcg_iterations": uw_iters, "kmax": 100,
                                              "what": "phase_unwrap(angle(lockin), sqrt(|lockin| / max)) per peak, device resident"},
            "consumers_ms": {
                "what": "SURVEY 8f rows on the C3 frame, device resident, one call each (float64, HBM-bound streaming kernels)",
                "phasegradient2Jac": j_ms, "phasegradient2Jac_gbs": 104.0 * SIZE * SIZE / (j_ms / 1e3) / 1e9,
                "props_from_Jac": p_ms, "props_from_Jac_gbs": 64.0 * SIZE * SIZE / (p_ms / 1e3) / 1e9,
                "gaussian_deconvolve_2_planes": d_ms2, "unit_cell_average_z2": c_ms, "fit_plane_huber": f_ms,
                "basis": "104 B/pixel (3 x (2 gradients + 1 weight) in, 4 out) and 64 B/pixel (4 in, 4 out) against hbm_gbs of MEASURED_PEAKS.json",
            },
        }
        # reference cuGPA on this GPU: the CuPy module itself if importable, else its torch transcription
        try:
            n_cand = 24
            k0 = ks[0]
            try:
                import cupy  # noqa: F401
                kind = "pyGPA.cuGPA.wfr2_grad_opt (CuPy)"
                raise ImportError("pyGPA itself is not available on the GPU box")
            except ImportError:
                from baseline import cugpa_torch
                kind = "torch.cuda transcription of pyGPA/cuGPA.py:41-87 (complex128, cuFFT, unfused)"
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
                     "value": units / step_s / 1e6, "unit": UNIT,
                     "sample": f"{n_cand} candidates of peak 0 on the 2048x2048 frame, API level (H2D + .get()), extrapolated linearly to 3 x 1681"}
        except Exception as exc:   # the baseline must never take the benchmark down
            cugpa = {"unavailable": repr(exc)}
        engine.release_workspaces()

    if rank == 0:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        sm_max = peaks.get("sm_max_mhz", 1965.0)
        fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12          # TFLOP/s, FFMA at the max SM clock
        # dominant kernel = the one with the largest summed duration inside the timed region
        dom = max(kernels, key=lambda k_: kernels[k_][0])
        d_ms, d_n = kernels[dom]
        mr = plans[0].mr
        if dom == "k_mr_interp":
            # multirate arg-max: per unit (pixel*kvec) the y-interpolation issues W_eff real-tap x
            # complex-sample MACs (4 flop each) + 3 flop for |sf|^2, the x-interpolation 1/S of that.
            S = mr["S"]
            w_eff = (11 + 10 * (S - 1)) / S
            flop_per_unit = 4 * w_eff * (1 + 1.0 / S) + 3
            basis = f"multirate form, stride {S}: 4*{w_eff:.2f}*(1+1/{S})+3 = {flop_per_unit:.1f} flop per pixel*kvec (tile halos not counted)"
        elif dom == "k_mr_pass2b":
            # split pass 2, coarse-rate stage: per COARSE output and candidate (2H+1) real-tap x complex-sample
            # MACs (4 flop) + one demodulation and one de-rotation (6 flop each); a unit is S^2 coarse outputs
            S, jb = mr["S"], 2 * plans[0].split["H"] + 1
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
            # k_mr_interp with pruning skips most (tile, plane, candidate) triples, so algorithmic flops / pruned
            # duration exceeds the pipe peak; the roofline fraction is taken on the launch that executes all of them
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
            "peak_source": f"148 SM x 128 FFMA lanes x 2 x {sm_max:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json; that file has no fp32 figure)",
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
            roofline["note"] = ("N > 1: the three peaks run on three streams of every rank, so these per-kernel event times "
                                "overlap and include sharing the SMs with the other peaks' kernels; the N = 1 line has the isolated ones")
        prof = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
        if os.path.exists(prof):
            roofline["traffic"] = json.load(open(prof)).get(dom)
        cpu = None
        if world == 1 and not args.no_cpu:
            n_cand = 8
            v, dt_cpu = cpu_sample(cfg["image"], cfg, n_cand, 1)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"first {n_cand} candidates of peak 0 on the full 2048x2048 frame ({dt_cpu:.1f} s), "
                             "oracle.wfr_sweep_klist = the reference's single-threaded float64 FFT loop"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "working set per step (3 x 1.44 GB of first-pass planes) exceeds L2; no flush needed",
                       "parallelism": f"k-grid sharded over {world} GPU(s)", "filter": f"{taps} taps (4.5 sigma)",
                       "argmax_form": (f"multirate, stride {mr['S']}" + (f", split pass 2 ({2 * plans[0].split['H'] + 1} coarse taps per candidate)" if plans[0].split else "") if mr else "direct"),
                       "pruning": "exact per-tile branch and bound on (results bit-identical to off; pipeline.sweep_ms_per_step_without_pruning gives the off time)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(gpu_launches), "roofline": roofline, "cpu_baseline": cpu,
            "pipeline": pipeline, "cugpa_equivalent": cugpa,
            "ms_per_2048_frame": ms_total / args.steps,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
