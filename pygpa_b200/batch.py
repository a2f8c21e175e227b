"""Frame-batch driver (BASELINE config 4): a time series of frames, each run through the whole
adaptive pipeline — sweep per primary k-vector, displacement field, Lawler-Fujita undistortion —
entirely on the device.  Frames are independent, so a multi-GPU job shards them over the ranks
with no collective on the data path (weak scaling); results are gathered by the caller.

The per-frame chain is the device-resident twin of
    u = extract_displacement_field(frame, kvecs)        (geometric_phase_analysis.py:907-932)
    corrected = undistort_image(frame, -u)               (:935-974; u = -extracted field, tests :63)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import engine, solvers

__all__ = ["FramePipeline", "shard_frames", "process_frames"]


class FramePipeline:
    """Plans (candidate axes, taps, scratch) for one frame shape, reused for every frame."""

    def __init__(self, shape, kvecs, sigma=None, kwscale=2.5, ksteps=3, n_grid=None, device=None, streams=1, graphs=False):
        """streams > 1: `submit` spreads consecutive frames over that many CUDA streams (each with its own scratch), so the
        kernels of one frame that cannot fill the GPU (1024^2 frames: FFT strips, spline prefilter, reductions) overlap
        with another frame's.
        graphs: `submit` captures the whole per-frame chain (~250 launches) into one CUDA graph per stream on its second
        use and replays it afterwards — the host then spends microseconds per frame instead of milliseconds.  Frames must
        be CUDA tensors of one shape and dtype; the returned tensors of a stream are overwritten by that stream's next
        frame (copy them out first)."""
        self.device = device or engine.require_cuda()
        self._streams = [torch.cuda.Stream(self.device) for _ in range(max(1, streams))] if (streams > 1 or graphs) else []
        self._turn = 0
        self._graphs = bool(graphs)
        self._slots = [None] * len(self._streams)       # per stream: dict(graph, static_in, out, undistort) once captured
        self._warm = [0] * len(self._streams)
        self.kvecs = np.asarray(kvecs, dtype=np.float64)
        norms = np.linalg.norm(self.kvecs, axis=1)
        self.kw = float(norms.mean() / kwscale)
        self.sigma = int(np.ceil(1 / norms.min())) if sigma is None else sigma
        self.kstep = 2 * self.kw / (n_grid - 0.5) if n_grid else self.kw / ksteps
        self.plans = []
        for pk in self.kvecs:
            wxs, wys = engine.grid_axes(pk[0], pk[1], self.kw, self.kstep)
            self.plans.append(engine.SweepPlan(shape, wxs, wys, self.sigma, device=self.device))

    def displacement(self, frame_dev):
        """frame (N, M) float32 CUDA tensor (mean already removed) -> u (2, N, M) float64 CUDA tensor,
        the field extract_displacement_field returns."""
        phs, wts = [], []
        for plan, pk in zip(self.plans, self.kvecs):
            res = plan.run(frame_dev, pk, engine.GRAD_NONE, want_kidx=False)
            ph, wt = solvers.phase_weight(res["lockin"], 2 * int(self.sigma))
            phs.append(ph)
            wts.append(wt)
        return solvers.displacement_from_phases(self.kvecs, torch.stack(phs), torch.stack(wts))

    def __call__(self, frame, undistort=True):
        """One frame (NumPy or tensor) -> dict(u, corrected) of CUDA tensors (float64)."""
        arr = frame if isinstance(frame, torch.Tensor) else np.asarray(frame, dtype=np.float64)
        f32 = engine.image_to_device(arr - arr.mean(), self.device)
        u = self.displacement(f32)
        out = {"u": u}
        if undistort:
            # the extracted field is minus the physical displacement (lock-in phase = -2 pi k.u); the ORIGINAL float64
            # frame is resampled, as undistort_image(frame, -u) does (zeros outside the frame, geometric_phase_analysis.py:973)
            out["corrected"] = solvers.undistort(solvers.to_device_f64(arr, self.device), -u)
        return out


    def submit(self, frame, undistort=True):
        """Like calling the pipeline, but on the next stream of the pool (round-robin).  The returned tensors are complete
        after `join()` (or a synchronisation of the device)."""
        if not self._streams:
            return self(frame, undistort)
        caller = torch.cuda.current_stream(self.device)
        slot_i = self._turn % len(self._streams)
        s = self._streams[slot_i]
        self._turn += 1
        s.wait_stream(caller)                      # the frame may have been produced on the caller's stream
        if self._graphs and isinstance(frame, torch.Tensor) and frame.is_cuda:
            slot = self._slots[slot_i]
            if slot is None and self._warm[slot_i] >= 1:
                slot = self._slots[slot_i] = self._capture(s, frame, undistort)
            if slot is not None and slot["undistort"] == undistort and slot["static_in"].shape == frame.shape \
                    and slot["static_in"].dtype == frame.dtype:
                with torch.cuda.stream(s):
                    slot["static_in"].copy_(frame)
                    slot["graph"].replay()
                return slot["out"]
        with torch.cuda.stream(s):
            res = self(frame, undistort)           # eager (also the warm-up that sizes this stream's scratch before a capture)
        self._warm[slot_i] += 1
        for v in res.values():
            v.record_stream(caller)
        return res

    def _capture(self, stream, frame, undistort):
        static_in = torch.empty_like(frame)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            static_in.copy_(frame)
        stream.synchronize()
        with torch.cuda.graph(graph, stream=stream):
            out = self(static_in, undistort)
        return {"graph": graph, "static_in": static_in, "out": out, "undistort": undistort}

    def join(self):
        """Make the caller's stream wait for everything submitted so far."""
        caller = torch.cuda.current_stream(self.device)
        for s in self._streams:
            caller.wait_stream(s)


def shard_frames(n_frames, world, rank):
    """Contiguous, balanced share of the frame indices for `rank`."""
    lo = (n_frames * rank) // world
    hi = (n_frames * (rank + 1)) // world
    return range(lo, hi)


def process_frames(frames, kvecs, sigma=None, n_grid=None, undistort=True, group=None):
    """Run this rank's share of `frames` (sequence or (T, N, M) array).  Returns {index: dict of
    NumPy arrays}; no communication happens here."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = shard_frames(len(frames), world, rank)
    pipe = None
    out = {}
    for t in mine:
        frame = np.asarray(frames[t])
        if pipe is None:
            pipe = FramePipeline(frame.shape, kvecs, sigma=sigma, n_grid=n_grid)
        res = pipe(frame, undistort=undistort)
        out[t] = {k: v.cpu().numpy() for k, v in res.items()}
    return out
