"""C4 frames/s on one GPU as a function of the number of streams the frames are spread over; host enqueue time beside it."""
import sys, time, torch
sys.path.insert(0, '.')
from pygpa_b200 import batch, engine, synth
dev = engine.require_cuda()
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ks_list = [a for a in sys.argv[2:]] or ["1", "2", "3", "1g", "2g", "3g"]      # "2g": two streams, one CUDA graph each
frames, ks = synth.frame_series_device(nf, dev, size=1024, t0=0, total=512)
ref = None
for k in ks_list:
    graphs = k.endswith("g"); k = int(k.rstrip("g"))
    pipe = batch.FramePipeline(frames.shape[1:], ks, sigma=10, n_grid=21, device=dev, streams=k, graphs=graphs)
    for rep in range(2):
        for i in range(min(8, nf)): pipe.submit(frames[i])
        pipe.join(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        res = None
        for i in range(nf): res = pipe.submit(frames[i])
        pipe.join()
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        chk = float(res["u"].double().abs().sum()), float(res["corrected"].abs().sum())
        captured = sum(sl is not None for sl in pipe._slots)
        if ref is None: ref = chk
        print(f"streams {k} graphs {captured}: {nf / ms * 1e3:7.1f} frames/s ({ms / nf:.3f} ms per frame on the GPU clock; host enqueue {(t1 - t0) / nf * 1e3:.3f} ms per frame, "
              f"drain {(t2 - t1) * 1e3:.1f} ms)  same={chk == ref}", flush=True)
    del pipe, res
    engine.release_workspaces()
