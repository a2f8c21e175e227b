"""Frame-batch driver (config 4 in miniature): device-resident chain == the NumPy-API chain == oracle."""
import numpy as np
import pytest

import oracle
from parity import DISP_TOL
from pygpa_b200 import batch, synth
from pygpa_b200 import geometric_phase_analysis as GPA

pytestmark = pytest.mark.gpu


def test_shard_frames():
    assert [len(batch.shard_frames(512, 8, r)) for r in range(8)] == [64] * 8
    assert sorted(i for r in range(3) for i in batch.shard_frames(10, 3, r)) == list(range(10))


def test_frame_series_matches_single_frame_api_and_oracle():
    shape = (128, 160)
    ks = synth.primary_ks(0.1, 7.0, 3)
    base = synth.smooth_random_field(shape, 0.05, seed=3)
    frames = []
    for t in range(3):
        u = base * (1 + 0.2 * np.sin(2 * np.pi * t / 3))
        frames.append(synth.lattice_image(shape, ks, u, noise=0.1, seed=100 + t) * (1 + 0.1 * t))
    res = batch.process_frames(frames, ks)
    assert sorted(res) == [0, 1, 2]
    for t, frame in enumerate(frames):
        u_api = GPA.extract_displacement_field(frame, ks)
        assert np.abs(res[t]["u"] - u_api).max() < 1e-9
        assert res[t]["corrected"].shape == shape
        # the ORIGINAL float64 frame is resampled, exactly as undistort_image(frame, -u) does
        assert np.abs(res[t]["corrected"] - GPA.undistort_image(frame, -u_api)).max() < 1e-9

    # against the oracle: EVERY pixel within 1e-3 px, except around pixels where the sweep legitimately picked another
    # candidate (near-tie, oracle top-2 gap < 1e-5): those carry a different lock-in phase, visible in u at that pixel
    import scipy.ndimage as ndi

    def sweep(im, s_, kx, ky, kw, kstep):
        return oracle.wfr_sweep(im, s_, kx, ky, kw, kstep, want_grad=False, return_diag=True)
    u_ref, gs_ref = oracle.extract_displacement_field(frames[1], ks, return_gs=True, sweep=sweep)
    _u, gs = GPA.extract_displacement_field(frames[1], ks, return_gs=True)
    flips = np.zeros(shape, dtype=bool)
    for g_, r_ in zip(gs, gs_ref):
        differs = ~np.all(g_['w'] == r_['w'], axis=0)
        gap = (r_['amp1'] - r_['amp2']) / r_['amp1']
        assert np.all(gap[differs] < 1e-5)
        flips |= differs
    assert flips.mean() < 1e-3
    near = ndi.binary_dilation(flips, iterations=2)
    assert np.abs(res[1]["u"] - u_ref).max(axis=0)[~near].max() < DISP_TOL


@pytest.mark.parametrize("streams,graphs", [(2, False), (1, True), (3, True)])
def test_stream_pool_and_graph_replay_equal_the_plain_pipeline(streams, graphs):
    """FramePipeline.submit spreads frames over CUDA streams (each with its own scratch) and can replay the per-frame chain
    as one CUDA graph per stream: both must return exactly what the plain call returns, frame after frame."""
    import torch
    from pygpa_b200 import engine
    dev = engine.require_cuda()
    shape = (128, 256)
    ks = synth.primary_ks(0.1, 7.0, 3)
    base = synth.smooth_random_field(shape, 0.05, seed=5)
    frames = [torch.from_numpy(synth.lattice_image(shape, ks, base * (1 + 0.1 * t), noise=0.1, seed=200 + t)).to(dev)
              for t in range(7)]
    plain = batch.FramePipeline(shape, ks, device=dev)
    want = [{k: v.clone() for k, v in plain(f).items()} for f in frames]
    pool = batch.FramePipeline(shape, ks, device=dev, streams=streams, graphs=graphs)
    for t, f in enumerate(frames):
        res = pool.submit(f)
        pool.join()
        torch.cuda.synchronize()
        for k in ("u", "corrected"):
            assert torch.equal(res[k], want[t][k]), (t, k)
    if graphs:
        assert all(slot is not None for slot in pool._slots[:min(streams, len(frames) // 2)])
    # several frames in flight at once, joined at the end
    outs = [{k: v.clone() for k, v in pool.submit(f).items()} if graphs else pool.submit(f) for f in frames[:streams]]
    pool.join()
    torch.cuda.synchronize()
    if not graphs:
        for t, res in enumerate(outs):
            assert torch.equal(res["u"], want[t]["u"])
