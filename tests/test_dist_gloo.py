"""Host-side logic of the multi-GPU k-grid sharding, exercised on CPU with the gloo backend and
world_size = 2: unit sharding, packed-key MAX merge with first-wins ties, exact SUM merge of the
payload.  The 'local sweep' of each rank is the oracle's amplitude for its share of planes."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from pygpa_b200 import dist as gdist
from pygpa_b200 import synth


def test_shard_units_balanced_and_complete():
    for world in (1, 2, 3, 4, 8):
        seen = np.zeros((3, 41), dtype=int)
        sizes = []
        for rank in range(world):
            ranges = gdist.shard_units(3, 41, world, rank)
            assert len(ranges) == 3
            sizes.append(sum(hi - lo for lo, hi in ranges))
            for p, (lo, hi) in enumerate(ranges):
                seen[p, lo:hi] += 1
        assert (seen == 1).all()
        assert max(sizes) - min(sizes) <= 1
    assert max(sum(hi - lo for lo, hi in gdist.shard_units(3, 41, 8, r)) for r in range(8)) == 16   # 96 % balance


def test_interleaved_shares_are_balanced_and_complete():
    for world in (2, 3, 4, 8):
        seen = np.zeros((3, 41), dtype=int)
        sizes = []
        for rank in range(world):
            total = 0
            for p, (lo, hi, step) in enumerate(gdist.shard_units_interleaved(3, 41, world, rank)):
                planes = list(range(lo, hi, step))
                seen[p, planes] += 1
                total += len(planes)
            sizes.append(total)
        assert (seen == 1).all() and max(sizes) - min(sizes) <= 1


def test_unequal_plane_counts_use_every_plane():
    """peaks with different np.arange lengths must each be swept over their OWN plane count."""
    for world in (1, 2, 3, 8):
        seen = [np.zeros(c, dtype=int) for c in (6, 7, 6)]
        for rank in range(world):
            for p, (lo, hi, st) in enumerate(gdist.shard_units_interleaved(3, (6, 7, 6), world, rank)):
                seen[p][lo:hi:st] += 1
            for p, (lo, hi) in enumerate(gdist.shard_units(3, (6, 7, 6), world, rank)):
                seen[p][lo:hi] += 1
        assert all((s == 2).all() for s in seen)


def test_pack_unpack_and_tie_break():
    amp2 = torch.tensor([0.0, 1.5, 1.5, 3.0e-20])
    idx = torch.tensor([7, 3, 2, 1680])
    key = gdist.pack_key(amp2, idx)
    a, i = gdist.unpack_key(key)
    assert torch.equal(a, amp2) and i.tolist() == [-1, 3, 2, 1680]
    assert key[2] > key[1]            # equal amplitude: the LOWER flat index wins the max
    assert key[1] > key[3] > key[0]   # amplitude ordering is preserved by the bit pattern


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, img, sigma, kx, ky, kw, kstep, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wxs, wys = oracle.candidate_axes(kx, ky, kw, kstep)
        ny = len(wys)
        (lo, hi), = gdist.shard_units(1, ny, world, rank)
        n, m = img.shape
        key = torch.zeros((n, m), dtype=torch.int64)
        lock = torch.zeros((n, m), dtype=torch.complex64)
        cands = {}
        # this rank's share of candidates, visited in a scrambled order: the merge must not care
        order = [(ix, iy) for iy in range(lo, hi) for ix in range(len(wxs))][::-1]
        for ix, iy in order:
            sf = oracle.lockin_fixed(img, (wxs[ix], wys[iy]), sigma)
            amp2 = torch.from_numpy((np.abs(sf) ** 2).astype(np.float32))
            flat = torch.full((n, m), ix * ny + iy, dtype=torch.int64)
            cand = gdist.pack_key(amp2, flat)
            cand = torch.where(amp2 > 0, cand, torch.zeros_like(cand))
            key = torch.maximum(key, cand)
            cands[ix * ny + iy] = sf
        gdist.merge_keys(key)
        _amp, idx = gdist.unpack_key(key)
        for flat, sf in cands.items():          # "finalize": only the winners this rank owns
            mine = idx == flat
            lock[mine] = torch.from_numpy(sf.astype(np.complex64))[mine]
        gdist.merge_payload([lock])
        if rank == 0:
            np.savez(os.path.join(out_dir, "merged.npz"), kidx=idx.numpy(), lockin=lock.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_merge_reproduces_single_process_argmax(tmp_path):
    shape = (40, 36)
    ks = synth.primary_ks(0.12, 7.0, 3)
    img = synth.lattice_image(shape, ks, synth.gaussian_bump(shape) * 0.5, noise=0.2, seed=4)
    img -= img.mean()
    kw, kstep = synth.sweep_params(ks, 5)
    sigma, k = 3, ks[0]
    mp.spawn(_worker, args=(2, _free_port(), img, sigma, k[0], k[1], kw, kstep, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "merged.npz")
    ref = oracle.wfr_sweep(img, sigma, k[0], k[1], kw, kstep, return_diag=True)
    # amplitudes were rounded to float32 before packing: only float32-level near-ties may differ
    same = got["kidx"] == ref["kidx"]
    gap = (ref["amp1"] - ref["amp2"]) / ref["amp1"]
    assert np.all(gap[~same] < 1e-6)
    assert same.mean() > 0.999
    wxs, wys = ref["wxs"], ref["wys"]
    x = np.arange(shape[0])[:, None]
    y = np.arange(shape[1])[None, :]
    ix, iy = got["kidx"] // len(wys), got["kidx"] % len(wys)
    unrot = ref["lockin"] * np.exp(2j * np.pi * ((wxs[ix] - k[0]) * x + (wys[iy] - k[1]) * y))
    assert np.abs(got["lockin"] - unrot)[same].max() < 1e-6 * np.abs(unrot).max() + 1e-7
