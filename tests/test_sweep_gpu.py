"""Parity of the CUDA lock-in / adaptive sweep (K1) with the oracle and the reference-made
golden fixtures.  Needs a B200; everything goes through the C ABI (ctypes)."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden
from parity import NEAR_TIE, check_sweep, wrap
from pygpa_b200 import cuGPA, engine, synth
from pygpa_b200 import geometric_phase_analysis as GPA

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def noisy_case():
    shape = (160, 128)
    ks = synth.primary_ks(0.1, 7.0, 3)
    u = synth.smooth_random_field(shape, 0.1, 5)
    img = synth.lattice_image(shape, ks, u, noise=0.3, seed=6)
    img -= img.mean()
    kw, kstep = synth.sweep_params(ks, 7)
    return dict(img=img, ks=ks, sigma=5, kw=kw, kstep=kstep)


def test_golden_fixture_all_peaks():
    g = load_golden("sweep_64x48.npz")
    img, ks = g["in_image"], g["in_ks"]
    sigma, kw, kstep = int(g["in_sigma"]), float(g["in_kw"]), float(g["in_kstep"])
    for i, k in enumerate(ks):
        diag = oracle.wfr_sweep(img, sigma, k[0], k[1], kw, kstep, return_diag=True)
        ref = dict(lockin=g["out_lockin"][i], w=g["out_w"][i], grad=g["out_grad"][i],
                   amp1=diag["amp1"], amp2=diag["amp2"])
        got = cuGPA.wfr2_grad_opt(img, sigma, k[0], k[1], kw, kstep)
        assert got["lockin"].dtype == np.complex128 and got["lockin"].shape == img.shape
        assert got["w"].shape == (2,) + img.shape and got["grad"].shape == img.shape + (2,)
        check_sweep(got, ref)
    got = GPA.optwfr2(img, sigma, ks[0][0], ks[0][1], kw, kstep)
    ref = dict(lockin=g["out_optwfr2_lockin"], w=g["out_optwfr2_w"], amp1=diag["amp1"] * 0 + 1, amp2=diag["amp2"] * 0)
    assert set(got) == {"lockin", "w"}
    diag0 = oracle.wfr_sweep(img, sigma, ks[0][0], ks[0][1], kw, kstep, return_diag=True, want_grad=False)
    ref["amp1"], ref["amp2"] = diag0["amp1"], diag0["amp2"]
    check_sweep(got, ref, check_grad=False)


@pytest.mark.parametrize("peak", [0, 1, 2])
def test_oracle_parity_nonsquare(noisy_case, peak):
    c = noisy_case
    k = c["ks"][peak]
    ref = oracle.wfr_sweep(c["img"], c["sigma"], k[0], k[1], c["kw"], c["kstep"], return_diag=True)
    got = cuGPA.wfr2_grad_opt(c["img"], c["sigma"], k[0], k[1], c["kw"], c["kstep"])
    stats = check_sweep(got, ref)
    assert stats["frac_mismatch"] < 1e-3


def test_forward_difference_gradient_mode(noisy_case):
    c = noisy_case
    k = c["ks"][0]
    ref = oracle.wfr_sweep(c["img"], c["sigma"], k[0], k[1], c["kw"], c["kstep"], grad_mode='diff', return_diag=True)
    got = cuGPA.wfr2_grad_opt(c["img"], c["sigma"], k[0], k[1], c["kw"], c["kstep"], grad='diff')
    assert np.isnan(got["grad"][-1, :, 0]).all() and np.isnan(got["grad"][:, -1, 1]).all()
    check_sweep(got, ref)
    only = cuGPA.wfr2_only_grad(c["img"], c["sigma"], tuple(k), c["kw"], c["kstep"], grad='diff')
    assert np.array_equal(np.isnan(only), np.isnan(got["grad"]))


def _grad_close(a, b, tol=1e-3):
    """phase gradients are defined modulo pi (geometric_phase_analysis.py:812); NaNs must coincide"""
    assert np.array_equal(np.isnan(a), np.isnan(b))
    d = np.abs(a - b)
    d = np.minimum(d, np.abs(d - np.pi))
    return np.nanmax(d) < tol


def test_gradient_callable_and_wfr2_grad_modes(noisy_case):
    """grad=<callable> (cuGPA.py:66-67) and geometric_phase_analysis.wfr2_grad(grad='diff' | callable) (:722-760) run
    unfused — one fixed lock-in per distinct winning candidate, the gradient function on the host — and must give what
    the reference's per-candidate loop gives wherever the same candidate wins."""
    c = noisy_case
    k = c["ks"][0]
    args = (c["img"], c["sigma"], k[0], k[1], c["kw"], c["kstep"])
    fused = cuGPA.wfr2_grad_opt(*args)
    unfused = cuGPA.wfr2_grad_opt(*args, grad=np.gradient)              # np.gradient returns the pair cp.gradient does
    assert np.array_equal(fused["lockin"], unfused["lockin"]) and np.array_equal(fused["w"], unfused["w"])
    assert _grad_close(fused["grad"], unfused["grad"])
    assert _grad_close(cuGPA.wfr2_only_grad(c["img"], c["sigma"], tuple(k), c["kw"], c["kstep"], grad=np.gradient), fused["grad"])
    assert set(cuGPA.wfr2_grad_single(*args, grad=np.gradient)) == {"lockin", "grad"}

    def one_sided(phase):
        return np.stack([np.diff(phase, axis=0, prepend=phase[:1]), np.diff(phase, axis=1, prepend=phase[:, :1])], axis=-1)
    for grad in (None, 'diff', one_sided):
        ref = oracle.wfr2_grad(*args, grad=grad)
        got = GPA.wfr2_grad(*args, grad=grad)
        same = np.all(got["w"] == ref["w"], axis=0)
        assert same.mean() > 0.999
        assert got["grad"].shape == c["img"].shape + (2,)
        assert _grad_close(np.where(same[..., None], got["grad"], 0), np.where(same[..., None], ref["grad"], 0))
        assert np.abs(np.angle(got["lockin"][same] * np.conj(ref["lockin"][same]))).max() < 1e-3
    with pytest.raises(ValueError):
        GPA.wfr2_grad(*args, grad='central')


def test_config2_quarter_size_near_tie_accounting():
    """C2 lattice at 256^2 with the full 21x21 grid: every k mismatch must sit on a near-tie."""
    cfg = synth.make_config('C2', size=256)
    k = cfg["ks"][0]
    ref = oracle.wfr_sweep(cfg["image"], cfg["sigma"], k[0], k[1], cfg["kw"], cfg["kstep"], return_diag=True)
    assert len(ref["wxs"]) == 21 and len(ref["wys"]) == 21
    got = cuGPA.wfr2_grad_opt(cfg["image"], cfg["sigma"], k[0], k[1], cfg["kw"], cfg["kstep"])
    stats = check_sweep(got, ref)
    assert stats["frac_mismatch"] < 1e-3 and stats["phase_err"] < 1e-3
    near = ((ref["amp1"] - ref["amp2"]) / ref["amp1"] < NEAR_TIE).mean()
    assert stats["frac_mismatch"] <= near


@pytest.mark.parametrize("shape,sigma,stride", [((160, 128), 5, 2), ((256, 192), 10, 4), ((256, 320), 22, 8)])
def test_multirate_and_direct_forms_agree(shape, sigma, stride):
    """The arg-max runs in the multirate form (decimate by S, interpolate) when frame and sigma allow;
    it must select the same candidates as the direct form and the oracle, near-ties excepted, for
    every supported stride."""
    from pygpa_b200 import _taps
    mr = _taps.multirate_taps(shape[0], shape[1], float(sigma))
    assert mr is not None and mr["S"] == stride
    ks = synth.primary_ks(0.5 / sigma, 7.0, 3)
    u = synth.smooth_random_field(shape, 0.1, seed=sigma)
    img = synth.lattice_image(shape, ks, u, noise=0.3, seed=sigma + 1)
    img -= img.mean()
    kw, kstep = synth.sweep_params(ks, 9)
    k = ks[1]
    dev = engine.require_cuda()
    d_img = engine.image_to_device(img, dev)
    wxs, wys = engine.grid_axes(k[0], k[1], kw, kstep)
    res = {}
    for method in ("direct", "multirate"):
        plan = engine.SweepPlan(d_img.shape, wxs, wys, sigma, device=dev, method=method)
        assert (plan.mr is not None) == (method == "multirate")
        res[method] = plan.run(d_img, k, want_w=True, out_f64=True)
    ref = oracle.wfr_sweep(img, sigma, k[0], k[1], kw, kstep, return_diag=True)
    gap = (ref["amp1"] - ref["amp2"]) / ref["amp1"]
    for method, r in res.items():
        got = {key: r[key].cpu().numpy() for key in ("lockin", "w", "grad")}
        stats = check_sweep(got, ref)
        assert stats["frac_mismatch"] < 2e-3, method
    differ = (res["direct"]["kidx"] != res["multirate"]["kidx"]).cpu().numpy()
    assert np.all(gap[differ] < NEAR_TIE)
    # chunked multirate run (2 planes resident at a time) is bit-identical to the unchunked one
    chunked = engine.SweepPlan(d_img.shape, wxs, wys, sigma, device=dev, method="multirate", planes_in_flight=2).run(d_img, k)
    assert torch.equal(chunked["key"], res["multirate"]["key"])


def test_tma_and_cp_async_tile_loads_agree():
    """The coarse tiles of k_mr_interp arrive by TMA box loads (interior tiles) or per-element cp.async gathers (tiles whose
    window wraps around the frame, or TMA off): the keys must be bit-identical."""
    cfg = synth.make_config('C2', size=512, n_grid=13)
    dev = engine.require_cuda()
    img = engine.image_to_device(cfg["image"], dev)
    k = cfg["ks"][1]
    wxs, wys = engine.grid_axes(k[0], k[1], cfg["kw"], cfg["kstep"])
    plan = engine.SweepPlan(img.shape, wxs, wys, cfg["sigma"], device=dev, method="multirate")
    a = plan.run(img, k)
    engine.set_tma(False)
    try:
        b = plan.run(img, k)
    finally:
        engine.set_tma(True)
    assert torch.equal(a["key"], b["key"])
    assert torch.equal(torch.view_as_real(a["lockin"]), torch.view_as_real(b["lockin"]))


def test_pruning_is_exact():
    """Branch-and-bound pruning of the multirate arg-max must not change a single key bit, on a
    structured frame (where it removes most candidates) and on pure noise (where it removes few)."""
    dev = engine.require_cuda()
    cfg = synth.make_config('C2', size=512)
    noise = np.random.default_rng(0).normal(size=(512, 512))
    for img in (cfg["image"], noise):
        d_img = engine.image_to_device(img, dev)
        for k in cfg["ks"][:2]:
            wxs, wys = engine.grid_axes(k[0], k[1], cfg["kw"], cfg["kstep"])
            plan = engine.SweepPlan(d_img.shape, wxs, wys, cfg["sigma"], device=dev, method="multirate")
            try:
                engine.set_pruning(False)
                ref = plan.run(d_img, k)
                engine.set_pruning(True)
                got = plan.run(d_img, k)
                again = plan.run(d_img, k)
            finally:
                engine.set_pruning(True)
            assert torch.equal(ref["key"], got["key"]) and torch.equal(got["key"], again["key"])
            assert torch.equal(torch.view_as_real(ref["lockin"]), torch.view_as_real(got["lockin"]))
            assert torch.equal(ref["grad"], got["grad"])


def test_many_rows_uses_wide_index_packing():
    """More than 256 candidate rows per plane switches the packed winner index from 8 to 16 bits."""
    shape = (96, 128)
    rng = np.random.default_rng(2)
    img = rng.normal(size=shape)
    wxs = np.linspace(0.05, 0.15, 300)
    wys = np.array([0.02, 0.03])
    dev = engine.require_cuda()
    d_img = engine.image_to_device(img, dev)
    res = {m: engine.SweepPlan(shape, wxs, wys, 10, device=dev, method=m).run(d_img, (0.1, 0.025)) for m in ("direct", "multirate")}
    same = (res["direct"]["kidx"] == res["multirate"]["kidx"]).float().mean().item()
    assert same > 0.995
    assert res["multirate"]["kidx"].max().item() >= 256 * 2       # high row indices survive the packing
    # spot check against the oracle on a few candidates around the winner of one pixel
    k = int(res["direct"]["kidx"][40, 50].item())
    ix, iy = divmod(k, 2)
    amp = [abs(oracle.lockin_fixed(img, (wxs[j], wys[iy]), 10)[40, 50]) for j in (max(ix - 1, 0), ix, min(ix + 1, 299))]
    assert amp[1] >= max(amp) * (1 - 1e-5)


def test_large_nonsquare_frame_smoke():
    """4096 x 2048 frame (larger than any other test, non-square): both forms run and agree."""
    shape = (4096, 2048)
    ks = synth.primary_ks(0.05, 7.0, 3)
    img = synth.lattice_image(shape, ks, noise=0.3, seed=9).astype(np.float32)
    kw, kstep = synth.sweep_params(ks, 3)
    k = ks[0]
    dev = engine.require_cuda()
    d_img = engine.image_to_device(img, dev)
    wxs, wys = engine.grid_axes(k[0], k[1], kw, kstep)
    a = engine.SweepPlan(shape, wxs, wys, 10, device=dev, method="direct").run(d_img, k)
    b = engine.SweepPlan(shape, wxs, wys, 10, device=dev, method="multirate").run(d_img, k)
    assert (a["kidx"] == b["kidx"]).float().mean().item() > 0.999
    assert a["kidx"].min().item() >= 0 and a["kidx"].max().item() < 9
    amax = a["lockin"].abs().max().item()
    agree = a["kidx"] == b["kidx"]
    assert (a["lockin"] - b["lockin"]).abs()[agree].max().item() < 1e-4 * amax


def test_variants_and_single(noisy_case):
    c = noisy_case
    k = c["ks"][1]
    full = cuGPA.wfr2_grad_opt(c["img"], c["sigma"], k[0], k[1], c["kw"], c["kstep"])
    single = cuGPA.wfr2_grad_single(c["img"], c["sigma"], k[0], k[1], c["kw"], c["kstep"])
    assert set(single) == {"lockin", "grad"}
    assert np.array_equal(single["lockin"], full["lockin"]) and np.array_equal(single["grad"], full["grad"])
    assert np.array_equal(cuGPA.wfr2_only_lockin(c["img"], c["sigma"], tuple(k), c["kw"], c["kstep"]), full["lockin"])
    assert np.array_equal(GPA.wfr2_only_lockin(c["img"], c["sigma"], k[0], k[1], c["kw"], c["kstep"]), full["lockin"])
    r = GPA.wfr(c["img"], c["sigma"], k[0], k[1], c["kw"], c["kstep"])
    assert np.array_equal(r["wx"], full["w"][0]) and np.allclose(r["r"], np.abs(full["lockin"]))


def test_fixed_reference_lockin(noisy_case):
    c = noisy_case
    for k in c["ks"]:
        ref = oracle.lockin_fixed(c["img"], k, 6)
        for got in (GPA.optGPA(c["img"], k, 6), GPA.GPA(c["img"], k[0], k[1], 6), np.asarray(cuGPA.cuGPA(c["img"], k, 6))):
            assert got.dtype == np.complex128
            assert np.abs(got - ref).max() < 1e-4 * np.abs(ref).max()
    stack = GPA.vecGPA(c["img"], c["ks"], 6)
    assert stack.shape == (3,) + c["img"].shape and stack.dtype == np.complex128
    for i, k in enumerate(c["ks"]):         # values: every plane is that k-vector's lock-in (geometric_phase_analysis.py:79-89)
        ref = oracle.lockin_fixed(c["img"], k, 6)
        assert np.abs(stack[i] - ref).max() < 1e-4 * np.abs(ref).max()
        assert np.array_equal(stack[i], GPA.optGPA(c["img"], k, 6))
    g = load_golden("fixed_64x64.npz")
    got = np.stack([GPA.optGPA(g["in_image"], k, int(g["in_sigma"])) for k in g["in_ks"]])
    assert np.abs(got - g["out_lockin"]).max() < 1e-4 * np.abs(g["out_lockin"]).max()
    strong = np.abs(g["out_lockin"]) > 0.05 * np.abs(g["out_lockin"]).max()
    assert np.abs(np.angle(got * np.conj(g["out_lockin"])))[strong].max() < 1e-3


def test_wfr3_candidate_list():
    g = load_golden("wfr3_64x48.npz")
    got = GPA.wfr3(g["in_image"], int(g["in_sigma"]), g["in_klist"], g["in_kref"])
    diag = oracle.wfr_sweep_klist(g["in_image"], int(g["in_sigma"]), g["in_klist"], g["in_kref"],
                                  want_grad=False, return_diag=True)
    ref = dict(lockin=g["out_lockin"], w=g["out_w"], amp1=diag["amp1"], amp2=diag["amp2"])
    check_sweep(got, ref, check_grad=False)


def _check_wfr4(got, image, sigma, klist, kref, dk, ref_lockin, ref_w):
    """wfr4 is path dependent per pixel: a pixel whose amplitude decisions all had a relative margin
    above the near-tie budget must follow the reference's path exactly."""
    diag = oracle.wfr4(image, sigma, klist, kref, dk, return_diag=True)
    assert np.array_equal(diag["w"], ref_w)
    same = np.all(got["w"] == ref_w, axis=0)
    clear = diag["margin"] > NEAR_TIE
    assert not (~same & clear).any(), f"{(~same & clear).sum()} pixels follow another path away from a near-tie"
    amax = np.abs(ref_lockin).max()
    assert np.abs(got["lockin"] - ref_lockin)[same].max() <= 1e-4 * amax
    m = same & (np.abs(ref_lockin) > 0.02 * amax)
    assert np.abs(np.angle(got["lockin"][m] * np.conj(ref_lockin[m]))).max() < 1e-3
    return same.mean()


def test_wfr4_neighbourhood_rule_matches_reference_fixture():
    g = load_golden("wfr4_64x48.npz")
    img, sigma, klist, kref, dk = g["in_image"], int(g["in_sigma"]), g["in_klist"], g["in_kref"], float(g["in_dk"])
    got = GPA.wfr4(img, sigma, klist, kref, dk)
    assert got["lockin"].dtype == np.complex128 and got["w"].shape == (2,) + img.shape
    assert _check_wfr4(got, img, sigma, klist, kref, dk, g["out_lockin"], g["out_w"]) > 0.99
    # reversed list: starts far from the peak, the neighbourhood rule decides which pixels ever move
    rev = klist[::-1].copy()
    got = GPA.wfr4(img, sigma, rev, kref, dk)
    assert _check_wfr4(got, img, sigma, rev, kref, dk, g["out_rev_lockin"], g["out_rev_w"]) > 0.99


def test_wfr4_chunked_planes_and_unreachable_candidates():
    """Workspace-limited plane chunks carry the (amplitude, held index) state through `key`; with a
    tiny dk no candidate but the first is ever reachable, so w stays klist[0] everywhere."""
    g = load_golden("wfr4_64x48.npz")
    img, sigma, klist, kref, dk = g["in_image"], int(g["in_sigma"]), g["in_klist"], g["in_kref"], float(g["in_dk"])
    dev = engine.require_cuda()
    d_img = engine.image_to_device(img, dev)
    full = engine.wfr4_sweep(d_img, sigma, klist, kref, dk)
    real_plan = engine._plan_planes

    def few_planes(n, m, n_rows, n_planes, rx, ry, device, planes_in_flight):
        return real_plan(n, m, n_rows, n_planes, rx, ry, device, 5)
    engine._plan_planes = few_planes
    try:
        engine.release_workspaces()
        part = engine.wfr4_sweep(d_img, sigma, klist, kref, dk)
    finally:
        engine._plan_planes = real_plan
        engine.release_workspaces()
    assert torch.equal(full["key"], part["key"]) and torch.equal(full["kidx"], part["kidx"])
    assert torch.equal(torch.view_as_real(full["lockin"]), torch.view_as_real(part["lockin"]))
    stuck = GPA.wfr4(img, sigma, klist, kref, 1e-9)
    assert np.all(stuck["w"][0] == klist[0, 0]) and np.all(stuck["w"][1] == klist[0, 1])
    ref = oracle.wfr4(img, sigma, klist, kref, 1e-9)
    assert np.abs(stuck["lockin"] - ref["lockin"]).max() <= 1e-4 * np.abs(ref["lockin"]).max()


def test_vec_variants_are_the_same_sweep(noisy_case):
    """wfr2_only_lockin_vec / wfr2_grad_vec (geometric_phase_analysis.py:705-719, 816-836) only batch
    the candidates differently (dask); their results are those of the plain functions."""
    c = noisy_case
    k = c["ks"][0]
    a = GPA.wfr2_grad_vec(c["img"], c["sigma"], k[0], k[1], c["kw"], c["kstep"])
    b = GPA.wfr2_grad_opt(c["img"], c["sigma"], k[0], k[1], c["kw"], c["kstep"])
    for key in ("lockin", "w", "grad"):
        assert np.array_equal(a[key], b[key])
    assert np.array_equal(GPA.wfr2_only_lockin_vec(c["img"], c["sigma"], k[0], k[1], c["kw"], c["kstep"]), b["lockin"])


def _run_device(img, plan, kref, **kw):
    return plan.run(img, kref, **kw)


def _same(a, b):
    return torch.equal(torch.view_as_real(a) if a.is_complex() else a, torch.view_as_real(b) if b.is_complex() else b)


@pytest.mark.parametrize("method", ["direct", "multirate"])
def test_plane_chunking_and_range_merge_are_bit_exact(noisy_case, method):
    """Resident-plane chunking and splitting the plane range over several calls (the multi-GPU
    k-grid sharding) must not change a single bit of the arg-max; with per-plan workspaces the
    finalized payload is bit-identical too."""
    c = noisy_case
    dev = engine.require_cuda()
    img = engine.image_to_device(c["img"], dev)
    k = c["ks"][2]
    wxs, wys = engine.grid_axes(k[0], k[1], c["kw"], c["kstep"])
    full = engine.SweepPlan(img.shape, wxs, wys, c["sigma"], device=dev, method=method).run(img, k, want_w=True)
    chunked = engine.SweepPlan(img.shape, wxs, wys, c["sigma"], planes_in_flight=3, device=dev, method=method).run(img, k, want_w=True)
    assert torch.equal(full["key"], chunked["key"]) and torch.equal(full["kidx"], chunked["kidx"])
    if method == "direct":
        for key in ("lockin", "grad", "w"):
            assert _same(full[key], chunked[key]), key
    else:   # chunked multirate falls back to the direct-form finalize: same winners, values to rounding
        assert _same(full["w"], chunked["w"])
        amax = full["lockin"].abs().max().item()
        assert (full["lockin"] - chunked["lockin"]).abs().max().item() < 1e-4 * amax
    # two "ranks": planes [0,3) and [3,ny), each with its own plan + private workspace
    ranges = ((0, 3), (3, len(wys)))
    plans = [engine.SweepPlan(img.shape, wxs, wys, c["sigma"], device=dev, method=method, private_ws=True) for _ in ranges]
    keys = []
    for plan, (lo, hi) in zip(plans, ranges):
        kk = torch.zeros(img.shape, dtype=torch.int64, device=dev)
        plan.argmax(img, kk, lo, hi)
        keys.append(kk)
    # keys are unsigned 64-bit with the top bit clear (|sf|^2 >= 0), so a signed max is the same
    merged = torch.maximum(keys[0], keys[1])
    assert torch.equal(merged, full["key"])
    parts = [plan.finalize(img, merged, k, plane_begin=lo, plane_end=hi, want_w=True, planes_valid=True)
             for plan, (lo, hi) in zip(plans, ranges)]
    assert _same(parts[0]["lockin"] + parts[1]["lockin"], full["lockin"])
    assert _same(parts[0]["grad"] + parts[1]["grad"], full["grad"])
    assert _same(parts[0]["w"] + parts[1]["w"], full["w"])


def test_zero_image_keeps_zeros():
    """|sf| never exceeds |0|: the reference leaves lockin = 0, w = 0, grad = 0."""
    img = np.zeros((40, 36))
    got = cuGPA.wfr2_grad_opt(img, 3, 0.1, 0.02, 0.04, 0.013)
    assert not got["lockin"].any() and not got["w"].any() and not got["grad"].any()
    dev = engine.require_cuda()
    wxs, wys = engine.grid_axes(0.1, 0.02, 0.04, 0.013)
    res = engine.SweepPlan(img.shape, wxs, wys, 3, device=dev).run(engine.image_to_device(img, dev), (0.1, 0.02))
    assert (res["kidx"] == -1).all()


def test_small_frame_filter_covers_whole_circle():
    """Frames narrower than the 4.5 sigma window: the filter radius is clamped to (n-1)//2.
    Odd axes then carry the reference's whole circular kernel (exact); on even axes the single
    antipodal tap (d = n/2) is dropped, a documented deviation of order g(n/2)/g(0)."""
    rng = np.random.default_rng(3)
    img = rng.normal(size=(25, 31))
    ref = oracle.wfr_sweep(img, 4, 0.2, 0.1, 0.05, 0.02, return_diag=True)
    got = cuGPA.wfr2_grad_opt(img, 4, 0.2, 0.1, 0.05, 0.02)
    check_sweep(got, ref)
    img = rng.normal(size=(24, 31))
    ref = oracle.wfr_sweep(img, 4, 0.2, 0.1, 0.05, 0.02, return_diag=True)
    got = cuGPA.wfr2_grad_opt(img, 4, 0.2, 0.1, 0.05, 0.02)
    same = np.all(got["w"] == ref["w"], axis=0)
    assert same.mean() > 0.97
    assert np.abs(got["lockin"] - ref["lockin"])[same].max() < 3e-2 * np.abs(ref["lockin"]).max()


def test_full_size_properties_config3():
    """BASELINE config 3 size (2048^2, 41x41 candidates), one peak: size-independent properties.
    (a) exact linearity: scaling the image by 2 leaves every k-index unchanged and doubles the
        lock-in bit-exactly (power-of-two scaling is exact in fp32);
    (b) the winner really is the arg-max: an independent float64 evaluation of the Gabor sum at
        sampled pixels reproduces the phase and is not beaten by the four neighbouring candidates;
    (c) circular-shift covariance of the selected k-vector (away from the frame seam)."""
    cfg = synth.make_config('C3')
    dev = engine.require_cuda()
    img64 = cfg["image"]
    img = engine.image_to_device(img64, dev)
    k = cfg["ks"][0]
    sigma = cfg["sigma"]
    wxs, wys = engine.grid_axes(k[0], k[1], cfg["kw"], cfg["kstep"])
    assert len(wxs) == 41 and len(wys) == 41
    plan = engine.SweepPlan(img.shape, wxs, wys, sigma, device=dev)
    a = plan.run(img, k)
    b = plan.run(img * 2, k)
    assert torch.equal(a["kidx"], b["kidx"])
    assert torch.equal(torch.view_as_real(a["lockin"]) * 2, torch.view_as_real(b["lockin"]))
    assert torch.equal(a["grad"], b["grad"])

    kidx = a["kidx"].cpu().numpy()
    lock = a["lockin"].cpu().numpy()
    assert kidx.min() >= 0 and kidx.max() < 41 * 41
    rng = np.random.default_rng(0)
    n = img64.shape[0]
    r = 6 * sigma
    d = np.arange(-r, r + 1)
    gw = np.exp(-d ** 2 / (2.0 * sigma ** 2)) / (sigma * np.sqrt(2 * np.pi))

    def gabor(x, y, wx, wy):
        xs, ys = (x + d) % n, (y + d) % n
        patch = img64[np.ix_(xs, ys)]
        return (gw[:, None] * gw[None, :] * patch * np.exp(2j * np.pi * (wx * xs[:, None] + wy * ys[None, :]))).sum()

    for _ in range(24):
        x, y = rng.integers(0, n, size=2)
        ix, iy = divmod(int(kidx[x, y]), 41)
        s = gabor(x, y, wxs[ix], wys[iy])
        rot = np.exp(-2j * np.pi * ((wxs[ix] - k[0]) * x + (wys[iy] - k[1]) * y))
        assert abs(np.angle(lock[x, y] * np.conj(s * rot))) < 1e-3
        assert abs(abs(lock[x, y]) - abs(s)) < 1e-4 * abs(s) + 1e-6
        for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            jx, jy = ix + dx, iy + dy
            if 0 <= jx < 41 and 0 <= jy < 41:
                assert abs(gabor(x, y, wxs[jx], wys[jy])) <= abs(s) * (1 + NEAR_TIE)

    # the carrier is evaluated at the true pixel index, so it is NOT periodic with the frame: the
    # covariance holds away (> R) from the frame seam and from where the shift moves that seam
    shift = (37, 1001)
    c = plan.run(engine.image_to_device(np.roll(img64, shift, axis=(0, 1)), dev), k)
    rolled = np.roll(kidx, shift, axis=(0, 1))
    idx = np.arange(n)
    far = [np.minimum.reduce([np.minimum((idx - s0) % n, (s0 - idx) % n) for s0 in (0, sh)]) > plan.rx + 1
           for sh in shift]
    m = far[0][:, None] & far[1][None, :]
    assert m.mean() > 0.6
    assert (c["kidx"].cpu().numpy() == rolled)[m].mean() > 0.999


def test_config5_size_8192_spot_check():
    """BASELINE config 5 size (8192^2, 41x41 candidates), one peak.  The frame is synthesised on the
    device (torch, float64: test plumbing only) because the NumPy generator needs ~40 s at this size; an
    independent float64 Gabor sum on patches copied back checks phase, amplitude and that no
    neighbouring candidate beats the winner.  fp32 carriers would lose 5e-4 rad at x ~ 8192: the
    kernels build them in fp64 with range reduction, which this test would catch."""
    dev = engine.require_cuda()
    n, sigma, ng = 8192, 10, 41
    ks = synth.primary_ks(0.05, 7.0, 3)
    kw, kstep = synth.sweep_params(ks, ng)
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.arange(n, dtype=torch.float64, device=dev)[:, None]
    y = torch.arange(n, dtype=torch.float64, device=dev)[None, :]
    ux = 6.0 * torch.sin(2 * np.pi * (1.3 * x + 0.4 * y) / n) * torch.cos(2 * np.pi * 0.9 * y / n)
    uy = 5.0 * torch.cos(2 * np.pi * (0.7 * x - 1.1 * y) / n)
    img64 = torch.zeros((n, n), dtype=torch.float64, device=dev)
    for k in ks:
        img64 += torch.cos(2 * np.pi * (k[0] * (x + ux) + k[1] * (y + uy)))
    img64 += 0.3 * torch.randn((n, n), dtype=torch.float64, device=dev, generator=g)
    img64 -= img64.mean()
    del ux, uy
    img = img64.float()
    k = ks[0]
    wxs, wys = engine.grid_axes(k[0], k[1], kw, kstep)
    assert len(wxs) == ng and len(wys) == ng
    plan = engine.SweepPlan((n, n), wxs, wys, sigma, device=dev)
    assert plan.mr is not None
    a = plan.run(img, k)
    kidx = a["kidx"]
    assert kidx.min().item() >= 0 and kidx.max().item() < ng * ng
    # the winners follow the local lattice: a smooth map, not the grid centre everywhere
    assert torch.unique(kidx).numel() > 50
    r = 6 * sigma
    d = np.arange(-r, r + 1)
    gw = np.exp(-d ** 2 / (2.0 * sigma ** 2)) / (sigma * np.sqrt(2 * np.pi))
    rng = np.random.default_rng(1)
    pts = [(5, 8190), (8191, 3), (4096, 4096)] + [tuple(rng.integers(0, n, size=2)) for _ in range(9)]
    for px, py in pts:
        xs, ys = (px + d) % n, (py + d) % n
        patch = img64[torch.as_tensor(xs, device=dev)][:, torch.as_tensor(ys, device=dev)].cpu().numpy()

        def gabor(wx, wy):
            return (gw[:, None] * gw[None, :] * patch * np.exp(2j * np.pi * (wx * xs[:, None] + wy * ys[None, :]))).sum()
        ix, iy = divmod(int(kidx[px, py].item()), ng)
        s = gabor(wxs[ix], wys[iy])
        rot = np.exp(-2j * np.pi * ((wxs[ix] - k[0]) * px + (wys[iy] - k[1]) * py))
        lock = complex(a["lockin"][px, py].item())
        assert abs(np.angle(lock * np.conj(s * rot))) < 1e-3
        assert abs(abs(lock) - abs(s)) < 1e-4 * abs(s) + 1e-6
        for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            jx, jy = ix + dx, iy + dy
            if 0 <= jx < ng and 0 <= jy < ng:
                assert abs(gabor(wxs[jx], wys[jy])) <= abs(s) * (1 + NEAR_TIE)
    del a, plan
    engine.release_workspaces()
    torch.cuda.empty_cache()


@pytest.mark.parametrize("shape,n_grid", [((256, 256), 21), ((192, 328), 13), ((520, 136), 9)])
def test_split_pass2_matches_single_stage(shape, n_grid):
    """Split pass 2 (anchor stage shared by a plane's candidates + coarse-rate stage per candidate)
    against every candidate's own full-rate pass 2: same winners except at near-ties, same winning
    amplitudes to 1e-5 — on the rows next to the frame edge too, where the carrier of the wrapped
    samples jumps — and both within the oracle's tolerances.  Shapes: coarse grid not a multiple of
    the 32-column / 128-row CTA tiles."""
    ks = synth.primary_ks(0.05, 7.0, 3)
    u = synth.smooth_random_field(shape, 0.25, seed=11)
    img = synth.lattice_image(shape, ks, u, noise=0.3, seed=12)
    img -= img.mean()
    kw, kstep = synth.sweep_params(ks, n_grid)
    dev = engine.require_cuda()
    d_img = engine.image_to_device(img, dev)
    for k in ks[:2]:
        wxs, wys = engine.grid_axes(k[0], k[1], kw, kstep)
        split = engine.SweepPlan(shape, wxs, wys, 10, device=dev, method="multirate")
        xonly = engine.SweepPlan(shape, wxs, wys, 10, device=dev, method="multirate", split_y=False)
        single = engine.SweepPlan(shape, wxs, wys, 10, device=dev, method="multirate-single")
        assert split.split is not None and split.split_y is not None and single.split is None and single.split_y is None
        assert xonly.split is not None and xonly.split_y is None
        a = split.run(d_img, k, want_w=True, out_f64=True)
        b = single.run(d_img, k, want_w=True, out_f64=True)
        c = xonly.run(d_img, k, want_w=True, out_f64=True)
        ref = oracle.wfr_sweep(img, 10, k[0], k[1], kw, kstep, return_diag=True)
        gap = (ref["amp1"] - ref["amp2"]) / ref["amp1"]
        for r in (a, b, c):
            check_sweep({key: r[key].cpu().numpy() for key in ("lockin", "w", "grad")}, ref)
        assert np.all(gap[(a["kidx"] != c["kidx"]).cpu().numpy()] < NEAR_TIE)
        differ = (a["kidx"] != b["kidx"]).cpu().numpy()
        assert np.all(gap[differ] < NEAR_TIE)
        amp_a = (a["key"] >> 32).to(torch.int32).view(torch.float32).cpu().numpy()
        amp_b = (b["key"] >> 32).to(torch.int32).view(torch.float32).cpu().numpy()
        rel = np.abs(amp_a - amp_b) / amp_b.max()
        assert rel.max() < 2e-5, f"winning |sf|^2 differs by {rel.max():.3g} (rows {np.argwhere(rel > 2e-5)[:4]})"
        # chunked split run is bit-identical to the unchunked one
        chunked = engine.SweepPlan(shape, wxs, wys, 10, device=dev, method="multirate", planes_in_flight=3).run(d_img, k)
        assert torch.equal(chunked["key"], a["key"])
