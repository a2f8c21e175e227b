"""Generate tests/golden/*.npz with the UNMODIFIED reference (run in the build container).

    python -m oracle.gen_golden

Every array under an ``out_`` key is the return value of a reference function
(pyGPA.geometric_phase_analysis / pyGPA.phase_unwrap imported from /root/reference);
the ``in_`` keys are the exact inputs, so the fixtures do not depend on the synthetic
generator staying bit-stable.  Fixtures are small (whole directory < 2 MB).
"""
import os
import numpy as np

from . import _refimport
from pygpa_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sorted_klist(centre, dk, half):
    """A (2 half + 1)^2 grid of spacing dk around `centre`, nearest first — what
    generate_klists(..., sort_list=True) (geometric_phase_analysis.py:865-889) feeds wfr4."""
    ax = dk * np.arange(-half, half + 1)
    kl = np.stack(np.meshgrid(ax, ax, indexing='ij'), axis=-1).reshape(-1, 2)
    order = np.argsort(np.linalg.norm(kl, axis=1), kind='stable')
    return np.asarray(centre) + kl[order]


def gen_wfr4(gpa, pu):
    """wfr4 (geometric_phase_analysis.py:839-862) on the frame of sweep_64x48.npz."""
    g = dict(np.load(os.path.join(OUT, "sweep_64x48.npz")))
    img, ks, sigma = g["in_image"], g["in_ks"], int(g["in_sigma"])
    dk = 0.008
    klist = sorted_klist(ks[1], dk, 3)
    r = gpa.wfr4(img, sigma, klist, ks[1], dk)
    r_far = gpa.wfr4(img, sigma, klist[::-1].copy(), ks[1], dk)     # starts far out: most pixels get stuck
    np.savez_compressed(os.path.join(OUT, "wfr4_64x48.npz"), in_image=img, in_sigma=sigma, in_klist=klist,
                        in_kref=ks[1], in_dk=dk, out_lockin=r['lockin'], out_w=r['w'],
                        out_rev_lockin=r_far['lockin'], out_rev_w=r_far['w'])


def gen_props(gpa, pu):
    """phasegradient2J / props_from_Jac (property_extract.py:69-101, 137-178) on the gradients of
    sweep_64x48.npz; `in_ks_ani` is a strained copy of the k-vectors so the isotropic re-referencing
    (calc_diff_from_isotropic) is not a no-op."""
    pe = _refimport.load_property_extract()
    g = dict(np.load(os.path.join(OUT, "sweep_64x48.npz")))
    ks, grads, w = g["in_ks"], g["out_grad"], np.abs(g["out_lockin"])
    ks_ani = ks @ np.array([[1.03, 0.02], [-0.01, 0.98]]).T
    w0 = w.copy()
    w0[:, :2] = 0.0             # zero-weight pixels: J = 0, Jac = identity
    w0[1:, 2:4] = 0.0           # rank-1 pixels
    nm = 0.5
    J_iso = pe.phasegradient2J(ks_ani, grads, w, nm)
    Jac = np.eye(2) + J_iso
    np.savez_compressed(
        os.path.join(OUT, "props_64x48.npz"), in_ks=ks, in_ks_ani=ks_ani, in_grads=grads, in_weights=w,
        in_weights_rankdef=w0, in_nmperpixel=nm,
        out_J_iso=J_iso,
        out_J_plain=pe.phasegradient2J(ks_ani, grads, w, nm, iso_ref=False),
        out_J_sorted=pe.phasegradient2J(ks_ani, grads, w, nm, sort=1),
        out_J_sorted_neg=pe.phasegradient2J(ks_ani, grads, w, nm, sort=-1),
        out_J_rankdef=pe.phasegradient2J(ks_ani, grads, w0, nm),
        out_Jac=pe.phasegradient2Jac(ks_ani, grads, w, nm),
        out_props=pe.props_from_Jac(Jac),
        out_props_diff=pe.props_from_Jac(Jac, refangle=3.0, refscale=2.0, diff=True),
        out_props_rankdef=pe.props_from_Jac(np.eye(2) + pe.phasegradient2J(ks_ani, grads, w0, nm)),
        out_props_from_J=pe.props_from_J(J_iso, refangle=-1.5, refscale=0.7))


def gen_iterate(gpa, pu):
    """iterate_GPA (geometric_phase_analysis.py:116-154) and mathtools.fit_plane (mathtools.py:30-47):
    a 96 x 80 lattice whose true k-vectors are 3 % off the ones handed to iterate_GPA."""
    import pyGPA.mathtools as mt
    shape = (96, 80)
    ks_true = synth.primary_ks(0.1, 7.0, 3)
    ks_guess = ks_true * 1.03
    u = synth.gaussian_bump(shape) * 0.5
    img = synth.lattice_image(shape, ks_true, u, noise=0.2, seed=21)
    img = img - img.mean()
    sigma = 6
    prs, w, corr = gpa.iterate_GPA(img, ks_guess, sigma, edge=4, iters=2, kmax_iter=15, kmax=60)
    prs0, w0, corr0 = gpa.iterate_GPA(img, ks_guess, sigma, edge=0, iters=1, kmax_iter=10, kmax=20)
    rng = np.random.default_rng(4)
    xx, yy = np.meshgrid(np.arange(70), np.arange(50), indexing='ij')
    plane = 0.31 * xx - 0.17 * yy + 2.5 + 0.4 * rng.normal(size=xx.shape)
    out = rng.uniform(size=xx.shape) < 0.08
    plane[out] += 15 * rng.normal(size=out.sum())
    np.savez_compressed(os.path.join(OUT, "iterate_96x80.npz"), in_image=img, in_ks=ks_guess, in_ks_true=ks_true,
                        in_sigma=sigma, out_prs=prs, out_w=w, out_corr=corr,
                        out_prs_edge0=prs0, out_w_edge0=w0, out_corr_edge0=corr0,
                        in_plane=plane, out_plane_fit=mt.fit_plane(plane),
                        out_delta_k=gpa.fit_delta_k(plane))


def gen_ucell(gpa, pu):
    """unit_cell_average / expand_unitcell (unit_cell_averaging.py:132-249), the constructions of the
    reference's tests/test_unit_cell_averaging.py at reduced size (r_k = 0.05, 96 x 80, z = 2 and 3)."""
    import pyGPA.unit_cell_averaging as uc
    shape = (96, 80)
    ks3 = synth.primary_ks(0.05, 7.0, 3)
    ks = ks3[:2]
    u = synth.gaussian_bump(shape) * 0.8
    img = synth.lattice_image(shape, ks3, None, second_order=0.3)
    img = img / img.max()
    img_d = synth.lattice_image(shape, ks3, u, second_order=0.3)
    img_d = img_d / img_d.max()
    img_nan = img.copy()
    img_nan[10:30, 20:50] = np.nan
    out = {}
    for z in (2, 3):
        cell = uc.unit_cell_average(img, ks, z=z)
        out[f"out_cell_z{z}"] = cell
        out[f"out_expand_z{z}"] = uc.expand_unitcell(cell, ks, shape, z=z)
        cell_d = uc.unit_cell_average(img_d, ks, z=z, u=u)
        out[f"out_cell_def_z{z}"] = cell_d
        out[f"out_expand_def_z{z}"] = uc.expand_unitcell(cell_d, ks, shape, z=z, u=u)
    out["out_cell_nan"] = uc.unit_cell_average(img_nan, ks, z=2)
    out["out_expand_z2_zoom"] = uc.expand_unitcell(out["out_cell_z2"], ks, (120, 100), z=2, z2=1.5)
    np.savez_compressed(os.path.join(OUT, "ucell_96x80.npz"), in_image=img, in_image_def=img_d, in_image_nan=img_nan,
                        in_ks=ks, in_u=u, **out)


def gen_base(gpa, pu):
    # ---- adaptive sweep: wfr2_grad_opt + optwfr2, non-square frame, 3 peaks -------------
    shape = (64, 48)
    ks = synth.primary_ks(0.12, 7.0, 3)
    u = synth.gaussian_bump(shape) * 0.6
    img = synth.lattice_image(shape, ks, u, second_order=0.3, noise=0.25, seed=11)
    img = img - img.mean()
    sigma = 4
    kw = float(np.linalg.norm(ks, axis=1).mean() / 2.5)
    kstep = kw / 3                                    # the reference tests' choice (tests/...:87-89)
    gs = [gpa.wfr2_grad_opt(img, sigma, k[0], k[1], kw, kstep) for k in ks]
    g2 = gpa.optwfr2(img, sigma, ks[0][0], ks[0][1], kw, kstep)
    np.savez_compressed(
        os.path.join(OUT, "sweep_64x48.npz"),
        in_image=img, in_ks=ks, in_sigma=sigma, in_kw=kw, in_kstep=kstep,
        out_lockin=np.stack([g['lockin'] for g in gs]),
        out_w=np.stack([g['w'] for g in gs]),
        out_grad=np.stack([g['grad'] for g in gs]),
        out_optwfr2_lockin=g2['lockin'], out_optwfr2_w=g2['w'])

    # ---- wfr3 on an explicit k-list --------------------------------------------------------
    rng = np.random.default_rng(5)
    klist = ks[1] + 0.03 * rng.uniform(-1, 1, size=(9, 2))
    g3 = gpa.wfr3(img, sigma, klist, ks[1])
    np.savez_compressed(os.path.join(OUT, "wfr3_64x48.npz"), in_image=img, in_sigma=sigma,
                        in_klist=klist, in_kref=ks[1], out_lockin=g3['lockin'], out_w=g3['w'])

    # ---- adaptive tail: reconstruct_u_inv_from_phases / extract_displacement_field -----------
    phases = np.stack([np.angle(g['lockin']) for g in gs])
    weights = np.stack([np.abs(g['lockin']) for g in gs])
    mask = np.zeros(shape)
    mask[2 * sigma:-2 * sigma, 2 * sigma:-2 * sigma] = 1
    weights_m = weights * (mask + 1e-6)
    u_fp = gpa.reconstruct_u_inv_from_phases(ks, phases, weights_m)
    u_fp_unw = gpa.reconstruct_u_inv_from_phases(ks, phases, weights_m, weighted_unwrap=False)
    grads = np.stack([g['grad'] for g in gs])
    u_fp_pre = gpa.reconstruct_u_inv_from_phases(ks, grads, weights_m, pre_diff=True)
    u_edf = gpa.extract_displacement_field(img, ks, sigma=sigma)
    np.savez_compressed(
        os.path.join(OUT, "tail_64x48.npz"), in_ks=ks, in_phases=phases, in_weights=weights_m,
        in_grads=grads, in_image=img, in_sigma=sigma,
        out_u=u_fp, out_u_unweighted_unwrap=u_fp_unw, out_u_prediff=u_fp_pre, out_u_edf=u_edf)

    # ---- fixed-reference path (config 1 in miniature) ------------------------------------------
    shape1 = (64, 64)
    ks1 = synth.primary_ks(0.1, 7.0, 3)
    u1 = synth.gaussian_bump(shape1)
    img1 = synth.lattice_image(shape1, ks1, u1, noise=0.1, seed=3)
    sig1 = 6
    rs = np.stack([gpa.optGPA(img1, k, sig1) for k in ks1])
    amps = np.abs(rs)
    unw = np.stack([pu.phase_unwrap(np.angle(r), np.sqrt(a / a.max()), kmax=25) for r, a in zip(rs, amps)])
    wdef = amps.copy()
    wdef[:, :3] = 0.0            # rank-0 pixels
    wdef[1:, 3:6] = 0.0          # rank-1 pixels
    np.savez_compressed(
        os.path.join(OUT, "fixed_64x64.npz"), in_image=img1, in_ks=ks1, in_sigma=sig1,
        in_weights_rankdef=wdef,
        out_lockin=rs, out_unwrapped=unw,
        out_u_unweighted=gpa.reconstruct_u_inv(ks1, unw),
        out_u_weighted=gpa.reconstruct_u_inv(ks1, unw, amps),
        out_u_rankdef=gpa.reconstruct_u_inv(ks1, unw, wdef),
        out_u_two_ks=gpa.reconstruct_u_inv(ks1, unw, use_only_ks=[0, 2]))

    # ---- phase unwrap known answers (tests/test_phase_unwrap.py, N reduced to 64) ---------------
    n = 64
    xx, yy = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
    psi0 = (yy + xx) / (4 * np.sqrt(2))
    psi = pu._wrapToPi(psi0)
    gauss = np.exp(-((xx - n // 2) ** 2 + (yy - n // 2) ** 2) / (0.3 * n ** 2))
    rng = np.random.default_rng(7)
    shape_r = (40, 56)                                # non-square: exercises the swapped N/M scale
    psi_r = pu._wrapToPi(3 * rng.normal(size=shape_r).cumsum(axis=0).cumsum(axis=1) / 20)
    w_r = rng.uniform(0.05, 1, size=shape_r)
    np.savez_compressed(
        os.path.join(OUT, "unwrap.npz"), in_psi=psi, in_psi0=psi0, in_gauss=gauss,
        in_psi_r=psi_r, in_w_r=w_r,
        out_ramp_k1=pu.phase_unwrap(psi, np.ones_like(psi), kmax=1),
        out_ramp_unweighted=pu.phase_unwrap(psi, None, kmax=30),
        out_ramp_gauss=pu.phase_unwrap(psi, gauss),
        out_r_k5=pu.phase_unwrap(psi_r, w_r, kmax=5),
        out_r_k100=pu.phase_unwrap(psi_r, w_r, kmax=100),
        out_r_unweighted=pu.phase_unwrap(psi_r, None),
        out_r_prediff_k7=pu.phase_unwrap_prediff(np.diff(psi_r, axis=1), np.diff(psi_r, axis=0), w_r, kmax=7),
        out_r_prediff_unweighted=pu.phase_unwrap_prediff(np.diff(psi_r, axis=1), np.diff(psi_r, axis=0)))

    # ---- Lawler-Fujita ----------------------------------------------------------------------------
    shape_l = (48, 40)
    ks_l = synth.primary_ks(0.11, 7.0, 3)
    u_l = synth.gaussian_bump(shape_l) * 0.8
    u_l[1] = 0.4 * np.sin(2 * np.pi * np.arange(shape_l[0])[:, None] / shape_l[0]) * np.ones(shape_l)
    img_l = synth.lattice_image(shape_l, ks_l, u_l)
    np.savez_compressed(
        os.path.join(OUT, "lawler_fujita_48x40.npz"), in_u=u_l, in_image=img_l,
        out_invert_edge0=gpa.invert_u_overlap(u_l),
        out_invert_edge3_it5=gpa.invert_u_overlap(u_l, iters=5, edge=3),
        out_undistorted=gpa.undistort_image(img_l, u_l))


SECTIONS = {"base": gen_base, "wfr4": gen_wfr4, "props": gen_props, "iterate": gen_iterate, "ucell": gen_ucell}


def main(argv=None):
    """python -m oracle.gen_golden [section ...]   (default: every section)"""
    import sys
    names = list(argv if argv is not None else sys.argv[1:]) or list(SECTIONS)
    gpa, pu = _refimport.load()
    os.makedirs(OUT, exist_ok=True)
    for name in names:
        SECTIONS[name](gpa, pu)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
