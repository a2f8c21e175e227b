"""Comparators shared by the GPU parity tests.

Tolerances are BASELINE.json's: selected k-index identical except at documented near-ties,
phase within 1e-3 rad, displacement within 1e-3 px.

Near-tie (hard part 1): a pixel is a near-tie iff the oracle's two
largest candidate amplitudes differ by less than NEAR_TIE relative.  Only there may the
selected candidate differ (fp32 arithmetic and the 4.5-sigma truncation perturb |sf| by a
few 1e-6 relative); everywhere else it must be bit-identical.
"""
import numpy as np

NEAR_TIE = 1e-5
PHASE_TOL = 1e-3      # rad
DISP_TOL = 1e-3       # px


def wrap(x):
    return (x + np.pi) % (2 * np.pi) - np.pi


def check_sweep(got, ref, check_grad=True, amp_floor=0.02):
    """got: dict from the CUDA path; ref: oracle.wfr_sweep(..., return_diag=True).
    Returns a stats dict; raises AssertionError on violation."""
    same = np.all(got['w'] == ref['w'], axis=0)
    gap = (ref['amp1'] - ref['amp2']) / np.maximum(ref['amp1'], 1e-300)
    bad = ~same & ~(gap < NEAR_TIE)
    assert not bad.any(), f"{bad.sum()} pixels pick another k away from a near-tie (max gap {gap[~same].max():.3g})"
    amax = np.abs(ref['lockin']).max()
    # where the same candidate was picked the signals must agree
    dl = np.abs(got['lockin'] - ref['lockin'])
    assert dl[same].max() <= 1e-4 * amax, f"lock-in differs by {dl[same].max() / amax:.3g} of max"
    m = same & (np.abs(ref['lockin']) > amp_floor * amax)
    ph = np.abs(np.angle(got['lockin'][m] * np.conj(ref['lockin'][m])))
    assert ph.max() < PHASE_TOL, f"phase error {ph.max():.3g} rad"
    stats = dict(mismatch=int((~same).sum()), frac_mismatch=float((~same).mean()),
                 phase_err=float(ph.max()), lockin_err=float(dl[same].max() / amax))
    if check_grad:
        fin = np.isfinite(ref['grad'])
        assert np.array_equal(fin, np.isfinite(got['grad'])), "NaN pattern of grad differs"
        dg = np.abs(wrap(2 * (got['grad'] - ref['grad'])) / 2)      # grad is defined mod pi
        mg = m[..., None] & fin
        assert dg[mg].max() < PHASE_TOL, f"gradient error {dg[mg].max():.3g}"
        stats['grad_err'] = float(dg[mg].max())
    return stats
