"""K5 on the GPU: phasegradient2J / props_from_Jac through the C ABI against the reference-generated
fixture and the oracle (the per-pixel code is additionally checked on the CPU by test_props_host.py)."""
import numpy as np
import pytest

import oracle
from conftest import load_golden
from pygpa_b200 import property_extract as PE

pytestmark = pytest.mark.gpu


def _circ(d, period):
    return np.minimum(np.abs(d) % period, period - np.abs(d) % period)


def test_phasegradient2J_matches_reference_fixture():
    g = load_golden("props_64x48.npz")
    ks, grads, w, nm = g["in_ks_ani"], g["in_grads"], g["in_weights"], float(g["in_nmperpixel"])
    tol = dict(rtol=1e-9, atol=1e-11)
    assert np.allclose(PE.phasegradient2J(ks, grads, w, nm), g["out_J_iso"], **tol)
    assert np.allclose(PE.phasegradient2J(ks, grads, w, nm, iso_ref=False), g["out_J_plain"], **tol)
    assert np.allclose(PE.phasegradient2J(ks, grads, w, nm, sort=1), g["out_J_sorted"], **tol)
    assert np.allclose(PE.phasegradient2J(ks, grads, w, nm, sort=-1), g["out_J_sorted_neg"], **tol)
    jr = PE.phasegradient2J(ks, grads, g["in_weights_rankdef"], nm)
    assert np.allclose(jr, g["out_J_rankdef"], **tol)
    assert np.all(jr[:2] == 0.0)                                   # zero weights -> minimum-norm 0
    jac = PE.phasegradient2Jac(ks, grads, w, nm)
    assert jac.shape == grads.shape[1:3] + (2, 2) and np.allclose(jac, g["out_Jac"], **tol)
    with pytest.raises(ValueError):
        PE.phasegradient2J(ks[:2], grads[:2], w[:2], nm)


def test_props_from_Jac_matches_reference_fixture():
    g = load_golden("props_64x48.npz")
    tol = dict(rtol=1e-9, atol=1e-7)
    assert np.allclose(PE.props_from_Jac(g["out_Jac"]), g["out_props"], **tol)
    assert np.allclose(PE.props_from_Jac(g["out_Jac"], refangle=3.0, refscale=2.0, diff=True), g["out_props_diff"], **tol)
    pr = PE.props_from_J(g["out_J_rankdef"])
    assert np.allclose(pr, g["out_props_rankdef"], **tol)
    assert np.array_equal(pr[:, 0, 0], [0., 0., 1., 1.])
    assert np.allclose(PE.props_from_J(g["out_J_iso"], refangle=-1.5, refscale=0.7), g["out_props_from_J"], **tol)
    one = PE.props_from_Jac(g["out_Jac"][5, 7])                    # a single 2x2 matrix, like the reference allows
    assert one.shape == (4,) and np.allclose(one, g["out_props"][:, 5, 7], **tol)


def test_props_pipeline_at_size_against_oracle():
    """sweep gradients -> J -> props for a 256 x 192 frame, device path vs oracle on the same inputs."""
    rng = np.random.default_rng(11)
    n, m = 256, 192
    ks = np.array([[0.1, 0.012], [0.039, 0.093], [-0.06, 0.081]])
    grads = rng.uniform(-0.3, 0.3, size=(3, n, m, 2))
    w = rng.uniform(0.0, 1.0, size=(3, n, m))
    J = PE.phasegradient2J(ks, grads, w, 2.0)
    Jo = oracle.phasegradient2J(ks, grads, w, 2.0)
    assert np.allclose(J, Jo, rtol=1e-8, atol=1e-10)
    pr, po = PE.props_from_J(Jo), oracle.props_from_J(Jo)
    kappa = po[3]
    ok = kappa - 1 > 1e-6                       # the angles are ill-conditioned as the anisotropy vanishes
    assert np.abs(pr[0] - po[0])[ok].max() < 1e-6 and _circ(pr[1] - po[1], 180.0)[ok].max() < 1e-6
    assert np.allclose(pr[2:], po[2:], rtol=1e-10)
