"""K7 on the GPU: unit_cell_average / expand_unitcell against fixtures produced by the unmodified
reference (oracle/gen_golden.py, section 'ucell') and the reference test's own accuracy bounds."""
import numpy as np
import pytest

import oracle
from conftest import load_golden
from pygpa_b200 import synth
from pygpa_b200 import unit_cell_averaging as UC

pytestmark = pytest.mark.gpu


def _same_cell(got, ref):
    assert got.shape == ref.shape and np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.abs(got - ref)[ok].max() < 1e-11


@pytest.mark.parametrize("z", [2, 3])
def test_project_and_expand_match_reference_fixture(z):
    g = load_golden("ucell_96x80.npz")
    ks, u, shape = g["in_ks"], g["in_u"], g["in_image"].shape
    cell = UC.unit_cell_average(g["in_image"], ks, z=z)
    _same_cell(cell, g[f"out_cell_z{z}"])
    assert np.abs(UC.expand_unitcell(g[f"out_cell_z{z}"], ks, shape, z=z) - g[f"out_expand_z{z}"]).max() < 1e-11
    cell_d = UC.unit_cell_average(g["in_image_def"], ks, u=u, z=z)
    _same_cell(cell_d, g[f"out_cell_def_z{z}"])
    exp_d = UC.expand_unitcell(cell_d, ks, shape, z=z, u=u)
    assert np.abs(exp_d - g[f"out_expand_def_z{z}"]).max() < 1e-10


def test_nan_mask_zoomed_expand_and_generated_function():
    g = load_golden("ucell_96x80.npz")
    ks = g["in_ks"]
    _same_cell(UC.unit_cell_average(g["in_image_nan"], ks, z=2), g["out_cell_nan"])
    got = UC.expand_unitcell(g["out_cell_z2"], ks, (120, 100), z=2, z2=1.5)
    assert np.abs(got - g["out_expand_z2_zoom"]).max() < 1e-11
    f = UC.unit_cell_average(None, ks, z=2, only_generate_func=True)
    _same_cell(f(g["in_image_def"], np.moveaxis(g["in_u"], 0, -1)), g["out_cell_def_z2"])


@pytest.mark.parametrize("z", [2, 3])
def test_reference_round_trip_accuracy(z):
    """tests/test_unit_cell_averaging.py:10-25 of the reference (r_k = 0.02, 200 x 200, order 2): the
    expanded average reproduces the lattice; plus oracle parity at that size."""
    shape = (200, 200)
    ks3 = synth.primary_ks(0.02, 7.0, 3)
    img = synth.lattice_image(shape, ks3, None, second_order=0.3)
    img = img / img.max()
    cell = UC.unit_cell_average(img, ks3[:2], z=z)
    back = UC.expand_unitcell(cell, ks3[:2], shape, z=z)
    # the reference bounds the maximum by 0.11 on latticegen's lattice; this synthetic one (sharper
    # second-order terms) gives 0.18 / 0.23 in the reference implementation itself (oracle, z = 2 / 3)
    assert np.abs(img - back).mean() < 5e-3 and np.abs(img - back).max() < 0.25
    ref_cell = oracle.unit_cell_average(img, ks3[:2], z=z)
    _same_cell(cell, ref_cell)
    assert np.abs(back - oracle.expand_unitcell(ref_cell, ks3[:2], shape, z=z)).max() < 1e-10
