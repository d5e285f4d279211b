"""CPU tests (no GPU) of the localization-mask path: the C
restatement against golden vectors made by the reference's own GridMask /
Map2Masks / Preconditioning objects, and bit-for-bit against the compiled
reference where oracle/_ref exists."""
import os

import numpy as np
import pytest

import mask_cases as mc
from conftest import ROOT, bits_equal
from oracle.oracle import Ref, ghosts_for, synthetic_masks, synthetic_orbitals


@pytest.fixture(scope="module")
def gmask():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_masks.npz")))


@pytest.mark.parametrize("op", [0, 1])
@pytest.mark.parametrize("dt,tag", [(np.float64, "f64"), (np.float32, "f32")])
def test_masked_precond_matches_golden(port, gmask, op, dt, tag):
    nc = mc.GID_TABLE.shape[1]
    res = synthetic_orbitals(nc, mc.DIMS, dt)
    for lt in (0, 2):
        mt = mc.tables_from_arrays(gmask, "masks_op%d_g%d" % (op, ghosts_for(lt)), op)
        for bc in ((1, 1, 1), (0, 0, 0)):
            ref = gmask["precond_op%d_lap%d_%s_bc%d%d%d" % ((op, lt, tag) + bc)]
            got = port.precond_mg_masked(lt, mc.LEVELS, res, mc.LL, mc.GAMMA, mt, bc)
            assert bits_equal(got, ref), (lt, bc)
            # masks matter: the unmasked V-cycle differs
            assert not np.array_equal(port.precond_mg(lt, mc.LEVELS, res, mc.LL, mc.GAMMA, bc),
                                      ref)


@pytest.mark.parametrize("op", [0, 1])
@pytest.mark.parametrize("dt,tag", [(np.float64, "f64"), (np.float32, "f32")])
def test_app_mask_noghost_matches_golden(port, gmask, op, dt, tag):
    nc = mc.GID_TABLE.shape[1]
    res = synthetic_orbitals(nc, mc.DIMS, dt)
    mt = mc.tables_from_arrays(gmask, "masks_op%d_g1" % op, op)
    got = port.app_mask_noghost(res, mt)
    assert bits_equal(got, gmask["appmask_op%d_%s" % (op, tag)])
    # the empty slot (gid -1) of slab 0 is zeroed, the radius-200 orbital kept
    s0 = mc.DIMS[0] // mc.SUBDIVX
    assert not got[2, :s0].any() and np.array_equal(got[3], res[3])


def test_ghosted_and_noghost_mask_agree(port, gmask):
    """GridFuncVector::app_mask on a ghosted block = LocGridOrbitals::applyMask
    on its interior; zero slabs also clear the y/z ghosts of their planes."""
    nc = mc.GID_TABLE.shape[1]
    res = synthetic_orbitals(nc, mc.DIMS, np.float32)
    for op in (0, 1):
        mt = mc.tables_from_arrays(gmask, "masks_op%d_g1" % op, op)
        gh = port.trade_boundaries(res, 1)
        m = port.app_mask_ghosted(gh, 1, mt)
        assert bits_equal(port.strip_ghosts(m, 1), port.app_mask_noghost(res, mt))
        s0 = mc.DIMS[0] // mc.SUBDIVX
        assert not m[2, 1:1 + s0].any()          # whole planes incl. ghosts
        assert np.array_equal(m[2, 0], gh[2, 0])  # x ghost plane untouched


@pytest.mark.skipif(not Ref.available(), reason="compiled reference not present")
@pytest.mark.parametrize("op", [0, 1])
@pytest.mark.parametrize("lt", [0, 2])
def test_masked_port_is_bit_exact_against_compiled_reference(port, op, lt):
    R = Ref()
    g = ghosts_for(lt)
    dims, ll, levels, subdivx = (16, 8, 24), (5.0, 2.5, 7.0), 1, 4
    rng = np.random.default_rng(11 + op)
    ngid = 7
    centers = rng.uniform(0, 1, (ngid, 3)) * np.array(ll)
    radii = rng.uniform(0.8, 2.5, ngid)
    gid_table = np.array([[0, 1, 2], [3, -1, 2], [4, 5, -1], [0, 6, 1]])
    h = R.masks_create(dims, ll, g, levels, subdivx, op, list(range(ngid)), centers, radii)
    mt = R.masks_tables(h, gid_table)
    # the numpy restatement of the mask generation agrees with GridMask::init
    ms = synthetic_masks(dims, ll, levels, subdivx, gid_table, centers, radii, op, ghosts=g)
    for l in range(levels + 1):
        assert np.array_equal(np.maximum(mt.state[l], 0), np.maximum(ms.state[l], 0))
        for k, v in mt.values[l].items():
            assert np.abs(v - ms.values[l][k]).max() < 1e-6
    for dt in (np.float32, np.float64):
        res = synthetic_orbitals(3, dims, dt)
        for bc in ((1, 1, 1), (0, 0, 0), (1, 0, 1)):
            a = R.precond_mg_masked(lt, res, 0.41, h, gid_table, bc)
            b = port.precond_mg_masked(lt, levels, res, ll, 0.41, mt, bc)
            assert bits_equal(a, b), (dt, bc)
        assert bits_equal(R.app_mask_noghost(res, h, gid_table), port.app_mask_noghost(res, mt))
