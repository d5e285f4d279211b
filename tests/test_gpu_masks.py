"""GPU parity tests of the LocGridOrbitals path (SURVEY a19, a15-a17 per
x-slab): localization masks inside the V-cycle and on orbital blocks, and the
per-slab contractions, through the C ABI against the oracle and against golden
vectors produced by the reference's own GridMask / Map2Masks /
Preconditioning objects.

Bars: the literal V-cycle and every mask application are BIT-IDENTICAL; the
fused V-cycle within 5e-6 of the output max norm (FP32 bar 1e-5)."""
import os

import numpy as np
import pytest
import torch

import mask_cases as mc
from conftest import ROOT, bits_equal, rel_inf
from oracle.oracle import ghosts_for, synthetic_masks, synthetic_orbitals

pytestmark = pytest.mark.gpu

TDT = {np.float64: torch.float64, np.float32: torch.float32}
MG_TOL = 5e-6


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def H():
    from mgmol_b200 import host as h
    return h


@pytest.fixture(scope="module")
def gmask():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_masks.npz")))


def _precond(H, lt, levels, res, dims, ll, bc, mt, mode, gamma, twice=False):
    grid = H.Grid(dims, ll, H.ghosts_for(lt), bc)
    masks = H.Masks.from_tables(grid, mt)
    orb = H.Orbitals(grid, res.shape[0], TDT[res.dtype.type], dev(res))
    pc = H.OrbitalsPreconditioning()
    pc.setup(orb, levels, lt, masks)
    pc.set_mode(mode)
    pc.gamma_ = gamma
    pc.precond_mg(orb)
    assert pc.last_mode() == mode
    if twice:
        pc.precond_mg(orb)
    out = host(orb.psi())
    pc.close()
    masks.close()
    return out


@pytest.mark.parametrize("op", [0, 1])
@pytest.mark.parametrize("dt,tag", [(np.float64, "f64"), (np.float32, "f32")])
@pytest.mark.parametrize("lt", [0, 2])
@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 0, 0)])
def test_masked_precond_against_golden(H, gmask, op, dt, tag, lt, bc):
    """Masks generated and applied by the reference itself (golden)."""
    mt = mc.tables_from_arrays(gmask, "masks_op%d_g%d" % (op, ghosts_for(lt)), op)
    res = synthetic_orbitals(mc.GID_TABLE.shape[1], mc.DIMS, dt)
    ref = gmask["precond_op%d_lap%d_%s_bc%d%d%d" % ((op, lt, tag) + bc)]
    lit = _precond(H, lt, mc.LEVELS, res, mc.DIMS, mc.LL, bc, mt, 1, mc.GAMMA)
    assert bits_equal(lit, ref)
    fus = _precond(H, lt, mc.LEVELS, res, mc.DIMS, mc.LL, bc, mt, 2, mc.GAMMA)
    assert rel_inf(fus, ref) <= MG_TOL
    # structure survives the fused kernels exactly: zero slabs stay zero
    assert np.array_equal(fus == 0, ref == 0) or rel_inf(fus, ref) <= MG_TOL
    s0 = mc.DIMS[0] // mc.SUBDIVX
    assert not fus[2, :s0].any()


def _case(dims, ll, levels, subdivx, ncolors, ngid, op, g, seed):
    rng = np.random.default_rng(seed)
    centers = rng.uniform(0, 1, (ngid, 3)) * np.array(ll)
    radii = rng.uniform(0.2, 0.45, ngid) * min(ll)
    gid_table = rng.integers(-1, ngid, (subdivx, ncolors))
    return gid_table, synthetic_masks(dims, ll, levels, subdivx, gid_table, centers, radii, op,
                                      ghosts=g)


@pytest.mark.parametrize("op", [0, 1])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("lt", [0, 2])
@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 0, 0), (1, 0, 1)])
@pytest.mark.parametrize("levels,dims,subdivx", [(1, (16, 24, 32), 2), (2, (32, 16, 64), 4),
                                                 (0, (8, 8, 8), 1)])
def test_masked_precond_vs_oracle(H, port, op, dt, lt, bc, levels, dims, subdivx):
    ll, N = tuple(0.25 * d for d in dims), 5
    gid_table, mt = _case(dims, ll, levels, subdivx, N, 9, op, ghosts_for(lt), 5 + op)
    res = synthetic_orbitals(N, dims, dt)
    ref = port.precond_mg_masked(lt, levels, res, ll, 0.31, mt, bc)
    lit = _precond(H, lt, levels, res, dims, ll, bc, mt, 1, 0.31)
    assert bits_equal(lit, ref)
    if bc[0] == bc[1] == bc[2]:  # mixed boxes stay literal
        fus = _precond(H, lt, levels, res, dims, ll, bc, mt, 2, 0.31)
        assert rel_inf(fus, ref) <= MG_TOL
        ref2 = port.precond_mg_masked(lt, levels, ref, ll, 0.31, mt, bc)
        fus2 = _precond(H, lt, levels, res, dims, ll, bc, mt, 2, 0.31, twice=True)
        assert rel_inf(fus2, ref2) <= 2 * MG_TOL


@pytest.mark.parametrize("op", [0, 1])
@pytest.mark.parametrize("lt", [0, 2])
def test_masked_precond_fused_equals_literal_larger(H, op, lt):
    """64 x 64 x 128, 12 colors, 4 slabs, 2 levels: fused against our own
    bit-exact literal sequence (sizes the CPU oracle would take long on)."""
    dims, ll, levels, subdivx, N = (64, 64, 128), (16.0, 16.0, 32.0), 2, 4, 12
    gid_table, mt = _case(dims, ll, levels, subdivx, N, 20, op, ghosts_for(lt), 21)
    res = synthetic_orbitals(N, dims, np.float32)
    lit = _precond(H, lt, levels, res, dims, ll, (1, 1, 1), mt, 1, 0.3)
    fus = _precond(H, lt, levels, res, dims, ll, (1, 1, 1), mt, 2, 0.3)
    assert rel_inf(fus, lit) <= MG_TOL
    assert np.abs(lit).max() > 0


@pytest.mark.parametrize("op", [0, 1])
@pytest.mark.parametrize("dt,tag", [(np.float64, "f64"), (np.float32, "f32")])
def test_app_mask(H, port, gmask, op, dt, tag):
    """LocGridOrbitals::applyMask (no ghosts) and GridFuncVector::app_mask
    (ghosted, every level), bit-identical."""
    from mgmol_b200._lib import lib, check
    mt = mc.tables_from_arrays(gmask, "masks_op%d_g1" % op, op)
    nc = mc.GID_TABLE.shape[1]
    res = synthetic_orbitals(nc, mc.DIMS, dt)
    grid = H.Grid(mc.DIMS, mc.LL, 1)
    masks = H.Masks.from_tables(grid, mt)
    orb = H.LocGridOrbitals(grid, len(mc.GIDS), mc.GID_TABLE, TDT[dt], dev(res), masks)
    idx = orb.getIterativeIndex()
    orb.applyMask()
    assert orb.getIterativeIndex() == idx + 1
    assert bits_equal(host(orb.psi()), gmask["appmask_op%d_%s" % (op, tag)])
    for level in range(mc.LEVELS + 1):
        d = tuple(x >> level for x in mc.DIMS)
        for g in (1, 2):
            u = port.trade_boundaries(synthetic_orbitals(nc, d, dt, first=level), g)
            t = dev(u)
            check(lib().mgb_gfv_app_mask(0 if dt == np.float32 else 1, masks.handle_, level, g,
                                         t.data_ptr(), nc, None))
            assert bits_equal(host(t), port.app_mask_ghosted(u, g, mt, level))
    # no mask set: no-op, like map2masks_ == nullptr
    t = dev(res)
    check(lib().mgb_app_mask(0 if dt == np.float32 else 1, None, 0, t.data_ptr(),
                             int(np.prod(mc.DIMS)), nc, None))
    assert bits_equal(host(t), res)
    masks.close()


def test_masks_errors(H):
    from mgmol_b200._lib import MgbError
    grid = H.Grid((16, 16, 16), (4.0, 4.0, 4.0), 1)
    with pytest.raises(MgbError):
        H.Masks(grid, 2, 3, 4, 0)            # 16 >> 2 not divisible by 3 slabs
    with pytest.raises(MgbError):
        H.Masks(grid, 1, 2, 4, 7)            # unknown mask operation
    m = H.Masks(grid, 1, 2, 4, 0)
    with pytest.raises(MgbError):
        m.set(0, 2, 0, 1)                    # iloc out of range
    with pytest.raises(MgbError):
        m.set(0, 0, 0, 2, None)              # state 2 without values
    orb = H.Orbitals(grid, 4, torch.float32)
    pc = H.OrbitalsPreconditioning()
    with pytest.raises(MgbError):
        pc.setup(orb, 2, 0, m)               # V-cycle needs 3 levels, set has 2
    m.close()


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_locgrid_contractions_per_slab(H, port, dt):
    """getLocalOverlap / computeLocalProduct / multiplyByMatrix /
    multiply_by_matrix of LocGridOrbitals: one contraction per x-slab
    (src/LocGridOrbitals.cc:750-898,1400-1424,1504-1604)."""
    dims, ll, subdivx, N, numst = (16, 8, 12), (4.0, 2.0, 3.0), 4, 5, 9
    rng = np.random.default_rng(17)
    gid_table = rng.integers(-1, numst, (subdivx, N))
    phi = synthetic_orbitals(N, dims, dt)
    other = synthetic_orbitals(N, dims, dt, first=31)
    grid = H.Grid(dims, ll, 1)
    orb = H.LocGridOrbitals(grid, numst, gid_table, TDT[dt], dev(phi))
    oth = H.LocGridOrbitals(grid, numst, gid_table, TDT[dt], dev(other))
    s0 = dims[0] // subdivx
    tol = 1e-12 if dt == np.float64 else 3e-6   # float: 3xTF32 tensor tiles
    ss = host(orb.getLocalOverlap())
    pp = host(orb.computeLocalProduct(oth))
    pt = host(orb.computeLocalProduct(oth, transpose=True))
    for iloc in range(subdivx):
        a = phi[:, iloc * s0:(iloc + 1) * s0]
        b = other[:, iloc * s0:(iloc + 1) * s0]
        ex = port.gemm_tn(a, a, grid.vel())
        assert np.abs(ss[iloc] - ex).max() <= tol * np.abs(ex).max()
        assert np.array_equal(ss[iloc], ss[iloc].T)
        ex = port.gemm_tn(a, b, grid.vel())
        assert np.abs(pp[iloc] - ex).max() <= tol * np.abs(ex).max()
        assert np.abs(pt[iloc] - ex.T).max() <= tol * np.abs(ex).max()
    # matrixToLocalMatrix: literal loops of the reference
    M = rng.standard_normal((numst, numst))
    for iloc in range(subdivx):
        lm = np.zeros((N, N))
        for j in range(N):
            for i in range(N):
                gi, gj = gid_table[iloc][i], gid_table[iloc][j]
                if gi != -1 and gj != -1:
                    lm[i, j] = M[gi, gj]
        assert np.array_equal(host(orb.matrixToLocalMatrix(iloc, dev(M))), lm)
    out = host(orb.multiply_by_matrix(dev(M), product=oth))
    for iloc in range(subdivx):
        lm = host(orb.matrixToLocalMatrix(iloc, dev(M)))
        ex = port.gemm_nn(phi[:, iloc * s0:(iloc + 1) * s0], lm)
        got = out[:, iloc * s0:(iloc + 1) * s0]
        assert np.abs(got - ex).max() <= (1e-13 if dt == np.float64 else 3e-6) * np.abs(ex).max()
