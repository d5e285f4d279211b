"""GPU parity tests: every CUDA path, called through the C ABI, against the
oracle (our C restatement, itself pinned bit-for-bit to the reference's
compiled sources) on the same seeded inputs, and against the committed golden
vectors of the compiled reference.

Tolerances (BASELINE.json north_star): H psi elementwise within 1e-12 (FP64) /
1e-5 (FP32), measured relative to the per-orbital max norm of the reference
result (entries near zero carry the cancellation error of the stencil's large
terms).  The literal kernels (ghosted-block path and the
generic fused kernel) are required to be BIT-IDENTICAL.
"""
import os

import numpy as np
import pytest
import torch

from conftest import bits_equal, rel_elem, rel_inf
from oracle.oracle import synthetic_orbitals, synthetic_potential

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-12, np.float32: 1e-5}
TDT = {np.float64: torch.float64, np.float32: torch.float32}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def H():
    from mgmol_b200 import host as h
    return h


def _hpsi_gpu(H, lap_type, phi, v, ll, bc, path):
    from mgmol_b200._lib import lib, check
    grid = H.Grid(phi.shape[1:], ll, H.ghosts_for(lap_type), bc)
    lap = H.LapFactory.createLap(grid, lap_type)
    dphi, dv = dev(phi), dev(v)
    out = torch.full_like(dphi, float("nan"))
    check(lib().mgb_hpsi_force_path(path))
    try:
        lap.applyWithPot(dphi, dv, out)
        used = lib().mgb_hpsi_last_path()
    finally:
        lib().mgb_hpsi_force_path(0)
    torch.cuda.synchronize()
    return host(out), used


# --------------------------------------------------------------------------
# fused H psi
# --------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("lap_type", [0, 2])
@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 0, 0)])
@pytest.mark.parametrize("dims,N", [((12, 8, 16), 3), ((40, 40, 40), 4), ((16, 24, 32), 5),
                                    ((8, 16, 256), 2), ((20, 12, 20), 1)])
def test_hpsi_tma_and_generic(H, port, dt, lap_type, bc, dims, N):
    ll = (6.0, 5.0, 7.0)
    phi = synthetic_orbitals(N, dims, dt)
    v = synthetic_potential(dims)
    ref = port.hpsi(lap_type, phi, v, ll, bc)
    got2, used2 = _hpsi_gpu(H, lap_type, phi, v, ll, bc, 2)
    assert used2 == 2
    assert bits_equal(got2, ref), "generic fused kernel must be bit-identical"
    got1, used1 = _hpsi_gpu(H, lap_type, phi, v, ll, bc, 1)
    assert used1 == 1
    err = rel_inf(got1, ref)
    assert err <= TOL[dt], "TMA kernel rel err %g" % err
    # the elementwise relative error on the entries above 1e-3 of the orbital's max (the
    # others are cancellation residue of the stencil's large terms): reported, and bounded
    # by the max-norm bar / 1e-3
    el = rel_elem(got1, ref, 1e-3)
    print("H psi %s lap %d bc %s: max-norm rel %.2e, elementwise rel (|ref| >= 1e-3 max) %.2e"
          % (np.dtype(dt).name, lap_type, bc, err, el))
    assert el <= TOL[dt] / 1e-3


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("lap_type", [0, 2])
@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 0, 0)])
@pytest.mark.parametrize("zboxes", [False, True])
@pytest.mark.parametrize("dims,N", [((16, 24, 32), 5), ((8, 16, 64), 3), ((32, 32, 128), 2)])
def test_hpsi_peer3d_self_neighbours(H, port, dt, lap_type, bc, zboxes, dims, N):
    """mgb_hpsi_peer3d on one rank: every one of the 26 neighbours is the box itself (the
    periodic wrap) or beyond a Dirichlet end, so the kernel's halo machinery -- boxes from
    the neighbour table, the z-halo column arrays, the padded ghosted potential -- runs
    without a second GPU and must reproduce mgb_hpsi's TMA kernel bit for bit (same
    arithmetic per point) and the oracle within the path's tolerance."""
    from mgmol_b200._lib import lib, check
    ll = (4.0, 6.0, 8.0)
    phi = synthetic_orbitals(N, dims, dt)
    v = synthetic_potential(dims)
    g = H.ghosts_for(lap_type)
    grid = H.Grid(dims, ll, g, bc)
    dphi, dv = dev(phi), dev(v)
    ref = torch.empty_like(dphi)
    check(lib().mgb_hpsi_force_path(1))
    try:
        H.LapFactory.createLap(grid, lap_type).applyWithPot(dphi, dv, ref)
    finally:
        lib().mgb_hpsi_force_path(0)
    gv = H.GridFuncVector(grid, 1, torch.float64)
    gv.assign(dv[None].contiguous())
    gv.trade_boundaries()
    out = torch.full_like(dphi, float("nan"))
    if zboxes:
        os.environ["MGB_HPSI_FORCE_ZBOXES"] = "1"
    try:
        class _NoComm:
            handle = None
        H.LapFactory.createLap(grid, lap_type).applyWithPotPeer(_NoComm, dphi, dv, out,
                                                               vghost=gv.data)
    finally:
        os.environ.pop("MGB_HPSI_FORCE_ZBOXES", None)
    assert "PEER=2" in lib().mgb_hpsi_last_kernel().decode()
    assert torch.equal(out, ref)
    exp = port.hpsi(lap_type, phi, v, ll, bc)
    assert rel_inf(host(out), exp) <= TOL[dt]


@pytest.mark.parametrize("cfg", ["8,1,1,2,0", "8,1,2,3,0", "4,2,2,2,7", "4,1,3,2,5",
                                 "8,2,1,2,16", "4,4,1,4,3"])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_hpsi_tma_configs(H, port, dt, cfg):
    """tile shape, orbitals per CTA, pipeline depth and x-chunking must not
    change the result beyond rounding."""
    dims, N, ll = (24, 32, 32), 5, (5.0, 6.0, 6.5)
    phi = synthetic_orbitals(N, dims, dt)
    v = synthetic_potential(dims)
    for lap_type in (0, 2):
        c = cfg
        if lap_type == 2 and cfg.startswith("8,"):
            c = "4," + cfg[2:]
        os.environ["MGB_HPSI_CFG"] = c
        try:
            for bc in ((1, 1, 1), (0, 0, 0)):
                ref = port.hpsi(lap_type, phi, v, ll, bc)
                got, used = _hpsi_gpu(H, lap_type, phi, v, ll, bc, 1)
                assert used == 1
                assert rel_inf(got, ref) <= TOL[dt], (c, lap_type, bc)
        finally:
            del os.environ["MGB_HPSI_CFG"]


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("lap_type", [0, 2])
@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 0, 0), (1, 0, 1), (0, 1, 1), (1, 1, 0)])
def test_hpsi_ghosted_composition_bit_exact(H, port, dt, lap_type, bc):
    dims, N, ll = (12, 10, 14), 3, (3.0, 2.5, 3.5)
    phi = synthetic_orbitals(N, dims, dt)
    v = synthetic_potential(dims)
    ref = port.hpsi(lap_type, phi, v, ll, bc)
    got, used = _hpsi_gpu(H, lap_type, phi, v, ll, bc, 3)
    assert used == 3
    assert bits_equal(got, ref)
    # automatic dispatch picks a valid path for every boundary condition
    got0, used0 = _hpsi_gpu(H, lap_type, phi, v, ll, bc, 0)
    assert rel_inf(got0, ref) <= TOL[dt]


@pytest.mark.parametrize("dt,tag", [(np.float64, "f64"), (np.float32, "f32")])
def test_hpsi_against_golden_reference_vectors(H, golden, dt, tag):
    dims = tuple(int(x) for x in golden["dims"])
    ll = tuple(float(x) for x in golden["ll"])
    N = int(golden["nfunc"])
    phi = synthetic_orbitals(N, dims, dt)
    v = synthetic_potential(dims)
    for lap_type in (0, 2):
        for bc in ((1, 1, 1), (0, 0, 0), (1, 0, 1)):
            ref = golden["hpsi_lap%d_%s_bc%d%d%d" % ((lap_type, tag) + bc)]
            got, used = _hpsi_gpu(H, lap_type, phi, v, ll, bc, 0)
            assert rel_inf(got, ref) <= TOL[dt], (lap_type, bc, used)
            if used != 1:
                assert bits_equal(got, ref)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("lap_type", [0, 2])
@pytest.mark.parametrize("chunk", [0, 1, 3, 4])
def test_hpsi_host_buffers_pipeline(H, port, dt, lap_type, chunk):
    """mgb_hpsi_host (host blocks in, host blocks out; H2D / kernel / D2H
    pipelined over blocks of `chunk` orbitals) against the oracle and,
    bit-for-bit, against the resident-device call."""
    dims, ll, N = (16, 24, 32), (4.0, 6.0, 8.0), 10
    phi = synthetic_orbitals(N, dims, dt)
    v = synthetic_potential(dims)
    grid = H.Grid(dims, ll, H.ghosts_for(lap_type))
    lap = H.LapFactory.createLap(grid, lap_type)
    hphi = torch.from_numpy(phi.copy()).pin_memory()
    hv = torch.from_numpy(v.copy()).pin_memory()
    hout = torch.full(hphi.shape, float("nan"), dtype=hphi.dtype).pin_memory()
    lap.applyWithPotHost(hphi, hv, hout, chunk)
    ref = port.hpsi(lap_type, phi, v, ll)
    assert rel_inf(hout.numpy(), ref) <= TOL[dt]
    dout = torch.empty_like(dev(phi))
    lap.applyWithPot(dev(phi), dev(v), dout)
    assert bits_equal(hout.numpy(), host(dout))
    # pageable (unpinned) buffers work too
    out2 = torch.full(hphi.shape, float("nan"), dtype=hphi.dtype)
    lap.applyWithPotHost(torch.from_numpy(phi.copy()), torch.from_numpy(v.copy()), out2, chunk)
    assert bits_equal(out2.numpy(), hout.numpy())


def test_hamiltonian_cache_and_errors(H):
    """Hamiltonian::applyLocal recomputes only when the iterative indices
    change (src/Hamiltonian.cc:56-71); operators without applyWithPot are
    refused (src/pb/Lap.h:35-39)."""
    from mgmol_b200._lib import lib, MgbError
    dims, ll = (16, 16, 16), (4.0, 4.0, 4.0)
    grid = H.Grid(dims, ll, 1)
    phi = H.Orbitals(grid, 3, torch.float64, dev(synthetic_orbitals(3, dims)))
    ham = H.Hamiltonian()
    ham.setup(grid, 0)
    ham.potential(H.Potentials(dev(synthetic_potential(dims))))
    n0 = lib().mgb_launch_count()
    h1 = ham.applyLocal(phi)
    n1 = lib().mgb_launch_count()
    assert n1 > n0
    h2 = ham.applyLocal(phi)
    assert lib().mgb_launch_count() == n1 and h2 is h1
    phi.incrementIterativeIndex()
    ham.applyLocal(phi)
    assert lib().mgb_launch_count() > n1
    for bad in (1, 3, 4):
        lap = H.LapFactory.createLap(grid.with_ghosts(H.ghosts_for(bad)), bad)
        with pytest.raises(MgbError):
            lap.applyWithPot(phi.psi(), ham.potential().vtot(), torch.empty_like(phi.psi()))


# --------------------------------------------------------------------------
# GridFuncVector seam: bit-exact kernels on ghosted blocks
# --------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("g", [1, 2])
@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 0, 0), (1, 0, 1), (0, 1, 0), (0, 0, 1)])
def test_trade_boundaries_bit_exact(H, port, dt, g, bc):
    dims = (9, 12, 10)
    phi = synthetic_orbitals(3, dims, dt)
    grid = H.Grid(dims, (1.0, 1.0, 1.0), g, bc)
    gfv = H.GridFuncVector(grid, 3, TDT[dt])
    gfv.assign(dev(phi))
    gfv.trade_boundaries()
    assert bits_equal(host(gfv.data), port.trade_boundaries(phi, g, bc))
    back = torch.empty_like(dev(phi))
    gfv.getValues(back)
    ref = port.strip_ghosts(port.trade_boundaries(phi, g, bc), g)
    assert bits_equal(host(back), ref)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_fd_kernels_bit_exact(H, port, dt):
    dims, ll = (13, 9, 20), (3.1, 2.2, 4.7)
    h = tuple(l / d for l, d in zip(ll, dims))
    phi = synthetic_orbitals(4, dims, dt)
    for lap_type, g in ((0, 1), (1, 1), (2, 2)):
        grid = H.Grid(dims, ll, g)
        a = H.GridFuncVector(grid, 4, TDT[dt])
        b = H.GridFuncVector(grid, 4, TDT[dt])
        a.assign(dev(phi))
        a.applyLap(lap_type, b)
        ref = port.fdkernel(lap_type, port.trade_boundaries(phi, g), g, h)
        assert bits_equal(host(b.data), ref), lap_type
    grid = H.Grid(dims, ll, 1)
    a = H.GridFuncVector(grid, 4, TDT[dt])
    b = H.GridFuncVector(grid, 4, TDT[dt])
    a.assign(dev(phi))
    a.applyRHS(0, b)
    assert bits_equal(host(b.data), port.fdkernel(100, port.trade_boundaries(phi, 1), 1, h))


@pytest.mark.parametrize("dt,tag", [(np.float64, "f64"), (np.float32, "f32")])
def test_fd_kernels_against_golden(H, golden, dt, tag):
    """all six batched FD kernels, incl. 6th and 8th order, against the
    compiled reference's outputs."""
    dims = tuple(int(x) for x in golden["dims"])
    ll = tuple(float(x) for x in golden["ll"])
    N = int(golden["nfunc"])
    phi = synthetic_orbitals(N, dims, dt)
    for kind, g in ((0, 1), (1, 1), (2, 2), (3, 3), (4, 4)):
        grid = H.Grid(dims, ll, g)
        a = H.GridFuncVector(grid, N, TDT[dt])
        b = H.GridFuncVector(grid, N, TDT[dt])
        a.assign(dev(phi))
        a.applyLap(kind, b)
        assert bits_equal(host(b.data), golden["fd%d_%s" % (kind, tag)]), kind
        a.trade_boundaries()
        assert bits_equal(host(a.data), golden["trade_g%d_%s_bc111" % (g, tag)]) if g <= 2 else True


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_mg_transfer_bit_exact(H, port, dt):
    dims, g = (12, 8, 16), 1
    phi = synthetic_orbitals(3, dims, dt)
    grid = H.Grid(dims, (1.0, 1.0, 1.0), g)
    fine = H.GridFuncVector(grid, 3, TDT[dt])
    coarse = H.GridFuncVector(grid.coarse_grid(), 3, TDT[dt])
    fine.assign(dev(phi))
    fine.restrict3D(coarse)
    gv = port.trade_boundaries(phi, g)
    cref = port.restrict3D(gv, g)
    assert bits_equal(host(coarse.data), cref)
    coarse.set_updated_boundaries(False)
    fine.extend3D(coarse)
    assert bits_equal(host(fine.data), port.extend3D(cref, gv, g))


def test_mgkernels_reference_unit_test(H):
    """tests/testMGkernels.cc:8-85 through the GPU path."""
    g, nf = 2, 5
    grid = H.Grid((16, 16, 16), (1.0, 1.0, 1.0), g)
    const = np.stack([np.full((16, 16, 16), 1.0 + i) for i in range(nf)])
    fine = H.GridFuncVector(grid, nf)
    coarse = H.GridFuncVector(grid.coarse_grid(), nf)
    fine.assign(dev(const))
    fine.restrict3D(coarse)
    coarse.set_updated_boundaries(False)
    out = H.GridFuncVector(grid, nf)
    out.extend3D(coarse)
    res = torch.empty((nf, 16, 16, 16), dtype=torch.float64, device="cuda")
    out.getValues(res)
    for i in range(nf):
        assert np.abs(host(res)[i] - (1.0 + i)).max() < 1e-8


@pytest.mark.parametrize("lap_type,g", [(0, 1), (1, 1), (2, 2)])
def test_jacobi_bit_exact(H, port, lap_type, g):
    dims, ll = (12, 8, 16), (3.0, 2.0, 4.0)
    v = synthetic_orbitals(3, dims, np.float32)
    f = synthetic_orbitals(3, dims, np.float32, first=9)
    grid = H.Grid(dims, ll, g)
    lap = H.LapFactory.createLap(grid, lap_type)
    gv = H.GridFuncVector(grid, 3, torch.float32)
    gf = H.GridFuncVector(grid, 3, torch.float32)
    gw = H.GridFuncVector(grid, 3, torch.float32)
    gv.assign(dev(v))
    gf.assign(dev(f))
    gv.jacobi(lap_type, gf, gw, lap.jacobiFactor())
    rv, rw = port.jacobi(lap_type, port.add_ghosts(v, g), port.add_ghosts(f, g),
                         np.zeros_like(port.add_ghosts(v, g)), g, ll, lap.jacobiFactor())
    assert bits_equal(host(gv.data), rv)
    assert bits_equal(host(gw.data), rw)


# --------------------------------------------------------------------------
# multigrid preconditioner
# --------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("lap_type", [0, 2])
@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 0, 0)])
@pytest.mark.parametrize("levels", [1, 2])
def test_precond_mg_bit_exact(H, port, dt, lap_type, bc, levels):
    dims, ll, N = (16, 24, 32), (4.0, 6.0, 8.0), 3
    res = synthetic_orbitals(N, dims, dt)
    grid = H.Grid(dims, ll, H.ghosts_for(lap_type), bc)
    orb = H.Orbitals(grid, N, TDT[dt], dev(res))
    pc = H.OrbitalsPreconditioning()
    pc.setup(orb, levels, lap_type)
    pc.set_mode(1)  # the literal reference-shaped sequence
    pc.gamma_ = 0.31
    pc.precond_mg(orb)
    assert pc.last_mode() == 1
    ref = port.precond_mg(lap_type, levels, res, ll, 0.31, bc)
    assert bits_equal(host(orb.psi()), ref)
    # a second application reuses the resident work blocks
    pc.precond_mg(orb)
    assert bits_equal(host(orb.psi()), port.precond_mg(lap_type, levels, ref, ll, 0.31, bc))
    pc.close()


@pytest.mark.parametrize("dt,tag", [(np.float64, "f64"), (np.float32, "f32")])
def test_precond_against_golden(H, golden, dt, tag):
    dims = tuple(int(x) for x in golden["dims"])
    ll = tuple(float(x) for x in golden["ll"])
    N = int(golden["nfunc"])
    for lap_type in (0, 2):
        for bc in ((1, 1, 1), (0, 0, 0)):
            for lev in (1, 2):
                res = synthetic_orbitals(N, dims, dt)
                grid = H.Grid(dims, ll, H.ghosts_for(lap_type), bc)
                orb = H.Orbitals(grid, N, TDT[dt], dev(res))
                pc = H.OrbitalsPreconditioning()
                pc.setup(orb, lev, lap_type)
                pc.set_mode(1)
                pc.gamma_ = 0.37
                pc.precond_mg(orb)
                ref = golden["precond_lap%d_lev%d_%s_bc%d%d%d" % ((lap_type, lev, tag) + bc)]
                assert bits_equal(host(orb.psi()), ref)
                pc.close()
                # the fused kernels (float difference-form stencils): 1e-6 of the
                # per-function max norm, ten times inside the FP32 bar of 1e-5
                orb = H.Orbitals(grid, N, TDT[dt], dev(res))
                pc = H.OrbitalsPreconditioning()
                pc.setup(orb, lev, lap_type)
                pc.gamma_ = 0.37
                pc.precond_mg(orb)
                if pc.last_mode() == 2:
                    assert rel_inf(host(orb.psi()), ref) <= MG_TOL
                else:
                    assert bits_equal(host(orb.psi()), ref)
                pc.close()


# fused V-cycle vs the literal float V-cycle, relative to the max norm of the
# OUTPUT (which is several times smaller than the intermediate iterates the
# one-ulp differences are made on): observed 2e-7 .. 2e-6, bar 1e-5 (FP32)
MG_TOL = 5e-6


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("lap_type", [0, 2])
@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 0, 0)])
@pytest.mark.parametrize("levels,dims", [(0, (8, 8, 8)), (1, (16, 24, 32)), (2, (16, 24, 32)),
                                         (2, (32, 16, 64)), (1, (8, 12, 16))])
def test_precond_mg_fused(H, port, dt, lap_type, bc, levels, dims):
    """Fused V-cycle (one pass per Jacobi sweep, no-ghost blocks) against the
    oracle's literal Preconditioning<float>::mg."""
    ll, N = tuple(0.25 * d for d in dims), 5
    res = synthetic_orbitals(N, dims, dt)
    grid = H.Grid(dims, ll, H.ghosts_for(lap_type), bc)
    orb = H.Orbitals(grid, N, TDT[dt], dev(res))
    pc = H.OrbitalsPreconditioning()
    pc.setup(orb, levels, lap_type)
    pc.set_mode(2)
    pc.gamma_ = 0.31
    pc.precond_mg(orb)
    assert pc.last_mode() == 2
    ref = port.precond_mg(lap_type, levels, res, ll, 0.31, bc)
    assert rel_inf(host(orb.psi()), ref) <= MG_TOL
    pc.precond_mg(orb)  # resident work blocks are reused
    ref2 = port.precond_mg(lap_type, levels, ref, ll, 0.31, bc)
    assert rel_inf(host(orb.psi()), ref2) <= MG_TOL
    pc.close()


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("lap_type", [0, 2])
@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 0, 0)])
@pytest.mark.parametrize("zboxes", [False, True])
@pytest.mark.parametrize("levels,dims", [(1, (16, 24, 32)), (2, (16, 24, 32)), (2, (32, 32, 64))])
def test_precond_mg_fused_3d_self_neighbours(H, dt, lap_type, bc, zboxes, levels, dims):
    """The any-decomposition variants of the fused V-cycle kernels (Jacobi with the neighbour
    table and z-halo column arrays, restriction and prolongation through the 8 neighbour
    blocks) on one rank, with the box itself as every neighbour: bit-identical to the
    single-rank fused kernels."""
    ll = (4.0, 6.0, 8.0)
    N = 3
    res = synthetic_orbitals(N, dims, dt)
    grid = H.Grid(dims, ll, H.ghosts_for(lap_type), bc)

    def run():
        r = H.Orbitals(grid, N, TDT[dt], dev(res))
        pc = H.OrbitalsPreconditioning()
        pc.setup(r, levels, lap_type)
        pc.set_mode(2)
        pc.gamma_ = 0.3
        pc.precond_mg(r)
        pc.close()
        return r.psi()

    ref = run()
    os.environ["MGB_MG_FORCE_3D"] = "1"
    if zboxes:
        os.environ["MGB_HPSI_FORCE_ZBOXES"] = "1"
    try:
        got = run()
    finally:
        os.environ.pop("MGB_MG_FORCE_3D", None)
        os.environ.pop("MGB_HPSI_FORCE_ZBOXES", None)
    assert torch.equal(got, ref)


@pytest.mark.parametrize("bc", [(1, 0, 1), (0, 1, 1), (1, 1, 0)])
def test_precond_mg_mixed_bc_stays_literal(H, port, bc):
    """Mixed periodic/Dirichlet boxes: the reference's result depends on ghost
    values its zeroing skips, so these boxes run the literal sequence
    (bit-identical) and refuse the fused mode."""
    from mgmol_b200._lib import MgbError
    dims, ll, N = (16, 24, 32), (4.0, 6.0, 8.0), 3
    res = synthetic_orbitals(N, dims, np.float32)
    grid = H.Grid(dims, ll, 1, bc)
    orb = H.Orbitals(grid, N, torch.float32, dev(res))
    pc = H.OrbitalsPreconditioning()
    pc.setup(orb, 2, 0)
    with pytest.raises(MgbError):
        pc.set_mode(2)
    pc.gamma_ = 0.31
    pc.precond_mg(orb)
    assert pc.last_mode() == 1
    assert bits_equal(host(orb.psi()), port.precond_mg(0, 2, res, ll, 0.31, bc))
    pc.close()


@pytest.mark.parametrize("cfg", ["4,1,1,5,0", "2,2,2,4,5", "4,2,3,6,7", "2,4,1,7,4"])
@pytest.mark.parametrize("lap_type", [0, 2])
def test_precond_mg_fused_tile_configs(H, port, lap_type, cfg):
    """Every tile shape of the Jacobi kernel (rows per thread, row groups,
    functions per CTA, ring depth, x chunk) gives the same numbers."""
    dims, ll, N = (16, 16, 32), (4.0, 4.0, 8.0), 4
    res = synthetic_orbitals(N, dims, np.float32)
    ref = port.precond_mg(lap_type, 1, res, ll, 0.31, (1, 1, 1))
    os.environ["MGB_MG_CFG"] = cfg
    try:
        grid = H.Grid(dims, ll, H.ghosts_for(lap_type))
        orb = H.Orbitals(grid, N, torch.float32, dev(res))
        pc = H.OrbitalsPreconditioning()
        pc.setup(orb, 1, lap_type)
        pc.set_mode(2)
        pc.gamma_ = 0.31
        pc.precond_mg(orb)
        assert rel_inf(host(orb.psi()), ref) <= MG_TOL
        pc.close()
    finally:
        del os.environ["MGB_MG_CFG"]


@pytest.mark.parametrize("lap_type", [0, 2])
def test_precond_mg_fused_equals_literal_full_size(H, lap_type):
    """128^3 (H2O_64's grid): fused against the literal path, which the small
    tests pin bit-for-bit to the reference."""
    n, N = 128, 3
    grid = H.Grid((n, n, n), (23.5,) * 3, H.ghosts_for(lap_type))
    g = torch.Generator(device="cuda").manual_seed(5)
    res = torch.rand((N, n, n, n), generator=g, device="cuda", dtype=torch.float32) - 0.5
    x = torch.arange(n, device="cuda", dtype=torch.float32) / n
    res += torch.cos(2 * np.pi * x)[None, :, None, None] * torch.sin(4 * np.pi * x)[None, None, None, :]
    outs = {}
    for mode in (1, 2):
        orb = H.Orbitals(grid, N, torch.float32, res.clone())
        pc = H.OrbitalsPreconditioning()
        pc.setup(orb, 2, lap_type)
        pc.set_mode(mode)
        pc.gamma_ = 0.2
        pc.precond_mg(orb)
        assert pc.last_mode() == mode
        outs[mode] = orb.psi().double()
        pc.close()
    scale = outs[1].abs().amax(dim=(1, 2, 3), keepdim=True)
    assert float(((outs[1] - outs[2]).abs() / scale).max()) <= MG_TOL


def test_precond_errors(H):
    from mgmol_b200._lib import MgbError
    grid = H.Grid((12, 12, 12), (1.0, 1.0, 1.0), 1)
    orb = H.Orbitals(grid, 2)
    pc = H.OrbitalsPreconditioning()
    with pytest.raises(MgbError):  # 12 is not divisible by 2^3
        pc.setup(orb, 3, 0)
    with pytest.raises(MgbError):  # applyLap has no case 10
        H.OrbitalsPreconditioning().setup(orb, 1, 10)


# --------------------------------------------------------------------------
# dense contractions
# --------------------------------------------------------------------------
# 3xTF32 with chunked accumulation against the exact contraction, relative to
# |a_i| |b_j| (observed <= 1e-6; north star FP32 bar 1e-5)
F32_TC_TOL = 3e-6


def _exact_tn(a, b, alpha):
    a2 = a.reshape(a.shape[0], -1).astype(np.float64)
    b2 = b.reshape(b.shape[0], -1).astype(np.float64)
    return alpha * (a2 @ b2.T)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("N,dims", [(5, (6, 4, 8)), (37, (10, 12, 14)), (130, (16, 16, 24)),
                                    (256, (32, 32, 16))])
def test_gram_and_projection(H, port, dt, N, dims):
    a = synthetic_orbitals(N, dims, dt)
    b = synthetic_orbitals(N, dims, dt, first=1000)
    grid = H.Grid(dims, (2.0, 2.0, 2.0), 1)
    A = H.Orbitals(grid, N, TDT[dt], dev(a))
    B = H.Orbitals(grid, N, TDT[dt], dev(b))
    from mgmol_b200._lib import lib, check
    K = a[0].size
    eps = np.finfo(np.float64).eps
    ex = _exact_tn(a, a, grid.vel())
    na = np.sqrt(np.diag(ex))
    exp = _exact_tn(a, b, grid.vel())
    nb = np.sqrt(np.diag(_exact_tn(b, b, grid.vel())))
    # float operands: mode 0 = 3xTF32 tensor tiles on tcgen05 (F32_TC_TOL of |a||b|, bar
    # 1e-5), mode 2 = the same arithmetic on mma.sync, mode 1 = DMMA on widened operands
    # (double products and sums)
    for mode in ((0, 2, 1) if dt == np.float32 else (0,)):
        check(lib().mgb_set_f32_contraction(mode))
        try:
            rel = F32_TC_TOL if (dt == np.float32 and mode != 1) else 4 * K * eps
            S = host(A.computeGram())
            assert (np.abs(S - ex) <= rel * np.outer(na, na) + 1e-300).all()
            assert np.array_equal(S, S.T), "Gram must be exactly symmetric"
            P = host(A.computeLocalProduct(B))
            assert (np.abs(P - exp) <= rel * np.outer(na, nb) + 1e-300).all()
            if N <= 40:
                # tie-breaker: the reference's own double-accumulating loops
                ref = port.gemm_tn(a, b, grid.vel())
                assert np.abs(P - ref).max() <= rel * np.abs(np.outer(na, nb)).max()
        finally:
            check(lib().mgb_set_f32_contraction(0))


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("N,n,dims", [(5, 5, (6, 4, 8)), (37, 37, (10, 12, 14)),
                                      (130, 70, (16, 16, 24)), (256, 256, (32, 16, 16))])
def test_multiply_by_matrix(H, port, dt, N, n, dims):
    a = synthetic_orbitals(N, dims, dt)
    M = np.random.default_rng(7).standard_normal((N, n)) / np.sqrt(N)
    grid = H.Grid(dims, (2.0, 2.0, 2.0), 1)
    A = H.Orbitals(grid, N, TDT[dt], dev(a))
    out = H.Orbitals(grid, n, TDT[dt])
    A.multiplyByMatrix(dev(M), out)
    got = host(out.psi())
    ex = np.einsum("lj,lxyz->jxyz", M, a.astype(np.float64))
    scale = np.abs(ex).max()
    # float: 3xTF32 tensor tiles with float coefficients (bar 1e-5)
    tol = 1e-13 if dt == np.float64 else 2e-6
    assert np.abs(got - ex).max() <= tol * scale * max(1, N / 64)
    if N <= 40:
        ref = port.gemm_nn(a, M)
        assert np.abs(got.astype(np.float64) - ref).max() <= tol * scale
    if dt == np.float32:
        # the DMMA path on widened operands reproduces the reference's sums
        from mgmol_b200._lib import lib, check
        check(lib().mgb_set_f32_contraction(1))
        try:
            A.multiplyByMatrix(dev(M), out)
            assert np.abs(host(out.psi()) - ex).max() <= 2e-7 * scale * max(1, N / 64)
        finally:
            check(lib().mgb_set_f32_contraction(0))


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("N,dims", [(300, (32, 32, 32)), (520, (16, 32, 64)), (129, (64, 64, 16))])
def test_contractions_many_tiles_streamk(H, dt, N, dims):
    """Several 128-wide tiles with ragged edges and a K range long enough to
    be shared between CTAs (stream-K segments, half-work diagonal Gram tiles,
    fix-up): K*eps-scaled bound against the exact FP64 contraction, exact
    symmetry, run-to-run determinism, and beta accumulation."""
    from mgmol_b200._lib import lib, check
    a = synthetic_orbitals(N, dims, dt)
    b = synthetic_orbitals(N, dims, dt, first=2000)
    grid = H.Grid(dims, (2.0, 2.0, 2.0), 1)
    A = H.Orbitals(grid, N, TDT[dt], dev(a))
    B = H.Orbitals(grid, N, TDT[dt], dev(b))
    K = a[0].size
    eps = np.finfo(np.float64).eps
    rel = F32_TC_TOL if dt == np.float32 else 4 * K * eps
    ex = _exact_tn(a, a, grid.vel())
    na = np.sqrt(np.diag(ex))
    S = host(A.computeGram())
    assert (np.abs(S - ex) <= rel * np.outer(na, na) + 1e-300).all()
    assert np.array_equal(S, S.T)
    assert bits_equal(S, host(A.computeGram())), "deterministic summation order"
    exp = _exact_tn(a, b, grid.vel())
    nb = np.sqrt(np.diag(_exact_tn(b, b, grid.vel())))
    bound = rel * np.outer(na, nb) + 1e-300
    P = host(A.computeLocalProduct(B))
    assert (np.abs(P - exp) <= bound).all()
    assert bits_equal(P, host(A.computeLocalProduct(B)))
    # C = alpha A^T B + beta C
    C0 = np.random.default_rng(5).standard_normal((N, N))
    c = dev(C0.T.copy())  # column-major C(i, j)
    check(lib().mgb_gemm_tn(0 if dt == np.float32 else 1, N, N, K, 0.5 * grid.vel(),
                            A.psi().data_ptr(), K, B.psi().data_ptr(), K, -2.0,
                            c.data_ptr(), N, None))
    got = host(c).T
    assert (np.abs(got - (0.5 * exp - 2.0 * C0)) <= bound + 4 * eps * np.abs(C0)).all()


@pytest.mark.parametrize("N,K,lda", [(128, 4096, 4096), (136, 4000, 4000), (130, 6144, 6148),
                                     (256, 32768, 32768), (304, 65536, 65600), (37, 1680, 1680),
                                     (8, 64, 64), (264, 2080, 2080)])
@pytest.mark.parametrize("positive", [False, True])
def test_f32_contractions_tcgen05(N, K, lda, positive):
    """k_gemm_tn_umma (tcgen05 kind::tf32, A from TMEM, TMA boxes of two K slabs) through
    mgb_syrk_t / mgb_gemm_tn: shapes that take the 5-D maps (rows % 8 == 0, K % 32 == 0) and
    shapes that take one slab per box, ragged tiles, a padded leading dimension, all-positive
    operands (the tensor core's truncating FP32 sums bias those), against the exact FP64
    contraction and against the mma.sync 3xTF32 kernel; exact symmetry, run-to-run determinism,
    beta accumulation.  Arithmetic of the reference: src/linear_algebra/mputils.cc:848-948."""
    from mgmol_b200._lib import lib, check
    g = torch.Generator(device="cuda").manual_seed(1000 + N)
    a = torch.rand((N, lda), generator=g, device="cuda", dtype=torch.float32)
    b = torch.rand((N, lda), generator=g, device="cuda", dtype=torch.float32)
    if not positive:
        a -= 0.5
        b -= 0.5
    ad, bd = a[:, :K].double(), b[:, :K].double()
    exg, exp = ad @ ad.t(), ad @ bd.t()
    na, nb = torch.sqrt(torch.diag(exg)), torch.sqrt((bd * bd).sum(1))
    S = torch.full((N, N), float("nan"), device="cuda", dtype=torch.float64)
    P = torch.full((N, N), float("nan"), device="cuda", dtype=torch.float64)
    res = {}
    for mode in (0, 2):
        check(lib().mgb_set_f32_contraction(mode))
        try:
            check(lib().mgb_syrk_t(0, N, K, 1.0, a.data_ptr(), lda, S.data_ptr(), N, None))
            check(lib().mgb_gemm_tn(0, N, N, K, 1.0, a.data_ptr(), lda, b.data_ptr(), lda, 0.0,
                                    P.data_ptr(), N, None))
            torch.cuda.synchronize()
            # column-major C(i, j): the row-major torch view is C^T
            assert bool(((S.t() - exg).abs() <= F32_TC_TOL * torch.outer(na, na)).all()), mode
            assert bool(((P.t() - exp).abs() <= F32_TC_TOL * torch.outer(na, nb)).all()), mode
            assert torch.equal(S, S.t())
            res[mode] = (S.clone(), P.clone())
            if mode == 0:
                S2, P2 = torch.empty_like(S), torch.empty_like(P)
                check(lib().mgb_syrk_t(0, N, K, 1.0, a.data_ptr(), lda, S2.data_ptr(), N, None))
                check(lib().mgb_gemm_tn(0, N, N, K, 1.0, a.data_ptr(), lda, b.data_ptr(), lda, 0.0,
                                        P2.data_ptr(), N, None))
                assert torch.equal(S, S2) and torch.equal(P, P2), "deterministic summation order"
                C0 = torch.randn((N, N), generator=g, device="cuda", dtype=torch.float64)
                Cb = C0.clone()
                check(lib().mgb_gemm_tn(0, N, N, K, 0.5, a.data_ptr(), lda, b.data_ptr(), lda, -2.0,
                                        Cb.data_ptr(), N, None))
                want = 0.5 * exp.t() - 2.0 * C0
                bound = F32_TC_TOL * torch.outer(nb, na) + 1e-15 * C0.abs()
                assert bool(((Cb - want).abs() <= bound).all())
        finally:
            check(lib().mgb_set_f32_contraction(0))
    # both kernels implement the same 3xTF32 arithmetic
    assert bool(((res[0][1] - res[2][1]).abs() <= 2 * F32_TC_TOL * torch.outer(nb, na)).all())


@pytest.mark.parametrize("m,n,k", [(4096, 128, 128), (5000, 130, 130), (32768, 256, 256),
                                   (8192, 70, 130), (1159, 264, 200), (640, 8, 8)])
@pytest.mark.parametrize("positive", [False, True])
def test_f32_multiply_tcgen05(m, n, k, positive):
    """k_gemm_nn_umma (Phi tile in TMEM with the points on the lanes, coefficient tiles by
    TMA) through mgb_gemm_nn: ragged point / orbital tiles, k not a multiple of 32, beta, against
    the exact FP64 product and the mma.sync 3xTF32 kernel (same arithmetic), deterministic.
    ExtendedGridOrbitals::multiplyByMatrix (src/ExtendedGridOrbitals.cc:448-498)."""
    from mgmol_b200._lib import lib, check
    g = torch.Generator(device="cuda").manual_seed(50 + n)
    a = torch.rand((k, m), generator=g, device="cuda", dtype=torch.float32)
    mc = torch.rand((n, k), generator=g, device="cuda", dtype=torch.float64)  # column-major k x n
    if not positive:
        a -= 0.5
        mc -= 0.5
    out0 = torch.rand((n, m), generator=g, device="cuda", dtype=torch.float32)
    alpha, beta = 0.75, -0.5
    ex = alpha * (mc @ a.double()) + beta * out0.double()
    scale = alpha * (mc.abs() @ a.double().abs()) + out0.double().abs()
    got = {}
    for mode in (0, 2):
        check(lib().mgb_set_f32_contraction(mode))
        try:
            out = out0.clone()
            check(lib().mgb_gemm_nn(0, m, n, k, alpha, a.data_ptr(), m, mc.data_ptr(), k, beta,
                                    out.data_ptr(), m, None))
            torch.cuda.synchronize()
            # 3xTF32 products + the rounding of the float result
            assert bool(((out.double() - ex).abs() <= F32_TC_TOL * scale).all()), mode
            got[mode] = out
            if mode == 0:
                out2 = out0.clone()
                check(lib().mgb_gemm_nn(0, m, n, k, alpha, a.data_ptr(), m, mc.data_ptr(), k, beta,
                                        out2.data_ptr(), m, None))
                assert torch.equal(out, out2), "deterministic"
        finally:
            check(lib().mgb_set_f32_contraction(0))
    assert bool(((got[0].double() - got[2].double()).abs() <= 2 * F32_TC_TOL * scale).all())


@pytest.mark.parametrize("N,ks,nslabs", [(16, 512, 3), (136, 2048, 2), (8, 96, 5), (12, 480, 4)])
def test_f32_slab_contractions_tcgen05(N, ks, nslabs):
    """mgb_syrk_t_slabs / mgb_gemm_tn_slabs (LocGridOrbitals: one contraction per x-slab, all slabs
    in one launch, src/LocGridOrbitals.cc:1504-1604) for float blocks on the tcgen05 kernel: the
    slab is the last index of the 5-D tensor maps (rows % 8 == 0 and K % 32 == 0) or of the 3-D
    maps (any shape)."""
    from mgmol_b200._lib import lib, check
    g = torch.Generator(device="cuda").manual_seed(77 + N)
    lda = ks * nslabs
    a = torch.rand((N, lda), generator=g, device="cuda", dtype=torch.float32) - 0.3
    b = torch.rand((N, lda), generator=g, device="cuda", dtype=torch.float32) - 0.6
    S = torch.full((nslabs, N, N), float("nan"), device="cuda", dtype=torch.float64)
    P = torch.full((nslabs, N, N), float("nan"), device="cuda", dtype=torch.float64)
    check(lib().mgb_syrk_t_slabs(0, N, ks, nslabs, 2.0, a.data_ptr(), lda, S.data_ptr(), N, None))
    check(lib().mgb_gemm_tn_slabs(0, N, N, ks, nslabs, 2.0, a.data_ptr(), lda, b.data_ptr(), lda, 0.0,
                                  P.data_ptr(), N, None))
    torch.cuda.synchronize()
    for i in range(nslabs):
        ad = a[:, i * ks:(i + 1) * ks].double()
        bd = b[:, i * ks:(i + 1) * ks].double()
        na, nb = torch.sqrt((ad * ad).sum(1)), torch.sqrt((bd * bd).sum(1))
        assert bool(((S[i].t() - 2.0 * ad @ ad.t()).abs() <= 2.0 * F32_TC_TOL * torch.outer(na, na)).all())
        assert torch.equal(S[i], S[i].t())
        assert bool(((P[i].t() - 2.0 * ad @ bd.t()).abs() <= 2.0 * F32_TC_TOL * torch.outer(na, nb)).all())


@pytest.mark.parametrize("m,n,K", [(264, 136, 4096), (8, 392, 2048), (130, 70, 6144), (384, 128, 65536)])
def test_f32_rectangular_projection_tcgen05(m, n, K):
    """mgb_gemm_tn with different row counts for A and B (computeLocalProduct against a block of
    other columns, addDotWithNcol2Matrix, src/ExtendedGridOrbitals.cc:1704-1752) on the tcgen05
    kernel, beta accumulation included."""
    from mgmol_b200._lib import lib, check
    g = torch.Generator(device="cuda").manual_seed(m + n)
    a = torch.rand((m, K), generator=g, device="cuda", dtype=torch.float32) - 0.4
    b = torch.rand((n, K), generator=g, device="cuda", dtype=torch.float32) - 0.6
    C0 = torch.randn((n, m), generator=g, device="cuda", dtype=torch.float64)  # column-major m x n
    C = C0.clone()
    check(lib().mgb_gemm_tn(0, m, n, K, 0.5, a.data_ptr(), K, b.data_ptr(), K, 1.5, C.data_ptr(), m, None))
    torch.cuda.synchronize()
    ad, bd = a.double(), b.double()
    want = 0.5 * (bd @ ad.t()) + 1.5 * C0
    na, nb = torch.sqrt((ad * ad).sum(1)), torch.sqrt((bd * bd).sum(1))
    assert bool(((C - want).abs() <= F32_TC_TOL * torch.outer(nb, na) + 1e-15 * C0.abs()).all())


def test_f32_contractions_tcgen05_full_size():
    """128^3 x 256 float (the ORBDTYPE float shape of H2O_64): Gram and Phi^T (H Phi) on the
    tcgen05 kernel against FP64 cuBLAS on the widened operands."""
    from mgmol_b200._lib import lib, check
    N, K = 256, 128 ** 3
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.rand((N, K), generator=g, device="cuda", dtype=torch.float32) - 0.5
    b = torch.rand((N, K), generator=g, device="cuda", dtype=torch.float32) - 0.25
    exg = torch.zeros((N, N), device="cuda", dtype=torch.float64)
    exp = torch.zeros((N, N), device="cuda", dtype=torch.float64)
    nb2 = torch.zeros(N, device="cuda", dtype=torch.float64)
    step = 1 << 19
    for k0 in range(0, K, step):
        ad, bd = a[:, k0:k0 + step].double(), b[:, k0:k0 + step].double()
        exg += ad @ ad.t()
        exp += ad @ bd.t()
        nb2 += (bd * bd).sum(1)
    na, nb = torch.sqrt(torch.diag(exg)), torch.sqrt(nb2)
    S = torch.empty((N, N), device="cuda", dtype=torch.float64)
    P = torch.empty((N, N), device="cuda", dtype=torch.float64)
    check(lib().mgb_syrk_t(0, N, K, 1.0, a.data_ptr(), K, S.data_ptr(), N, None))
    check(lib().mgb_gemm_tn(0, N, N, K, 1.0, a.data_ptr(), K, b.data_ptr(), K, 0.0, P.data_ptr(), N,
                            None))
    torch.cuda.synchronize()
    assert bool(((S.t() - exg).abs() <= F32_TC_TOL * torch.outer(na, na)).all())
    assert bool(((P.t() - exp).abs() <= F32_TC_TOL * torch.outer(na, nb)).all())
    assert torch.equal(S, S.t())
    # Phi M over the block written by nothing else: b <- a M
    mc = torch.rand((N, N), generator=g, device="cuda", dtype=torch.float64) - 0.5
    check(lib().mgb_gemm_nn(0, K, N, N, 1.0, a.data_ptr(), K, mc.data_ptr(), N, 0.0, b.data_ptr(), K,
                            None))
    torch.cuda.synchronize()
    worst = 0.0
    for k0 in range(0, K, step):
        ad = a[:, k0:k0 + step].double()
        exm = mc @ ad
        scm = mc.abs() @ ad.abs()
        worst = max(worst, float(((b[:, k0:k0 + step].double() - exm).abs() / scm).max()))
    assert worst <= F32_TC_TOL


def test_contractions_full_size_against_cublas(H):
    """H2O_64 shape (128^3 x 256, FP64): Gram, Phi^T H Phi-shaped product and
    Phi M against cuBLAS DGEMM on the same operands, K*eps-scaled."""
    n, N = 128, 256
    K = n ** 3
    g = torch.Generator(device="cuda").manual_seed(11)
    phi = torch.rand((N, n, n, n), generator=g, device="cuda", dtype=torch.float64) - 0.5
    grid = H.Grid((n, n, n), (23.5, 23.5, 23.5), 2)
    A = H.Orbitals(grid, N, torch.float64, phi)
    S = A.computeGram()
    a2 = phi.view(N, -1)
    ref = grid.vel() * (a2 @ a2.t())
    eps = np.finfo(np.float64).eps
    nrm = torch.sqrt(torch.diag(ref))
    bound = 4 * K * eps * torch.outer(nrm, nrm)
    assert bool(((S - ref).abs() <= bound).all())
    assert torch.equal(S, S.t())
    hphi = torch.roll(phi, 1, dims=0).contiguous()
    P = A.computeLocalProduct(hphi)
    refp = grid.vel() * (a2 @ hphi.view(N, -1).t())
    assert bool(((P - refp).abs() <= bound.max()).all())
    M = torch.randn((N, N), generator=g, device="cuda", dtype=torch.float64) / np.sqrt(N)
    out = H.Orbitals(grid, N, torch.float64)
    A.multiplyByMatrix(M, out)
    refm = (M.t() @ a2).view_as(phi)
    assert float((out.psi() - refm).abs().max()) <= 1e-13 * float(refm.abs().max()) * (N / 64)


# --------------------------------------------------------------------------
# residual assembly
# --------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("lap_type", [0, 2])
@pytest.mark.parametrize("N,dims,bcolor", [(5, (12, 8, 16), 2), (37, (10, 12, 14), 32),
                                           (130, (16, 16, 24), 64)])
def test_compute_matB_and_addHlocal2matrix(H, port, dt, lap_type, N, dims, bcolor):
    """ExtendedGridOrbitals::computeMatB (src/ExtendedGridOrbitals.cc:901-967): matB = vel *
    Phi^T (B Phi), B the Mehrstellen right-hand-side stencil (identity for the 4th-order
    operator) applied in blocks of bcolor columns; Hamiltonian::addHlocal2matrix
    (src/Hamiltonian.cc:163-212): hij += Phi^T H_loc Phi."""
    ll = (3.0, 2.5, 4.0)
    phi = synthetic_orbitals(N, dims, dt)
    v = synthetic_potential(dims)
    grid = H.Grid(dims, ll, H.ghosts_for(lap_type))
    orb = H.Orbitals(grid, N, TDT[dt], dev(phi))
    lap = H.LapFactory.createLap(grid, lap_type)
    got = host(orb.computeMatB(lap, bcolor=bcolor))
    bphi = port.lap_rhs(lap_type, phi, ll)
    f = phi.reshape(N, -1).astype(np.float64)
    exact = grid.vel() * f @ bphi.reshape(N, -1).astype(np.float64).T
    scale = grid.vel() * np.abs(f) @ np.abs(bphi.reshape(N, -1).astype(np.float64)).T
    tol = 1e-13 if dt == np.float64 else 3e-6
    assert (np.abs(got - exact) / scale).max() <= tol
    # with a work block: one pass
    work = H.Orbitals(grid, N, TDT[dt])
    got2 = host(orb.computeMatB(lap, work=work))
    assert (np.abs(got2 - exact) / scale).max() <= tol
    ham = H.Hamiltonian()
    ham.setup(grid, lap_type)
    ham.potential(H.Potentials(dev(v)))
    hij = torch.zeros((N, N), dtype=torch.float64, device="cuda")
    ham.addHlocal2matrix(orb, orb, hij)
    ham.addHlocal2matrix(orb, orb, hij, force=True)
    once = host(ham.addHlocalij(orb))
    assert np.abs(host(hij) - 2 * once).max() <= 1e-12 * np.abs(once).max()


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 0, 0)])
@pytest.mark.parametrize("dims", [(12, 8, 16), (7, 9, 10), (32, 32, 32)])
def test_apply_b_bit_exact(H, port, dt, bc, dims):
    """Lap::rhs on a no-ghost block = FDkernelRHS_4th_Mehr1 after
    setDataWithGhosts + trade_boundaries, bit for bit."""
    N, ll = 3, tuple(0.3 * d for d in dims)
    phi = synthetic_orbitals(N, dims, dt)
    grid = H.Grid(dims, ll, 1, bc)
    out = torch.full((N,) + dims, float("nan"), dtype=TDT[dt], device="cuda")
    H.LapFactory.createLap(grid, 0).rhs(dev(phi), out)
    h = tuple(l / d for l, d in zip(ll, dims))
    ref = port.fdkernel(100, port.trade_boundaries(phi, 1, bc), 1, h, rhs_ghosts=0)
    assert bits_equal(host(out), ref)
    # Laph4MP: B2 = 2/3 + faces / 36 + edges / 72 (rhs_4th_Mehr2)
    out10 = torch.full_like(out, float("nan"))
    H.LapFactory.createLap(grid, 10).rhs(dev(phi), out10)
    assert bits_equal(host(out10), port.lap_rhs(10, phi, ll, bc))
    # non-compact operators: B = 1
    out2 = torch.empty_like(out)
    H.LapFactory.createLap(H.Grid(dims, ll, 2, bc), 2).rhs(dev(phi), out2)
    assert bits_equal(host(out2), phi)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("lap_type", [0, 2])
@pytest.mark.parametrize("N,dims", [(5, (12, 8, 16)), (37, (10, 12, 14)), (130, (16, 16, 24))])
def test_residual_using_hphi(H, port, dt, lap_type, N, dims):
    """res = (B psi) theta - H psi  (MGmol::computeResidualUsingHPhi) against the
    oracle's sequence: B per orbital, MPgemmNN, axpy(-1)."""
    ll = tuple(0.3 * d for d in dims)
    phi = synthetic_orbitals(N, dims, dt)
    v = synthetic_potential(dims)
    theta = np.random.default_rng(3).standard_normal((N, N)) / np.sqrt(N)
    theta = 0.5 * (theta + theta.T)
    grid = H.Grid(dims, ll, H.ghosts_for(lap_type))
    psi = H.Orbitals(grid, N, TDT[dt], dev(phi))
    ham = H.Hamiltonian()
    ham.setup(grid, lap_type)
    ham.potential(H.Potentials(dev(v)))
    hphi = ham.applyLocal(psi)
    res = H.Orbitals(grid, N, TDT[dt])
    H.computeResidualUsingHPhi(ham.lapOper(), psi, hphi, dev(theta), res)
    got = host(res.psi())
    hp = host(hphi.psi())
    if lap_type == 0:
        h = tuple(l / d for l, d in zip(ll, dims))
        bphi = port.fdkernel(100, port.trade_boundaries(phi, 1), 1, h, rhs_ghosts=0)
    else:
        bphi = phi
    ref = port.gemm_nn(bphi, theta) - hp          # MPgemmNN then axpy(-1., hphi)
    scale = np.abs(ref).max()
    tol = 1e-13 if dt == np.float64 else 2e-6
    assert np.abs(got.astype(np.float64) - ref).max() <= tol * scale * max(1, N / 64)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("N,dims", [(5, (12, 8, 16)), (37, (10, 12, 14)), (130, (24, 24, 24))])
def test_density_blas3(H, port, dt, N, dims):
    """rho += sum_j (Phi X)_j phi_j (Rho::computeRhoSubdomainUsingBlas3): the
    oracle's MPgemmNN product, then the reference's j-ordered accumulation with
    the product formed in ORBDTYPE."""
    phi = synthetic_orbitals(N, dims, dt)
    X = np.random.default_rng(12).standard_normal((N, N)) / N
    X = 0.5 * (X + X.T)
    grid = H.Grid(dims, (3.0, 2.5, 4.0), 1)
    A = H.Orbitals(grid, N, TDT[dt], dev(phi))
    rho0 = np.random.default_rng(13).uniform(0, 1, dims)
    rho = dev(rho0.copy())
    H.computeRhoUsingBlas3(A, dev(X), rho)
    product = port.gemm_nn(phi, X)
    ref = rho0.copy()
    for j in range(N):
        ref += (product[j] * phi[j]).astype(np.float64)
    exact = rho0 + np.einsum("ixyz,ij,jxyz->xyz", phi.astype(np.float64), X, phi.astype(np.float64))
    scale = np.abs(exact - rho0).max() + 1.0
    tol = 1e-13 if dt == np.float64 else 3e-6
    assert np.abs(host(rho) - ref).max() <= tol * scale * max(1, N / 64)
    assert np.abs(host(rho) - exact).max() <= (1e-12 if dt == np.float64 else 1e-5) * scale


@pytest.mark.parametrize("dt,tag", [(np.float64, "f64"), (np.float32, "f32")])
def test_f1_rows_against_golden(H, dt, tag):
    """B / B2, residual assembly and density against vectors produced by the
    compiled reference (tests/golden/make_golden_f1.py)."""
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                  "reference_f1.npz")))
    dims = tuple(int(x) for x in g["dims"])
    ll = tuple(float(x) for x in g["ll"])
    N = int(g["nfunc"])
    theta = g["theta"]
    phi = synthetic_orbitals(N, dims, dt)
    v = synthetic_potential(dims)
    for lt in (0, 10):
        for bc in ((1, 1, 1), (0, 0, 0)):
            out = torch.full((N,) + dims, float("nan"), dtype=TDT[dt], device="cuda")
            H.LapFactory.createLap(H.Grid(dims, ll, 1, bc), lt).rhs(dev(phi), out)
            assert bits_equal(host(out), g["rhs_lap%d_%s_bc%d%d%d" % ((lt, tag) + bc)])
    tol = 1e-13 if dt == np.float64 else 2e-6
    for lt in (0, 2):
        grid = H.Grid(dims, ll, H.ghosts_for(lt))
        psi = H.Orbitals(grid, N, TDT[dt], dev(phi))
        ham = H.Hamiltonian()
        ham.setup(grid, lt)
        ham.potential(H.Potentials(dev(v)))
        res = H.Orbitals(grid, N, TDT[dt])
        H.computeResidualUsingHPhi(ham.lapOper(), psi, ham.applyLocal(psi), dev(theta), res)
        ref = g["residual_lap%d_%s" % (lt, tag)].astype(np.float64)
        assert np.abs(host(res.psi()) - ref).max() <= tol * np.abs(ref).max()
    grid = H.Grid(dims, ll, 1)
    rho = torch.zeros(dims, dtype=torch.float64, device="cuda")
    H.computeRhoUsingBlas3(H.Orbitals(grid, N, TDT[dt], dev(phi)), dev(theta), rho)
    ref = g["rho_%s" % tag]
    assert np.abs(host(rho) - ref).max() <= tol * np.abs(ref).max()


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_diagonal_dot_products(H, port, dt):
    """computeDiagonalElementsDotProduct: one launch for all orbitals, double
    accumulation like MPdot."""
    dims, N = (12, 10, 16), 7
    a = synthetic_orbitals(N, dims, dt)
    b = synthetic_orbitals(N, dims, dt, first=50)
    grid = H.Grid(dims, (3.0, 2.5, 4.0), 1)
    A = H.Orbitals(grid, N, TDT[dt], dev(a))
    B = H.Orbitals(grid, N, TDT[dt], dev(b))
    ss = host(A.computeDiagonalElementsDotProduct(B))
    ex = grid.vel() * np.einsum("ixyz,ixyz->i", a.astype(np.float64), b.astype(np.float64))
    assert np.abs(ss - ex).max() <= 1e-13 * np.abs(ex).max() * a[0].size ** 0.5
    assert abs(A.dotProduct(B) - ex.sum()) <= 1e-12 * np.abs(ex).sum()


def test_anderson_mix_on_orbitals(H):
    """AndersonMix driven with device orbitals (C-ABI BLAS-1) equals the same
    mixer on host numpy vectors (which tests/test_anderson_cpu.py pins to the
    reference's AndersonMix)."""
    from mgmol_b200.mixing import AndersonMix
    from test_anderson_cpu import Vec
    dims, N, m, beta = (8, 8, 8), 3, 3, 0.8
    grid = H.Grid(dims, (2.0, 2.0, 2.0), 1)
    rng = np.random.default_rng(9)
    x0 = rng.standard_normal((N,) + dims)
    amat = np.abs(rng.standard_normal((N,) + dims)) + 0.5      # diagonal operator
    X = H.Orbitals(grid, N, torch.float64, dev(x0))
    W = H.Orbitals(grid, N, torch.float64)
    mix = AndersonMix(m, beta, X, lambda o: o.clone())
    xv, wv = Vec(x0.ravel()), Vec(x0.ravel())
    Vec.invs = grid.vel()
    mixv = AndersonMix(m, beta, xv, lambda v: Vec(v.u))
    for it in range(8):
        xh = host(X.psi())
        assert np.abs(xh.ravel() - xv.u).max() <= 1e-11 * np.abs(xv.u).max(), it
        lam = float((xh * amat * xh).sum() / (xh * xh).sum())
        r = 0.3 * (lam * xh - amat * xh)
        F = H.Orbitals(grid, N, torch.float64, dev(r))
        mix.update(F, W)
        lamv = float((xv.u * amat.ravel() * xv.u).sum() / (xv.u @ xv.u))
        fv = Vec(0.3 * (lamv * xv.u - amat.ravel() * xv.u))
        mixv.update(fv, wv)
    assert mix.mm_ == mixv.mm_


# --------------------------------------------------------------------------
# full-size, size-independent properties (BASELINE configs: 128^3 and 256^3)
# --------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("lap_type", [0, 2])
@pytest.mark.parametrize("n", [128, 256])
def test_full_size_tma_equals_bit_exact_kernel(H, dt, lap_type, n):
    """At the benchmark's grid the oracle takes minutes, so the TMA kernel is
    compared with the generic fused kernel, which the small-size tests pin
    bit-for-bit to the reference."""
    from oracle.oracle import H2O512_CELL
    dims, N = (n, n, n), 3
    ll = (H2O512_CELL * n / 256,) * 3
    g = torch.Generator(device="cuda").manual_seed(1234)
    phi = (torch.rand((N,) + dims, generator=g, device="cuda", dtype=torch.float64) - 0.5)
    x = torch.arange(n, device="cuda", dtype=torch.float64) / n
    phi += torch.cos(2 * np.pi * x)[None, :, None, None] * torch.cos(4 * np.pi * x)[None, None, None, :]
    phi = phi.to(TDT[dt]).contiguous()
    v = (torch.rand(dims, generator=g, device="cuda", dtype=torch.float64) * 2.5 - 2.0)
    from mgmol_b200._lib import lib, check
    grid = H.Grid(dims, ll, H.ghosts_for(lap_type))
    lap = H.LapFactory.createLap(grid, lap_type)
    outs = {}
    for path in (1, 2):
        check(lib().mgb_hpsi_force_path(path))
        try:
            out = torch.empty_like(phi)
            lap.applyWithPot(phi, v, out)
            assert lib().mgb_hpsi_last_path() == path
            outs[path] = out
        finally:
            lib().mgb_hpsi_force_path(0)
    scale = outs[2].abs().amax(dim=(1, 2, 3), keepdim=True).double()
    err = float(((outs[1].double() - outs[2].double()).abs() / scale).max())
    assert err <= TOL[dt], err


def test_full_size_linearity_and_translation(H):
    """H(a psi1 + b psi2) = a H psi1 + b H psi2 and periodic translation
    covariance at 128^3, FP64."""
    n, ll = 128, (23.5, 23.5, 23.5)
    grid = H.Grid((n, n, n), ll, 1)
    lap = H.LapFactory.createLap(grid, 0)
    g = torch.Generator(device="cuda").manual_seed(7)
    psi = torch.rand((2, n, n, n), generator=g, device="cuda", dtype=torch.float64) - 0.5
    v = torch.rand((n, n, n), generator=g, device="cuda", dtype=torch.float64) - 0.7
    hp = torch.empty_like(psi)
    lap.applyWithPot(psi, v, hp)
    comb = (0.3 * psi[0] - 1.7 * psi[1])[None].contiguous()
    hc = torch.empty_like(comb)
    lap.applyWithPot(comb, v, hc)
    lin = 0.3 * hp[0] - 1.7 * hp[1]
    assert float((hc[0] - lin).abs().max() / lin.abs().max()) < 1e-12
    sh = (5, 9, 3)
    psi_s = torch.roll(psi, sh, dims=(1, 2, 3)).contiguous()
    v_s = torch.roll(v, sh, dims=(0, 1, 2)).contiguous()
    hs = torch.empty_like(psi_s)
    lap.applyWithPot(psi_s, v_s, hs)
    assert torch.equal(hs, torch.roll(hp, sh, dims=(1, 2, 3)))


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_config_si4x4nanowire_shape(H, dt):
    """BASELINE configs[2] (examples/Si4x4nanowire): 128 x 128 x 32 grid, 40
    orbitals, Mehrstellen, 2 multigrid levels.  H psi: TMA kernel against the
    bit-exact generic kernel; V-cycle: fused against the literal sequence."""
    from mgmol_b200._lib import lib, check
    dims, N, lap_type = (128, 128, 32), 40, 0
    ll = (40.0, 40.0, 10.26)
    g = torch.Generator(device="cuda").manual_seed(4)
    phi = (torch.rand((N,) + dims, generator=g, device="cuda", dtype=torch.float64) - 0.5)
    phi = phi.to(TDT[dt]).contiguous()
    v = torch.rand(dims, generator=g, device="cuda", dtype=torch.float64) * 2.0 - 1.5
    grid = H.Grid(dims, ll, 1)
    lap = H.LapFactory.createLap(grid, lap_type)
    outs = {}
    for path in (1, 2):
        check(lib().mgb_hpsi_force_path(path))
        try:
            out = torch.empty_like(phi)
            lap.applyWithPot(phi, v, out)
            assert lib().mgb_hpsi_last_path() == path
            outs[path] = out
        finally:
            lib().mgb_hpsi_force_path(0)
    scale = outs[2].abs().amax(dim=(1, 2, 3), keepdim=True).double()
    assert float(((outs[1].double() - outs[2].double()).abs() / scale).max()) <= TOL[dt]
    res = {}
    for mode in (1, 2):
        orb = H.Orbitals(grid, N, TDT[dt], outs[2].clone())
        pc = H.OrbitalsPreconditioning()
        pc.setup(orb, 2, lap_type)
        pc.set_mode(mode)
        pc.gamma_ = 0.25
        pc.precond_mg(orb)
        res[mode] = orb.psi().double()
        pc.close()
    sc = res[1].abs().amax(dim=(1, 2, 3), keepdim=True)
    assert float(((res[2] - res[1]).abs() / sc).max()) <= MG_TOL


# --------------------------------------------------------------------------
# mputils BLAS-1 and the small utility entry points
# --------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 255, 4097, 3 * 20 * 20 * 20])
def test_blas1_against_mputils(dt, n):
    """MPaxpy / MPscal / MPdot (src/linear_algebra/mputils.cc:53-244): products
    in double, one rounding to the storage type per element -> axpy and scal
    bit-exact; the dot product is a double sum in another order -> 1e-14
    relative to sum|x y|."""
    import ctypes
    from mgmol_b200._lib import lib, check
    from oracle.oracle import Ref, _dt as odt, _ptr
    L = lib()
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(dt)
    y = rng.standard_normal(n).astype(dt)
    code = 1 if dt == np.float64 else 0
    for alpha in (-0.37, 1.0, 0.0, 2.5):
        dx, dy = dev(x), dev(y)
        check(L.mgb_axpy(code, n, alpha, ctypes.c_void_p(dx.data_ptr()),
                         ctypes.c_void_p(dy.data_ptr()), None))
        want = y + (alpha * x.astype(np.float64)).astype(dt)
        assert bits_equal(host(dy), want)
        if Ref.available() and dt == np.float32:
            r = y.copy()
            Ref().lib.ref_mpaxpy(odt(r.dtype), n, ctypes.c_double(alpha), _ptr(x), _ptr(r))
            assert bits_equal(host(dy), r)
        ds = dev(x)
        check(L.mgb_scal(code, n, alpha, ctypes.c_void_p(ds.data_ptr()), None))
        if alpha == 0.0:
            wants = np.zeros_like(x)
        else:
            wants = (alpha * x.astype(np.float64)).astype(dt)
        assert bits_equal(host(ds), wants)
    dx, dy = dev(x), dev(y)
    res = torch.zeros(1, dtype=torch.float64, device="cuda")
    check(L.mgb_dot(code, n, ctypes.c_void_p(dx.data_ptr()), ctypes.c_void_p(dy.data_ptr()),
                    ctypes.c_void_p(res.data_ptr()), None))
    xd, yd = x.astype(np.float64), y.astype(np.float64)
    want = float(np.dot(xd, yd))
    assert abs(float(res.item()) - want) <= 1e-14 * float(np.abs(xd * yd).sum()) + 1e-300
    if Ref.available():
        r = Ref().lib.ref_mpdot(odt(x.dtype), n, _ptr(x), _ptr(y))
        assert abs(float(res.item()) - r) <= 1e-13 * float(np.abs(xd * yd).sum())
    assert L.mgb_axpy(7, n, 1.0, ctypes.c_void_p(dx.data_ptr()),
                      ctypes.c_void_p(dy.data_ptr()), None) != 0


def test_version_and_host_registration():
    """mgb_host_register pins a caller's buffer so mgb_hpsi_host streams from it
    without staging; results are the same bits either way."""
    import ctypes
    from mgmol_b200._lib import lib, check, MgbGrid
    L = lib()
    assert L.mgb_version() >= 100
    dims, N = (16, 16, 16), 3
    phi = synthetic_orbitals(N, dims, np.float64)
    v = synthetic_potential(dims)
    g = MgbGrid()
    for d in range(3):
        g.dim[d] = g.gdim[d] = dims[d]
        g.h[d] = 0.3
        g.bc[d] = 1
        g.nproc[d] = 1
        g.coord[d] = 0
    g.ghosts = 2
    n = int(np.prod(dims))
    outs = []
    for pinned in (False, True):
        src = phi.copy()
        out = np.full_like(src, np.nan)
        if pinned:
            check(L.mgb_host_register(src.ctypes.data_as(ctypes.c_void_p), src.nbytes))
            check(L.mgb_host_register(out.ctypes.data_as(ctypes.c_void_p), out.nbytes))
        try:
            check(L.mgb_hpsi_host(0, 1, ctypes.byref(g), src.ctypes.data_as(ctypes.c_void_p), n,
                                  v.ctypes.data_as(ctypes.c_void_p),
                                  out.ctypes.data_as(ctypes.c_void_p), n, N, 0))
        finally:
            if pinned:
                check(L.mgb_host_unregister(src.ctypes.data_as(ctypes.c_void_p)))
                check(L.mgb_host_unregister(out.ctypes.data_as(ctypes.c_void_p)))
        outs.append(out)
    assert bits_equal(outs[0], outs[1])
    assert np.isfinite(outs[0]).all()
    assert L.mgb_host_register(None, 16) != 0
    assert L.mgb_host_unregister(None) == 0


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("N,dims", [(7, (12, 8, 16)), (40, (16, 16, 16)), (150, (16, 24, 32))])
def test_orthonormalize_loewdin(H, port, dt, N, dims):
    """ExtendedGridOrbitals::orthonormalizeLoewdin (src/ExtendedGridOrbitals.cc:
    1304-1358; transform of src/GramMatrix.cc:267-285): afterwards the Gram
    matrix is the identity, Phi moved by the symmetric S^-1/2 of the oracle's
    Gram matrix, and the iterative index advanced."""
    a = synthetic_orbitals(N, dims, dt)
    grid = H.Grid(dims, (2.0, 2.5, 3.0), 1)
    A = H.Orbitals(grid, N, TDT[dt], dev(a))
    idx = A.getIterativeIndex()
    P = host(A.orthonormalizeLoewdin())
    assert A.getIterativeIndex() > idx
    S = grid.vel() * np.einsum("ixyz,jxyz->ij", a.astype(np.float64), a.astype(np.float64))
    lam, V = np.linalg.eigh(S)
    Pref = (V / np.sqrt(lam)) @ V.T
    cond = lam[-1] / lam[0]
    eps = 1e-13 if dt == np.float64 else 2e-6
    assert np.abs(P - P.T).max() <= 1e-12 * np.abs(P).max()
    assert np.abs(P - Pref).max() <= eps * cond * np.abs(Pref).max()
    want = np.einsum("lj,lxyz->jxyz", Pref, a.astype(np.float64))
    got = host(A.psi()).astype(np.float64)
    assert np.abs(got - want).max() <= eps * cond * np.abs(want).max() * max(1, N / 64)
    G = host(A.computeGram())
    tolg = 1e-11 if dt == np.float64 else 2e-5
    assert np.abs(G - np.eye(N)).max() <= tolg * max(1.0, np.sqrt(cond))
    if N <= 40:
        # the same statement from the oracle's Gram matrix and Phi M
        Sp = port.gemm_tn(a, a, grid.vel())
        lp, Vp = np.linalg.eigh(Sp)
        ref = port.gemm_nn(a, (Vp / np.sqrt(lp)) @ Vp.T)
        assert np.abs(got - ref).max() <= eps * cond * np.abs(want).max()


# --------------------------------------------------------------------------
# non-local Kleinman-Bylander projectors (row f3)
# --------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("lap_type", [0, 2])
@pytest.mark.parametrize("dims,N,nions,radius", [((12, 10, 16), 5, 7, 0.9),
                                                 ((32, 32, 32), 9, 20, 1.4)])
def test_kb_projectors_against_oracle(H, port, dt, lap_type, dims, N, nions, radius):
    """computeKBpsi and computeHnlPhiAndAdd2HPhi (src/KBPsiMatrixSparse.cc:136-212,
    src/computeHij.cc:294-375) against the oracle's restatement of the same loops: the
    projections to a sum-order tolerance, the scatter BIT-IDENTICAL for float (the
    reference's (T) roundings are reproduced; for double its DAXPY is an un-vendored BLAS)."""
    from oracle.oracle import synthetic_kb_projectors
    ll = tuple(0.25 * n for n in dims)
    ions = synthetic_kb_projectors(dims, ll, nions, radius, dt)
    grid = H.Grid(dims, ll, H.ghosts_for(lap_type))
    kbp = H.KBProjectors(grid, TDT[dt])
    rows = [kbp.add_ion(i["nlindex"], i["proj"], i["coeff"]) for i in ions]
    kbp.commit()
    assert rows[0] == 0 and kbp.nrows() == sum(len(i["coeff"]) for i in ions)
    phi = synthetic_orbitals(N, dims, dt)
    lap = H.LapFactory.createLap(grid, lap_type)
    mehr = lap_type == 0
    # <beta | phi> (or <beta | B phi> with the Mehrstellen flag)
    kb = host(kbp.computeKBpsi(dev(phi), lapOper=lap if mehr else None))
    src = port.lap_rhs(lap_type, phi, ll) if mehr else phi
    kb_ref = port.kb_psi(ions, src, grid.vel())
    scale = np.abs(kb_ref).max()
    assert np.abs(kb - kb_ref).max() <= (1e-13 if dt == np.float64 else 1e-6) * scale
    # H phi += V_nl phi, fed with the ORACLE's projections so that the scatter is compared alone
    h0 = synthetic_orbitals(N, dims, dt, first=40)
    dh = dev(h0)
    kbp.computeHnlPhiAndAdd2HPhi(dev(kb_ref), dh, lapOper=lap)
    if mehr:
        v = port.kb_vnlpsi(ions, kb_ref, np.zeros_like(h0), add=False)
        exp = h0 + port.lap_rhs(0, v, ll)
    else:
        exp = port.kb_vnlpsi(ions, kb_ref, h0, add=True)
    if dt == np.float32:
        assert bits_equal(host(dh), exp)
    else:
        assert rel_inf(host(dh), exp) <= 1e-15
    kbp.close()
