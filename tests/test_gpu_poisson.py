"""Poisson multigrid on the device: PoissonMG over the
C-ABI operations, against the golden vectors of the compiled reference solver
(pb::SolverLap / Mgm / Vcycle) and against the same control flow run on the
oracle's operations."""
import os

import numpy as np
import pytest
import torch

from poisson_cases import CASES, DEFAULTS, DTYPES, LAPS, PCG_CASES, charge, guess, key
from poisson_twin import field_factory

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TDT = {np.float64: torch.float64, np.float32: torch.float32}


@pytest.fixture(scope="module")
def gpois():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_poisson.npz")))


def _solve_device(dims, ll, bc, kw, lt, dt):
    from mgmol_b200.host import Grid
    from mgmol_b200.poisson import PoissonMG
    par = dict(DEFAULTS, **kw)
    solver = PoissonMG(Grid(dims, ll, 1, bc), lt, TDT[dt])
    solver.setup(par["nu1"], par["nu2"], par["max_sweeps"], par["tol"], par["max_nlevels"])
    vh = torch.from_numpy(guess(dims, dt)).cuda()
    conv = solver.solve(vh, torch.from_numpy(charge(dims, bc, dt)).cuda())
    torch.cuda.synchronize()
    return vh.cpu().numpy(), conv, solver


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("lt", LAPS)
@pytest.mark.parametrize("dt", DTYPES, ids=["f64", "f32"])
def test_poisson_against_golden(gpois, case, lt, dt):
    tag, dims, ll, bc, kw = case
    from mgmol_b200._lib import lib
    n0 = lib().mgb_launch_count()
    vh, conv, solver = _solve_device(dims, ll, bc, kw, lt, dt)
    assert lib().mgb_launch_count() > n0          # the library's kernels did the work
    ref = gpois[key(tag, lt, dt)]
    st = gpois[key(tag, lt, dt) + "_stats"]
    # every grid operation is bit-exact against the reference kernel; only the
    # order of the sums inside norms and the average differs
    eps = 1e-13 if dt == np.float64 else 2e-6
    assert np.abs(vh.astype(np.float64) - ref).max() <= eps * np.abs(ref).max()
    assert conv == bool(st[0])
    assert solver.getNbSweeps() == int(st[1])
    assert solver.getFinalResidual() == pytest.approx(st[2], rel=1e-6)
    assert solver.getResidualReduction() == pytest.approx(st[4], rel=1e-6)


@pytest.mark.parametrize("lt", LAPS)
def test_poisson_device_equals_oracle_control_flow(port, lt):
    """The same PoissonMG code on device fields and on oracle fields."""
    from mgmol_b200.host import Grid
    from mgmol_b200.poisson import PoissonMG
    dims, ll, bc = (20, 12, 28), (3.0, 2.0, 4.5), (1, 0, 1)
    for dt in DTYPES:
        got, conv, solver = _solve_device(dims, ll, bc, {}, lt, dt)
        twin = PoissonMG(Grid(dims, ll, 1, bc), lt, dt, field=field_factory(port))
        vh = guess(dims, dt)
        conv2 = twin.solve(vh, charge(dims, bc, dt))
        eps = 1e-13 if dt == np.float64 else 2e-6
        assert conv == conv2 and solver.getNbSweeps() == twin.getNbSweeps()
        assert np.abs(got.astype(np.float64) - vh).max() <= eps * np.abs(vh).max()


def test_poisson_larger_box_residual():
    """64^3 periodic box (coarsens to 1^3): ten V(2,2) sweeps take the relative
    residual below 1e-8 and the solution has zero mean."""
    dims, ll, bc = (64, 64, 64), (8.0, 8.0, 8.0), (1, 1, 1)
    vh, conv, solver = _solve_device(dims, ll, bc, {}, 0, np.float64)
    assert solver.getNbSweeps() == 10 and not conv
    assert solver.getFinalRelativeResidual() < 1e-8
    assert abs(vh.mean()) < 1e-13 * np.abs(vh).max()


def _pcg_device(dims, ll, bc, kw, lt, dt):
    from mgmol_b200.host import Grid
    from mgmol_b200.poisson import PoissonPCG
    par = dict(DEFAULTS, **kw)
    solver = PoissonPCG(Grid(dims, ll, 1, bc), lt, TDT[dt])
    solver.setup(par["nu1"], par["nu2"], par["max_sweeps"], par["tol"], par["max_nlevels"])
    vh = torch.from_numpy(guess(dims, dt)).cuda()
    conv = solver.solve(vh, torch.from_numpy(charge(dims, bc, dt)).cuda())
    torch.cuda.synchronize()
    return vh.cpu().numpy(), conv, solver


@pytest.mark.parametrize("case", PCG_CASES, ids=[c[0] for c in PCG_CASES])
@pytest.mark.parametrize("lt", LAPS)
@pytest.mark.parametrize("dt", DTYPES, ids=["f64", "f32"])
def test_pcg_against_golden(gpois, case, lt, dt):
    """PCGSolver::solve (src/PCGSolver.cc:165-252) with the float multigrid
    preconditioner, on the device, against the compiled reference.  The CG
    scalars are ratios of dot products the device sums in another order."""
    tag, dims, ll, bc, kw = case
    vh, conv, solver = _pcg_device(dims, ll, bc, kw, lt, dt)
    ref = gpois[key(tag, lt, dt)]
    st = gpois[key(tag, lt, dt) + "_stats"]
    eps = 1e-9 if dt == np.float64 else 5e-6
    assert np.abs(vh.astype(np.float64) - ref).max() <= eps * np.abs(ref).max()
    assert conv == bool(st[0])
    assert solver.getFinalResidual() == pytest.approx(st[1], rel=1e-4)
    assert solver.getResidualReduction() == pytest.approx(st[2], rel=1e-4)


@pytest.mark.parametrize("solver", [0, 1], ids=["mg", "pcg"])
@pytest.mark.parametrize("lt", [0, 2])
@pytest.mark.parametrize("dt", DTYPES, ids=["f64", "f32"])
def test_poisson_c_abi_entry(gpois, solver, lt, dt):
    """mgb_poisson_solve: the C entry over the same solver code (the C++
    templates instantiated inside the library), against the golden vectors."""
    import ctypes
    from mgmol_b200._lib import lib, check
    from mgmol_b200.host import Grid
    cases = (CASES[0], CASES[2]) if solver == 0 else (PCG_CASES[0], PCG_CASES[2])
    for tag, dims, ll, bc, kw in cases:
        par = dict(DEFAULTS, **kw)
        grid = Grid(dims, ll, 1, bc)
        vh = torch.from_numpy(guess(dims, dt)).cuda()
        rho = torch.from_numpy(charge(dims, bc, dt)).cuda()
        stats = (ctypes.c_double * 5)()
        check(lib().mgb_poisson_solve(
            solver, lt, 1 if dt == np.float64 else 0, grid.ref(), ctypes.c_void_p(vh.data_ptr()),
            ctypes.c_void_p(rho.data_ptr()), par["nu1"], par["nu2"], par["max_sweeps"], par["tol"],
            par["max_nlevels"], stats))
        torch.cuda.synchronize()
        ref = gpois[key(tag, lt, dt)]
        st = gpois[key(tag, lt, dt) + "_stats"]
        if solver == 0:
            eps = 1e-13 if dt == np.float64 else 2e-6
            assert int(stats[1]) == int(st[1])
            assert stats[2] == pytest.approx(st[2], rel=1e-6)
        else:
            eps = 1e-9 if dt == np.float64 else 5e-6
            assert stats[2] == pytest.approx(st[1], rel=1e-4)
        assert np.abs(vh.cpu().numpy().astype(np.float64) - ref).max() <= eps * np.abs(ref).max()
        assert bool(stats[0]) == bool(st[0])
    # refusals come back as error codes, not aborts
    g3 = Grid((8, 8, 8), (1.0, 1.0, 1.0), 1)
    z = torch.zeros((8, 8, 8), dtype=torch.float64, device="cuda")
    p = ctypes.c_void_p(z.data_ptr())
    assert lib().mgb_poisson_solve(0, 3, 1, g3.ref(), p, p, 2, 2, 10, 1e-16, 10, None) != 0
    assert lib().mgb_poisson_solve(2, 0, 1, g3.ref(), p, p, 2, 2, 10, 1e-16, 10, None) != 0


@pytest.mark.parametrize("pcg,dt", [(False, np.float64), (False, np.float32), (True, np.float64)],
                         ids=["mg-f64", "mg-f32", "pcg-f64"])
def test_hartree_device_equals_oracle_control_flow(port, pcg, dt):
    """Hartree<T>::solve (src/Hartree.cc:27-112) on device fields and on oracle
    fields: two consecutive calls (the second starts from the first vh), the
    potential and the three integrals."""
    from mgmol_b200.host import Grid
    from mgmol_b200.poisson import Hartree
    tag, dims, ll, bc, _ = CASES[2]
    rho = charge(dims, bc, np.float64)
    rhoc = 0.3 * charge(dims, bc, np.float64, seed=11)
    dev_h = Hartree(Grid(dims, ll, 1, bc), 0, TDT[dt], pcg=pcg)
    cpu_h = Hartree(Grid(dims, ll, 1, bc), 0, dt, field=field_factory(port), pcg=pcg,
                    rho_dtype=np.float64, precond_dtype=np.float32)
    drho, drhoc = torch.from_numpy(rho).cuda(), torch.from_numpy(rhoc).cuda()
    eps = (1e-10 if pcg else 1e-13) if dt == np.float64 else 2e-6
    for _ in range(2):
        dev_h.setup(2, 2, 4, 1e-16, 10)
        cpu_h.setup(2, 2, 4, 1e-16, 10)
        dev_h.solve(drho, drhoc)
        cpu_h.solve(rho, rhoc)
        got = dev_h.vh().cpu().numpy().reshape(dims).astype(np.float64)
        want = cpu_h.vh().reshape(dims).astype(np.float64)
        assert np.abs(got - want).max() <= eps * np.abs(want).max()
        for a, b in ((dev_h.IntVhRho(), cpu_h.IntVhRho()), (dev_h.IntVhRhoc(), cpu_h.IntVhRhoc()),
                     (dev_h.IntVhRho_old(), cpu_h.IntVhRho_old())):
            assert a == pytest.approx(b, rel=1e-6, abs=1e-12)
