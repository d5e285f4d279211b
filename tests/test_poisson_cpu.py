"""Poisson multigrid without a GPU: the product's control
flow (mgmol_b200/poisson.py: SolverLap::solve -> Mgm -> Vcycle) driven by a numpy
field whose operations are the oracle's, against the compiled reference solver
and against the committed golden vectors."""
import os

import numpy as np
import pytest

from oracle.oracle import Ref
from poisson_cases import CASES, DEFAULTS, DTYPES, LAPS, PCG_CASES, charge, guess, key
from poisson_twin import field_factory

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gpois():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_poisson.npz")))


def _solve_twin(port, tag, dims, ll, bc, kw, lt, dt):
    from mgmol_b200.host import Grid
    from mgmol_b200.poisson import PoissonMG
    par = dict(DEFAULTS, **kw)
    solver = PoissonMG(Grid(dims, ll, 1, bc), lt, dt, field=field_factory(port))
    solver.setup(par["nu1"], par["nu2"], par["max_sweeps"], par["tol"], par["max_nlevels"])
    vh = guess(dims, dt)
    conv = solver.solve(vh, charge(dims, bc, dt))
    return vh, conv, solver


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("lt", LAPS)
@pytest.mark.parametrize("dt", DTYPES, ids=["f64", "f32"])
def test_control_flow_against_golden(port, gpois, case, lt, dt):
    tag, dims, ll, bc, kw = case
    vh, conv, solver = _solve_twin(port, tag, dims, ll, bc, kw, lt, dt)
    ref = gpois[key(tag, lt, dt)]
    st = gpois[key(tag, lt, dt) + "_stats"]
    # same operations in the same order: differences only from the order of
    # the sums inside norms and the average (1 ulp)
    eps = 1e-14 if dt == np.float64 else 1e-6
    assert np.abs(vh.astype(np.float64) - ref).max() <= eps * np.abs(ref).max()
    assert conv == bool(st[0])
    assert solver.getNbSweeps() == int(st[1])
    assert solver.getFinalResidual() == pytest.approx(st[2], rel=1e-6)
    assert solver.getFinalRelativeResidual() == pytest.approx(st[3], rel=1e-6)
    assert solver.getResidualReduction() == pytest.approx(st[4], rel=1e-6)


@pytest.mark.skipif(not Ref.available(), reason="compiled reference not present")
def test_golden_is_what_the_compiled_reference_returns(gpois):
    ref = Ref()
    for tag, dims, ll, bc, kw in CASES[:2]:
        for lt in LAPS:
            for dt in DTYPES:
                vh, conv, st = ref.poisson_solve(lt, guess(dims, dt), charge(dims, bc, dt), ll, bc,
                                                 **dict(DEFAULTS, **kw))
                assert np.array_equal(vh, gpois[key(tag, lt, dt)])
                assert conv == bool(gpois[key(tag, lt, dt) + "_stats"][0])


@pytest.mark.parametrize("lt", LAPS)
@pytest.mark.parametrize("dt", DTYPES, ids=["f64", "f32"])
def test_c_restatement_against_golden(port, gpois, lt, dt):
    """oracle/mgmol_oracle_poisson.inc (orc_poisson_mg, orc_poisson_pcg) against
    the compiled reference's outputs: the multigrid solver to the bit, the
    conjugate gradient within the reordering of its dot products."""
    for tag, dims, ll, bc, kw in CASES:
        vh, conv, st = port.poisson_solve(lt, guess(dims, dt), charge(dims, bc, dt), ll, bc,
                                          **dict(DEFAULTS, **kw))
        ref, rst = gpois[key(tag, lt, dt)], gpois[key(tag, lt, dt) + "_stats"]
        assert np.array_equal(vh, ref), (tag, lt, dt)
        assert conv == bool(rst[0]) and st[0] == rst[1]
        assert st[1] == pytest.approx(rst[2], rel=1e-12)
    for tag, dims, ll, bc, kw in PCG_CASES:
        vh, conv, st = port.pcg_solve(lt, guess(dims, dt), charge(dims, bc, dt), ll, bc,
                                      **dict(DEFAULTS, **kw))
        ref, rst = gpois[key(tag, lt, dt)], gpois[key(tag, lt, dt) + "_stats"]
        eps = 1e-11 if dt == np.float64 else 1e-6
        assert np.abs(vh.astype(np.float64) - ref).max() <= eps * np.abs(ref).max(), (tag, lt, dt)
        assert conv == bool(rst[0])
        assert st[0] == pytest.approx(rst[1], rel=1e-6)


def test_converges_to_the_discrete_solution(port):
    """A V(2,2) cycle contracts the residual by about an order of magnitude per
    sweep; the converged flag, the sweep count and the analytic solution of a
    single Fourier mode."""
    from mgmol_b200.host import Grid
    from mgmol_b200.poisson import PoissonMG
    dims, ll = (16, 16, 16), (4.0, 4.0, 4.0)
    x = np.arange(16) * 0.25
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    k = 2 * np.pi / 4
    rho = np.sin(k * X) * np.cos(k * Y) + 0.3 * np.sin(2 * k * Z)
    exact = np.sin(k * X) * np.cos(k * Y) / (2 * k * k) + 0.3 * np.sin(2 * k * Z) / (4 * k * k)
    for lt, err in ((0, 1e-4), (2, 3e-4), (1, 6e-3)):
        solver = PoissonMG(Grid(dims, ll, 1, (1, 1, 1)), lt, np.float64, field=field_factory(port))
        solver.setup(2, 2, 40, 1e-10, 10)
        vh = np.zeros(dims)
        assert solver.solve(vh, rho)
        assert solver.getNbSweeps() < 40
        assert solver.getFinalRelativeResidual() < 1e-10
        assert np.abs(vh - exact).max() < err
        assert abs(vh.mean()) < 1e-14


@pytest.mark.parametrize("case", PCG_CASES, ids=[c[0] for c in PCG_CASES])
@pytest.mark.parametrize("lt", LAPS)
@pytest.mark.parametrize("dt", DTYPES, ids=["f64", "f32"])
def test_pcg_control_flow_against_golden(port, gpois, case, lt, dt):
    """PoissonPCG (PCGSolver::solve + preconSolve in float) on oracle fields
    against the compiled reference."""
    from mgmol_b200.host import Grid
    from mgmol_b200.poisson import PoissonPCG
    tag, dims, ll, bc, kw = case
    par = dict(DEFAULTS, **kw)
    solver = PoissonPCG(Grid(dims, ll, 1, bc), lt, dt, field=field_factory(port),
                        precond_dtype=np.float32)
    solver.setup(par["nu1"], par["nu2"], par["max_sweeps"], par["tol"], par["max_nlevels"])
    vh = guess(dims, dt)
    conv = solver.solve(vh, charge(dims, bc, dt))
    ref = gpois[key(tag, lt, dt)]
    st = gpois[key(tag, lt, dt) + "_stats"]
    # the CG scalars are ratios of dot products summed in another order
    eps = 1e-10 if dt == np.float64 else 2e-6
    assert np.abs(vh.astype(np.float64) - ref).max() <= eps * np.abs(ref).max()
    assert conv == bool(st[0])
    assert solver.getFinalResidual() == pytest.approx(st[1], rel=1e-5)
    assert solver.getResidualReduction() == pytest.approx(st[2], rel=1e-5)


def test_pcg_zero_right_hand_side_returns_at_once(port):
    from mgmol_b200.host import Grid
    from mgmol_b200.poisson import PoissonPCG
    dims = (8, 8, 8)
    solver = PoissonPCG(Grid(dims, (2.0, 2.0, 2.0), 1), 0, np.float64, field=field_factory(port),
                        precond_dtype=np.float32)
    vh = np.zeros(dims)
    assert solver.solve(vh, np.zeros(dims)) and not vh.any()


def _gdot(a, b, bc):
    sl = tuple(slice(1 if c != 1 else 0, None) for c in bc)
    return float((a[sl].astype(np.float64) * b[sl].astype(np.float64)).sum())


@pytest.mark.skipif(not Ref.available(), reason="compiled reference not present")
@pytest.mark.parametrize("pcg", [False, True], ids=["mg", "pcg"])
@pytest.mark.parametrize("dt", DTYPES, ids=["f64", "f32"])
def test_hartree_sequence(port, pcg, dt):
    """Hartree<T>::solve (src/Hartree.cc:27-112): rhs = 4 pi (rho - rhoc) in the
    solver's precision, the solver started from the previous vh, and the three
    integrals -- against the same steps done with the compiled reference
    solver, over two consecutive calls."""
    import math
    from mgmol_b200.host import Grid
    from mgmol_b200.poisson import Hartree
    ref = Ref()
    for (tag, dims, ll, bc, _), lt in ((CASES[0], 0), (CASES[1], 2), (CASES[2], 1)):
        rho = charge(dims, bc, np.float64)
        rhoc = 0.3 * charge(dims, bc, np.float64, seed=11)
        har = Hartree(Grid(dims, ll, 1, bc), lt, dt, field=field_factory(port), pcg=pcg,
                      rho_dtype=np.float64, precond_dtype=np.float32)
        har.setup(2, 2, 5, 1e-16, 10)
        vel = float(np.prod([a / n for a, n in zip(ll, dims)]))
        vref = np.zeros(dims, dt)
        for call in range(2):
            old = vel * _gdot(vref, rho.astype(dt), bc)
            har.solve(rho, rhoc)
            rhs = ((rho - rhoc).astype(dt).astype(np.float64) * (4 * math.pi)).astype(dt)
            solve = ref.pcg_solve if pcg else ref.poisson_solve
            vref, _, st = solve(lt, vref, rhs, ll, bc, max_sweeps=5)
            eps = (1e-11 if pcg else 1e-14) if dt == np.float64 else 1e-6
            got = har.vh().reshape(dims).astype(np.float64)
            assert np.abs(got - vref).max() <= eps * np.abs(vref).max(), (tag, call)
            assert har.IntVhRho_old() == pytest.approx(old, rel=1e-9, abs=1e-300)
            assert har.IntVhRho() == pytest.approx(vel * _gdot(vref, rho.astype(dt), bc), rel=1e-9)
            assert har.IntVhRhoc() == pytest.approx(vel * _gdot(vref, rhoc.astype(dt), bc), rel=1e-9)
            assert har.getFinalResidual() == pytest.approx(st[-3] if not pcg else st[0], rel=1e-4)
        har.resetVh()
        assert not har.vh().any()


def test_refuses_what_is_not_built():
    from mgmol_b200.host import Grid
    from mgmol_b200.poisson import PoissonMG
    with pytest.raises(ValueError):
        PoissonMG(Grid((8, 8, 8), (1.0, 1.0, 1.0), 1), 3)
    with pytest.raises(ValueError):
        PoissonMG(Grid((8, 8, 8), (1.0, 1.0, 1.0), 1, (2, 2, 2)), 0)
    with pytest.raises(ValueError):
        PoissonMG(Grid((8, 8, 8), (1.0, 1.0, 1.0), 1, (1, 1, 1), (2, 1, 1), (0, 0, 0)), 0)
    from mgmol_b200.poisson import PoissonPCG
    with pytest.raises(ValueError):
        PoissonPCG(Grid((8, 8, 8), (1.0, 1.0, 1.0), 1), 4)
