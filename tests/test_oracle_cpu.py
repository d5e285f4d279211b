"""CPU tests (no GPU): the C restatement against the golden vectors produced
by the reference's own compiled sources, and -- where oracle/_ref exists --
bit-for-bit against that compiled reference itself."""
import numpy as np
import pytest

from conftest import bits_equal
from oracle.oracle import Ref, synthetic_orbitals, synthetic_potential

BCS = [(1, 1, 1), (0, 0, 0), (1, 0, 1)]


def _inputs(golden, dt):
    dims = tuple(int(x) for x in golden["dims"])
    ll = tuple(float(x) for x in golden["ll"])
    N = int(golden["nfunc"])
    return dims, ll, N, synthetic_orbitals(N, dims, dt), synthetic_potential(dims)


@pytest.mark.parametrize("dt,tag", [(np.float64, "f64"), (np.float32, "f32")])
def test_hpsi_matches_golden(port, golden, dt, tag):
    dims, ll, N, phi, v = _inputs(golden, dt)
    for lt in (0, 2):
        for bc in BCS:
            ref = golden["hpsi_lap%d_%s_bc%d%d%d" % ((lt, tag) + bc)]
            assert bits_equal(port.hpsi(lt, phi, v, ll, bc), ref), (lt, bc)


@pytest.mark.parametrize("dt,tag", [(np.float64, "f64"), (np.float32, "f32")])
def test_trade_boundaries_matches_golden(port, golden, dt, tag):
    dims, ll, N, phi, v = _inputs(golden, dt)
    for g in (1, 2):
        for bc in ((1, 1, 1), (0, 0, 0), (0, 1, 1)):
            ref = golden["trade_g%d_%s_bc%d%d%d" % ((g, tag) + bc)]
            assert bits_equal(port.trade_boundaries(phi, g, bc), ref), (g, bc)


@pytest.mark.parametrize("dt,tag", [(np.float64, "f64"), (np.float32, "f32")])
def test_fd_kernels_match_golden(port, golden, dt, tag):
    dims, ll, N, phi, v = _inputs(golden, dt)
    h = tuple(l / d for l, d in zip(ll, dims))
    for kind, g in ((0, 1), (1, 1), (2, 2), (100, 1)):
        gv = port.trade_boundaries(phi, g)
        assert bits_equal(port.fdkernel(kind, gv, g, h), golden["fd%d_%s" % (kind, tag)])


@pytest.mark.parametrize("dt,tag", [(np.float64, "f64"), (np.float32, "f32")])
def test_mg_transfer_matches_golden(port, golden, dt, tag):
    dims, ll, N, phi, v = _inputs(golden, dt)
    gv = port.trade_boundaries(phi, 1)
    c = port.restrict3D(gv, 1)
    assert bits_equal(c, golden["restrict_%s" % tag])
    assert bits_equal(port.extend3D(c, gv, 1), golden["extend_%s" % tag])


@pytest.mark.parametrize("dt,tag", [(np.float64, "f64"), (np.float32, "f32")])
def test_precond_matches_golden(port, golden, dt, tag):
    dims, ll, N, phi, v = _inputs(golden, dt)
    for lt in (0, 2):
        for bc in ((1, 1, 1), (0, 0, 0)):
            for lev in (1, 2):
                ref = golden["precond_lap%d_lev%d_%s_bc%d%d%d" % ((lt, lev, tag) + bc)]
                assert bits_equal(port.precond_mg(lt, lev, phi, ll, 0.37, bc), ref)


def test_lap_constants_match_golden(port, golden):
    dims = tuple(int(x) for x in golden["dims"])
    ll = tuple(float(x) for x in golden["ll"])
    for lt in (0, 1, 2):
        g = 2 if lt == 2 else 1
        assert np.array_equal(np.array(port.lap_constants(lt, dims, g, ll)),
                              golden["lapconst_%d" % lt])
    for lt, g in ((3, 3), (4, 4)):
        assert np.array_equal(np.array(port.lap_constants(lt, (8, 8, 8), g, (2., 2., 2.))),
                              golden["lapconst_%d" % lt])


@pytest.mark.parametrize("dt,tag", [(np.float64, "f64"), (np.float32, "f32")])
def test_contractions_match_golden(port, golden, dt, tag):
    a = synthetic_orbitals(5, (6, 4, 8), dt)
    b = synthetic_orbitals(5, (6, 4, 8), dt, first=11)
    M = np.random.default_rng(7).standard_normal((5, 5))
    got = port.gemm_tn(a, b, 0.37)
    ref = golden["gemm_tn_%s" % tag]
    if dt == np.float32:
        # float operands: the reference's own double-accumulating loops
        assert bits_equal(got, ref)
    else:
        # double operands: the reference calls an un-vendored DGEMM whose
        # summation order is unspecified ("parity unpinned" at the last bit)
        assert np.abs(got - ref).max() <= 64 * np.finfo(np.float64).eps * np.abs(ref).max()
    sy = golden["syrk_%s" % tag]
    assert np.abs(port.gemm_tn(a, a, 0.37) - sy).max() <= 1e-13 * np.abs(sy).max()
    assert bits_equal(port.gemm_nn(a, M), golden["gemm_nn_%s" % tag])


def test_reference_unit_tests_properties(port):
    """The reference's own unit tests for this path, restated:
    tests/testLaph4.cc:10-86 / testBatchLaph4.cc:7-110 (4th-order Laplacian of
    sin x + sin y + sin z vs k^2 sin, margin 2e-3) and tests/testMGkernels.cc:
    8-85 (restrict then extend of constant fields returns the constants,
    margin 1e-8)."""
    dims, g = (32, 24, 20), 2
    ll = (2 * np.pi,) * 3
    h = tuple(l / d for l, d in zip(ll, dims))
    x = np.arange(dims[0])[:, None, None] * h[0]
    y = np.arange(dims[1])[None, :, None] * h[1]
    z = np.arange(dims[2])[None, None, :] * h[2]
    u = (np.sin(x) + np.sin(y) + np.sin(z))[None].repeat(10, 0)
    lap = port.strip_ghosts(port.fdkernel(2, port.trade_boundaries(u, g), g, h), g)
    assert np.abs(lap - u).max() < 2e-3
    nf, g = 5, 2
    const = np.stack([np.full((16, 16, 16), 1.0 + i) for i in range(nf)])
    gv = port.trade_boundaries(const, g)
    c = port.restrict3D(gv, g)
    f = port.extend3D(c, np.zeros_like(gv), g)
    inner = port.strip_ghosts(f, g)
    for i in range(nf):
        assert np.abs(inner[i] - (1.0 + i)).max() < 1e-8


def test_ghost_values_reference_test(port):
    """tests/testTradeGhostValues.cc:27-198: every ghost value, edges and
    corners included, equals the periodic analytic function (eps 1e-8)."""
    for g in (1, 2):
        for dims in ((32, 24, 20), (24, 20, 32), (20, 32, 24)):
            ll = (1.0, 1.0, 1.0)
            nx, ny, nz = dims

            def fn(i, j, k):
                return (np.cos(2 * np.pi * i / nx) * np.cos(2 * np.pi * j / ny)
                        * np.cos(2 * np.pi * k / nz) + 1.111)
            i = np.arange(nx)[:, None, None]
            j = np.arange(ny)[None, :, None]
            k = np.arange(nz)[None, None, :]
            u = np.stack([fn(i, j, k) * (1 + f) for f in range(3)])
            gu = port.trade_boundaries(u, g)
            ig = np.arange(-g, nx + g)[:, None, None]
            jg = np.arange(-g, ny + g)[None, :, None]
            kg = np.arange(-g, nz + g)[None, None, :]
            for f in range(3):
                assert np.abs(gu[f] - fn(ig, jg, kg) * (1 + f)).max() < 1e-8


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built here")
def test_port_is_bit_exact_against_compiled_reference(port):
    """Pins the restatement against the reference's own compiled sources on
    fresh inputs (different from the golden ones)."""
    R = Ref()
    dims, ll, N = (16, 12, 20), (6.0, 5.0, 7.0), 3
    for dt in (np.float64, np.float32):
        phi = synthetic_orbitals(N, dims, dt, first=5)
        v = synthetic_potential(dims)
        for lt in (0, 2):
            for bc in BCS:
                assert bits_equal(port.hpsi(lt, phi, v, ll, bc), R.hpsi(lt, phi, v, ll, bc))
        for lt in (0, 2):
            for bc in ((1, 1, 1), (0, 0, 0)):
                big = synthetic_orbitals(2, (16, 8, 24), dt)
                assert bits_equal(port.precond_mg(lt, 2, big, (4., 2., 6.), 0.2, bc),
                                  R.precond_mg(lt, 2, big, (4., 2., 6.), 0.2, bc))


@pytest.mark.skipif(not Ref.available(), reason="compiled reference not present")
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("lap_type", [0, 10, 2])
@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 0, 0), (1, 0, 1)])
def test_lap_rhs_port_is_bit_exact_against_compiled_reference(port, dt, lap_type, bc):
    """Lap::rhs per orbital: B of Laph4M, B2 of Laph4MP, identity otherwise."""
    if lap_type == 2 and bc != (1, 1, 1):
        # B = 1: FDoper::rhs returns the function as it stands; after a Dirichlet
        # trade that includes the zeroed low layers -- the residual path never
        # calls rhs for these operators (applyB is false, src/MGmol.cc:1246)
        pytest.skip("identity right-hand side is only exercised periodic")
    R = Ref()
    dims, ll = (10, 8, 12), (3.0, 2.2, 4.1)
    phi = synthetic_orbitals(3, dims, dt)
    assert bits_equal(port.lap_rhs(lap_type, phi, ll, bc), R.lap_rhs(lap_type, phi, ll, bc))


@pytest.fixture(scope="module")
def gf1():
    import os
    from conftest import ROOT
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_f1.npz")))


@pytest.mark.parametrize("dt,tag", [(np.float64, "f64"), (np.float32, "f32")])
def test_f1_rows_match_golden(port, gf1, dt, tag):
    """Lap::rhs (B, B2), the residual sequence and the density sequence of the
    restatement against vectors made by the compiled reference."""
    dims = tuple(int(x) for x in gf1["dims"])
    ll = tuple(float(x) for x in gf1["ll"])
    N = int(gf1["nfunc"])
    theta = gf1["theta"]
    phi = synthetic_orbitals(N, dims, dt)
    v = synthetic_potential(dims)
    for lt in (0, 10):
        for bc in ((1, 1, 1), (0, 0, 0)):
            assert bits_equal(port.lap_rhs(lt, phi, ll, bc),
                              gf1["rhs_lap%d_%s_bc%d%d%d" % ((lt, tag) + bc)])
    tol = 1e-13 if dt == np.float64 else 2e-7
    for lt in (0, 2):
        hphi = port.hpsi(lt, phi, v, ll)
        bphi = port.lap_rhs(lt, phi, ll) if lt == 0 else phi
        res = (port.gemm_nn(bphi, theta) - hphi).astype(dt)
        ref = gf1["residual_lap%d_%s" % (lt, tag)]
        # the FP64 product goes through the reference's (unpinned) DGEMM
        assert np.abs(res.astype(np.float64) - ref).max() <= tol * np.abs(ref).max()
    product = port.gemm_nn(phi, theta)
    rho = np.zeros(dims)
    for j in range(N):
        rho += (product[j] * phi[j]).astype(np.float64)
    assert np.abs(rho - gf1["rho_%s" % tag]).max() <= tol * np.abs(gf1["rho_%s" % tag]).max()


@pytest.mark.skipif(not Ref.available(), reason="compiled reference not present")
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_blas1_statement_against_compiled_reference(dt):
    """The one-line statements the GPU BLAS-1 tests use (double product, one
    rounding to the storage type) are what MPaxpy / MPdot compute
    (src/linear_algebra/mputils.cc:130-244)."""
    import ctypes
    from oracle.oracle import _dt, _ptr
    R = Ref()
    rng = np.random.default_rng(3)
    n = 4097
    x = rng.standard_normal(n).astype(dt)
    y = rng.standard_normal(n).astype(dt)
    for alpha in (-0.37, 1.0, 0.0, 2.5):
        r = y.copy()
        R.lib.ref_mpaxpy(_dt(dt), n, ctypes.c_double(alpha), _ptr(x), _ptr(r))
        want = y + (alpha * x.astype(np.float64)).astype(dt)
        if dt == np.float32:
            assert bits_equal(r, want)
        else:   # DAXPY of the BLAS at hand may fuse the multiply-add
            assert np.allclose(r, want, rtol=0, atol=4e-16 * np.abs(want).max())
    d = R.lib.ref_mpdot(_dt(dt), n, _ptr(x), _ptr(y))
    xd, yd = x.astype(np.float64), y.astype(np.float64)
    assert abs(d - float(np.dot(xd, yd))) <= 1e-13 * float(np.abs(xd * yd).sum())
