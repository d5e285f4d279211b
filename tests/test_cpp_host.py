"""The C++ host side (include/mgmol_b200.hpp): compiles with a plain C++11
compiler against the C ABI (no CUDA headers), and -- on a GPU -- reproduces the
oracle through the reference's own class/method names."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cc")
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_mirror")


def build_cpp_test():
    from mgmol_b200 import build as b
    from oracle import oracle as orc
    b.build()
    orc.build(ref=False, port=True)
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC,
           "-L", os.path.join(ROOT, "mgmol_b200"), "-lmgmol_b200",
           "-L", os.path.join(ROOT, "oracle"), "-lmgmol_oracle",
           "-Wl,-rpath,$ORIGIN/../../mgmol_b200", "-Wl,-rpath,$ORIGIN/../../oracle",
           "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_cpp_host_header_compiles_without_cuda():
    exe = build_cpp_test()
    assert os.path.exists(exe)
    # without a GPU the program reports "no CUDA device" (exit 77), it does not
    # fall back to anything
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 77, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_cpp_host_mirror_parity():
    stale = (not os.path.exists(EXE)
             or os.path.getmtime(EXE) < os.path.getmtime(SRC)
             or os.path.getmtime(EXE) < os.path.getmtime(
                 os.path.join(ROOT, "include", "mgmol_b200.hpp")))
    exe = build_cpp_test() if stale else EXE
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok" in r.stdout


def test_cpp_anderson_mirror_matches_reference():
    """AndersonMix of include/mgmol_b200.hpp on host vectors against the
    trajectories of the reference's own AndersonMix<Solution> (golden)."""
    import numpy as np
    from mgmol_b200 import build as b
    b.build()
    src = os.path.join(ROOT, "tests", "cpp", "test_anderson_mirror.cc")
    exe = os.path.join(ROOT, "tests", "cpp", "test_anderson_mirror")
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src,
           "-L", os.path.join(ROOT, "mgmol_b200"), "-lmgmol_b200",
           "-Wl,-rpath,$ORIGIN/../../mgmol_b200", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_anderson.npz"))
    for n, m, beta in ((20, 1, 1.0), (20, 3, 1.0), (50, 4, 0.7)):
        ref = g["anderson_n%d_m%d_b%g" % (n, m, beta)]
        rng = np.random.default_rng(42)
        x = rng.uniform(0.0, 1.0, n)
        x[0], x[1] = 1.0, 2.0
        x /= np.linalg.norm(x)
        inp = "%d %d %r %d\n%s\n" % (n, m, beta, len(ref), " ".join(repr(float(v)) for v in x))
        r = subprocess.run([exe], input=inp, capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, r.stderr
        got = np.array([[float(v) for v in line.split()] for line in r.stdout.strip().splitlines()])
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max(), (n, m, beta)


def test_cpp_loewdin_transform_host():
    """loewdinTransform of include/mgmol_b200.hpp (Jacobi rotations in place of
    the reference's dsyev, src/GramMatrix.cc:267-299) against LAPACK: P =
    S^-1/2 and the inverse transform S^1/2, and refusal of an indefinite S."""
    import numpy as np
    from mgmol_b200 import build as b
    b.build()
    src = os.path.join(ROOT, "tests", "cpp", "test_loewdin_host.cc")
    exe = os.path.join(ROOT, "tests", "cpp", "test_loewdin_host")
    cmd = ["g++", "-std=c++11", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), src,
           "-L", os.path.join(ROOT, "mgmol_b200"), "-lmgmol_b200",
           "-Wl,-rpath,$ORIGIN/../../mgmol_b200", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr

    def run(S):
        n = S.shape[0]
        inp = "%d\n%s\n" % (n, " ".join(repr(float(v)) for v in S.T.ravel()))
        return subprocess.run([exe], input=inp, capture_output=True, text=True, timeout=120)

    rng = np.random.default_rng(11)
    for n, spread in ((1, 1.0), (2, 3.0), (7, 10.0), (40, 1e3), (130, 1e5)):
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        lam = np.geomspace(1.0, spread, n)
        S = (Q * lam) @ Q.T
        S = 0.5 * (S + S.T)
        r = run(S)
        assert r.returncode == 0, r.stdout + r.stderr
        rows = np.array([[float(v) for v in line.split()] for line in r.stdout.strip().splitlines()])
        P, R = rows[:n], rows[n:]
        w, V = np.linalg.eigh(S)
        assert np.abs(P - (V / np.sqrt(w)) @ V.T).max() <= 1e-12 * spread
        assert np.abs(R - (V * np.sqrt(w)) @ V.T).max() <= 1e-12 * np.sqrt(spread) * spread
        assert np.abs(P @ S @ P - np.eye(n)).max() <= 1e-11 * np.sqrt(spread)
        assert np.abs(P - P.T).max() <= 1e-13 * np.abs(P).max()
    bad = np.array([[1.0, 2.0], [2.0, 1.0]])
    r = run(bad)
    assert r.returncode == 3
    # the Loewdin part of the reference's own unit test (tests/testGramMatrix.cc:
    # 44-58,92-121): n = 213, diagonal 2, off-diagonal elements within 1/n;
    # B^-1/2 B^1/2 = 1 to 1e-14 and B^1/2 B^1/2 = B to 1e-12 (max norm)
    n = 213
    A = rng.uniform(-1.0 / n, 1.0 / n, (n, n))
    B = 0.5 * (A + A.T)
    np.fill_diagonal(B, 2.0)
    r = run(B)
    assert r.returncode == 0
    rows = np.array([[float(v) for v in line.split()] for line in r.stdout.strip().splitlines()])
    P, R = rows[:n], rows[n:]
    assert np.abs(P @ R - np.eye(n)).max() <= 1e-14
    assert np.abs(R @ R - B).max() <= 1e-12


# -- the C++ Poisson solvers (include/mgmol_b200_poisson.hpp) ---------------------
def _build_poisson_mirror():
    from mgmol_b200 import build as b
    from oracle import oracle as orc
    b.build()
    orc.build(ref=False, port=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_poisson_mirror.cc")
    exe = os.path.join(ROOT, "tests", "cpp", "test_poisson_mirror")
    cmd = ["g++", "-std=c++11", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "tests", "cpp"), src,
           "-L", os.path.join(ROOT, "mgmol_b200"), "-lmgmol_b200",
           "-L", os.path.join(ROOT, "oracle"), "-lmgmol_oracle",
           "-Wl,-rpath,$ORIGIN/../../mgmol_b200", "-Wl,-rpath,$ORIGIN/../../oracle", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _run_poisson_mirror(exe, mode, solver, case, lt, dt):
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from poisson_cases import DEFAULTS, charge, guess
    tag, dims, ll, bc, kw = case
    par = dict(DEFAULTS, **kw)
    head = "%d %d %d  %d %d %d  %r %r %r  %d %d %d  %d %d %d %r %d\n" % (
        (solver, lt, 1 if dt == np.float64 else 0) + tuple(dims) + tuple(ll) + tuple(bc)
        + (par["nu1"], par["nu2"], par["max_sweeps"], par["tol"], par["max_nlevels"]))
    body = "\n".join(repr(float(v)) for v in guess(dims, dt).ravel()) + "\n" + \
        "\n".join(repr(float(v)) for v in charge(dims, bc, dt).ravel()) + "\n"
    r = subprocess.run([exe, mode], input=head + body, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[:300], r.stderr[:300])
    lines = r.stdout.strip().splitlines()
    stats = [float(v) for v in lines[0].split()]
    vh = np.array([float(v) for v in lines[1:]]).reshape(dims)
    return vh, stats


def _check_poisson_mirror(mode, eps64, eps32, eps_pcg64, eps_pcg32, subset=False):
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from poisson_cases import CASES, PCG_CASES, key
    exe = _build_poisson_mirror()
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_poisson.npz"))
    plan = ((0, CASES), (1, PCG_CASES))
    laps = (0, 1, 2)
    if subset:      # one process (one CUDA context) per solve: keep the device leg short
        plan = ((0, [CASES[0], CASES[3]]), (1, [PCG_CASES[2]]))
        laps = (0, 2)
    for solver, cases in plan:
        for case in cases:
            for lt in laps:
                for dt in (np.float64, np.float32):
                    vh, st = _run_poisson_mirror(exe, mode, solver, case, lt, dt)
                    ref = g[key(case[0], lt, dt)]
                    rst = g[key(case[0], lt, dt) + "_stats"]
                    eps = ((eps64 if dt == np.float64 else eps32) if solver == 0
                           else (eps_pcg64 if dt == np.float64 else eps_pcg32))
                    assert np.abs(vh - ref).max() <= eps * np.abs(ref).max(), (case[0], lt, dt)
                    assert bool(st[0]) == bool(rst[0]), (case[0], lt, dt)
                    if solver == 0:
                        assert int(st[1]) == int(rst[1])
                        assert abs(st[2] - rst[2]) <= 1e-5 * rst[2]
                        assert abs(st[4] - rst[4]) <= 1e-5 * rst[4]
                    else:
                        assert abs(st[2] - rst[1]) <= 1e-4 * rst[1]
                        assert abs(st[4] - rst[2]) <= 1e-4 * rst[2]


def test_cpp_poisson_solvers_control_flow():
    """PoissonMG / PoissonPCG of include/mgmol_b200_poisson.hpp instantiated with
    host fields over the oracle's kernels, against the golden vectors of the
    compiled reference solvers (SolverLap/Mgm/Vcycle and PCGSolver)."""
    _check_poisson_mirror("cpu", 1e-14, 1e-6, 1e-10, 2e-6)


@pytest.mark.gpu
def test_cpp_poisson_solvers_on_device():
    """The same templates with the device field GridFunc<T> over the C ABI."""
    _check_poisson_mirror("gpu", 1e-13, 2e-6, 1e-9, 5e-6, subset=True)


def test_cpp_error_policy_throws_behind_c_entries():
    """the header templates end the run on a failed call (as the
    reference does); instantiated behind a C entry that returns codes
    (csrc/poisson.cu) they throw mgmol_b200::Error instead
    (MGMOL_B200_ERRORS_THROW) and the entry returns the MGB_E* code."""
    from mgmol_b200 import build as b
    b.build()
    src = os.path.join(ROOT, "tests", "cpp", "test_error_policy.cc")
    exe = os.path.join(ROOT, "tests", "cpp", "test_error_policy")
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src,
           "-L", os.path.join(ROOT, "mgmol_b200"), "-lmgmol_b200",
           "-Wl,-rpath,$ORIGIN/../../mgmol_b200", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "ok 3" in r.stdout
    # and the library's own instantiation is the throwing one
    with open(os.path.join(ROOT, "mgmol_b200", "csrc", "poisson.cu")) as f:
        text = f.read()
    assert text.index("#define MGMOL_B200_ERRORS_THROW") < text.index(
        '#include "mgmol_b200_poisson.hpp"')
    assert "catch (const mgmol_b200::Error&" in text


def _hartree_inputs(dims, bc, dt):
    import numpy as np
    from poisson_cases import charge, guess
    rho = charge(dims, bc, np.float64) + 0.3
    rhoc = np.full(dims, rho.mean()) + 0.05 * charge(dims, bc, np.float64, seed=11)
    if tuple(bc) == (1, 1, 1):
        rhoc += rho.mean() - rhoc.mean()          # neutral cell
    return guess(dims, dt), rho, rhoc


def _run_hartree_mirror(exe, mode, pcg, case, lt, dt):
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from poisson_cases import DEFAULTS
    tag, dims, ll, bc, kw = case
    par = dict(DEFAULTS, **kw)
    vh0, rho, rhoc = _hartree_inputs(dims, bc, dt)
    head = "%d %d %d  %d %d %d  %r %r %r  %d %d %d  %d %d %d %r %d\n" % (
        (3 if pcg else 2, lt, 1 if dt == np.float64 else 0) + tuple(dims) + tuple(ll) + tuple(bc)
        + (par["nu1"], par["nu2"], par["max_sweeps"], par["tol"], par["max_nlevels"]))
    body = "".join("\n".join(repr(float(v)) for v in a.ravel()) + "\n" for a in (vh0, rho, rhoc))
    r = subprocess.run([exe, mode], input=head + body, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[:300], r.stderr[:300])
    lines = r.stdout.strip().splitlines()
    stats = [[float(v) for v in lines[k].split()] for k in range(2)]
    return np.array([float(v) for v in lines[2:]]).reshape(dims), stats


def _check_hartree_mirror(mode, cases, eps64, eps32):
    """Hartree<Field, RField, Solver> of include/mgmol_b200_poisson.hpp, two consecutive
    solves (the second from the kept potential), against the Python host's Hartree over the
    oracle's field operations (mgmol_b200/poisson.py, itself compared with the device in
    tests/test_gpu_poisson.py): potential and the three integrals."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from mgmol_b200.host import Grid
    from mgmol_b200.poisson import Hartree
    from oracle.oracle import Port
    from poisson_cases import CASES, DEFAULTS, PCG_CASES
    from poisson_twin import field_factory
    exe = _build_poisson_mirror()
    port = Port()
    byname = {c[0]: c for c in CASES + PCG_CASES}
    for tag, lt, dt, pcg in cases:
        case = byname[tag]
        _, dims, ll, bc, kw = case
        par = dict(DEFAULTS, **kw)
        vh, stats = _run_hartree_mirror(exe, mode, pcg, case, lt, dt)
        vh0, rho, rhoc = _hartree_inputs(dims, bc, dt)
        h = Hartree(Grid(dims, ll, 1, bc), lt, dt, field=field_factory(port), pcg=pcg,
                    rho_dtype=np.float64, precond_dtype=np.float32)
        h.setup(par["nu1"], par["nu2"], par["max_sweeps"], par["tol"], par["max_nlevels"])
        h.set_vh(vh0)
        eps = eps64 if dt == np.float64 else eps32
        for k in range(2):
            conv = h.solve(rho, rhoc)
            want = [float(conv), h.IntVhRho_old(), h.IntVhRho(), h.IntVhRhoc()]
            scale = max(abs(w) for w in want[1:]) + 1e-300
            assert stats[k][0] == want[0], (tag, lt, dt, pcg, k)
            for a, b in zip(stats[k][1:4], want[1:]):
                assert abs(a - b) <= 10 * eps * scale, (tag, lt, dt, pcg, k, a, b)
        ref = np.asarray(h.vh()).reshape(dims).astype(np.float64)
        assert np.abs(vh - ref).max() <= eps * np.abs(ref).max(), (tag, lt, dt, pcg)


def test_cpp_hartree_control_flow():
    import numpy as np
    _check_hartree_mirror("cpu", [("per", 0, np.float64, False), ("mix", 2, np.float64, False),
                                  ("dir", 1, np.float32, False), ("pcg_per", 0, np.float64, True),
                                  ("pcg_mix", 2, np.float32, True)], 1e-13, 2e-6)


@pytest.mark.gpu
def test_cpp_hartree_on_device():
    import numpy as np
    _check_hartree_mirror("gpu", [("per", 0, np.float64, False), ("pcg_mix", 0, np.float64, True)],
                          1e-12, 5e-6)
