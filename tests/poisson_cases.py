"""Cases shared by the Poisson multigrid tests and the golden-vector generator."""
import numpy as np

# (tag, dims, lattice, bc, kwargs of SolverLap::setup)
CASES = [
    ("per", (24, 16, 12), (5.0, 4.0, 3.5), (1, 1, 1), {}),
    ("dir", (12, 20, 8), (3.0, 4.0, 2.0), (0, 0, 0), {}),
    ("mix", (12, 20, 8), (3.0, 4.0, 2.0), (1, 1, 0), {}),
    ("deep_per", (16, 16, 16), (4.0, 4.0, 4.0), (1, 1, 1), {}),       # down to 1^3
    ("deep_dir", (16, 16, 16), (4.0, 4.0, 4.0), (0, 1, 0), {}),
    ("conv", (24, 16, 12), (5.0, 4.0, 3.5), (1, 1, 1),
     dict(nu1=2, nu2=2, max_sweeps=30, tol=1e-4, max_nlevels=10)),
    ("shallow", (16, 16, 16), (4.0, 4.0, 4.0), (1, 1, 1),
     dict(nu1=1, nu2=3, max_sweeps=4, tol=1e-16, max_nlevels=2)),
]
# PCGSolver cases: few iterations, so that the comparison happens well above the
# rounding floor of the residual (a stagnating CG is not reproducible to more
# than its floor between two summation orders)
PCG_CASES = [
    ("pcg_per", (24, 16, 12), (5.0, 4.0, 3.5), (1, 1, 1), dict(max_sweeps=4)),
    ("pcg_dir", (12, 20, 8), (3.0, 4.0, 2.0), (0, 0, 0), dict(max_sweeps=4)),
    ("pcg_mix", (16, 16, 16), (4.0, 4.0, 4.0), (1, 1, 0), dict(max_sweeps=3)),
    ("pcg_mix2", (16, 16, 16), (4.0, 4.0, 4.0), (0, 1, 1), dict(max_sweeps=3, nu1=1, nu2=1)),
    ("pcg_conv", (24, 16, 12), (5.0, 4.0, 3.5), (1, 1, 1), dict(max_sweeps=30, tol=1e-5)),
    ("pcg_shallow", (16, 16, 16), (4.0, 4.0, 4.0), (1, 1, 1), dict(max_sweeps=3, max_nlevels=1)),
]
LAPS = (0, 1, 2)
DTYPES = (np.float64, np.float32)
DEFAULTS = dict(nu1=2, nu2=2, max_sweeps=10, tol=1e-16, max_nlevels=10)


def charge(dims, bc, dtype, seed=7):
    """A smooth-plus-noise charge; zero mean when the box is fully periodic."""
    rng = np.random.default_rng(seed + sum(dims))
    x, y, z = np.meshgrid(*(np.arange(n) / n for n in dims), indexing="ij")
    rho = np.sin(2 * np.pi * x) * np.cos(4 * np.pi * y) + 0.5 * np.cos(2 * np.pi * (z + 0.1))
    rho = rho + 0.2 * rng.standard_normal(dims)
    if tuple(bc) == (1, 1, 1):
        rho = rho - rho.mean()
    return np.ascontiguousarray(rho, dtype)


def guess(dims, dtype, seed=3):
    rng = np.random.default_rng(seed + sum(dims))
    return np.ascontiguousarray(0.01 * rng.standard_normal(dims), dtype)


def key(tag, lap_type, dtype):
    return "%s_lap%d_%s" % (tag, lap_type, "f64" if dtype == np.float64 else "f32")
