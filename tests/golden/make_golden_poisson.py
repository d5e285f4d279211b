"""Golden vectors of the reference's Poisson solvers (SolverLap / Mgm / Vcycle, PCGSolver)
from the compiled reference (oracle/_ref), for the GPU box where /root/reference
is absent.  Run here:  python tests/golden/make_golden_poisson.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.oracle import Ref  # noqa: E402
from poisson_cases import (CASES, DEFAULTS, DTYPES, LAPS, PCG_CASES, charge, guess,  # noqa: E402
                           key)


def main():
    ref = Ref()
    out = {}
    for tag, dims, ll, bc, kw in CASES:
        par = dict(DEFAULTS, **kw)
        for lt in LAPS:
            for dt in DTYPES:
                vh, conv, stats = ref.poisson_solve(lt, guess(dims, dt), charge(dims, bc, dt), ll,
                                                    bc, **par)
                out[key(tag, lt, dt)] = vh
                out[key(tag, lt, dt) + "_stats"] = np.array((float(conv),) + stats)
    for tag, dims, ll, bc, kw in PCG_CASES:
        par = dict(DEFAULTS, **kw)
        for lt in LAPS:
            for dt in DTYPES:
                vh, conv, stats = ref.pcg_solve(lt, guess(dims, dt), charge(dims, bc, dt), ll, bc,
                                                **par)
                out[key(tag, lt, dt)] = vh
                out[key(tag, lt, dt) + "_stats"] = np.array((float(conv),) + stats)
    path = os.path.join(ROOT, "tests", "golden", "reference_poisson.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
