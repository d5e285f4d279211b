"""Golden vectors of the reference's AndersonMix<Solution> (src/AndersonMix.cc
compiled unmodified, oracle/Makefile): the trial solutions of a fixed-point
iteration on a diagonal eigenproblem, the shape of the reference's own
tests/Anderson/testAndersonMix.cc.  Run in the build container only.

    python tests/golden/make_golden_anderson.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from anderson_case import run_case  # noqa: E402
from oracle.oracle import Ref  # noqa: E402


def main():
    R = Ref()
    out = {}
    for n, m, beta in ((20, 1, 1.0), (20, 3, 1.0), (50, 4, 0.7)):
        class Mixer:
            def __init__(self, x0):
                self.h = R.anderson_create(m, beta, x0)

            def update(self, x, f, invs):
                return R.anderson_update(self.h, f, invs)
        xs = run_case(n, Mixer, iters=12)
        out["anderson_n%d_m%d_b%g" % (n, m, beta)] = np.stack(xs)
    np.savez_compressed(os.path.join(HERE, "reference_anderson.npz"), **out)
    print("wrote %d arrays" % len(out))


if __name__ == "__main__":
    main()
