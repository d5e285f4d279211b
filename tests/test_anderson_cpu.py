"""CPU tests of the AndersonMix mirror (mgmol_b200/mixing.py)
against the reference's own AndersonMix<Solution> (golden vectors, and the
compiled reference where oracle/_ref exists)."""
import os

import numpy as np
import pytest

from anderson_case import run_case
from conftest import ROOT
from mgmol_b200.mixing import AndersonMix
from oracle.oracle import Ref


class Vec:
    """numpy vector with the interface AndersonMix needs; dotProduct carries the
    weight the reference's Solution::dotProduct has (invs)."""
    invs = 1.0

    def __init__(self, u):
        self.u = np.array(u, np.float64)

    def assign(self, y):
        self.u[:] = y.u

    def __isub__(self, y):
        self.u -= y.u
        return self

    def dotProduct(self, y):
        return Vec.invs * float(self.u @ y.u)

    def axpy(self, a, y):
        self.u += a * y.u

    def scal(self, a):
        self.u *= a


def mirror_mixer(m, beta):
    class Mixer:
        def __init__(self, x0):
            self.x = Vec(x0)
            self.work = Vec(x0)
            self.mix = AndersonMix(m, beta, self.x, lambda v: Vec(v.u))

        def update(self, x, f, invs):
            Vec.invs = invs
            fv = Vec(f)
            self.mix.update(fv, self.work)
            return self.x.u.copy(), fv.u.copy()
    return Mixer


@pytest.mark.parametrize("n,m,beta", [(20, 1, 1.0), (20, 3, 1.0), (50, 4, 0.7)])
def test_anderson_matches_golden(n, m, beta):
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_anderson.npz"))
    ref = g["anderson_n%d_m%d_b%g" % (n, m, beta)]
    xs = np.stack(run_case(n, mirror_mixer(m, beta), iters=len(ref)))
    # same LAPACK-free arithmetic up to the m x m solve: 1e-10 relative
    assert np.abs(xs - ref).max() <= 1e-10 * np.abs(ref).max()
    # the extrapolation does something: it differs from plain x += f
    plain = np.stack(run_case(n, mirror_mixer(0, beta), iters=len(ref)))
    assert np.abs(plain - ref).max() > 1e-3


@pytest.mark.skipif(not Ref.available(), reason="compiled reference not present")
def test_anderson_against_compiled_reference_random_problem():
    R = Ref()
    n, m, beta = 64, 5, 0.9

    class RefMixer:
        def __init__(self, x0):
            self.h = R.anderson_create(m, beta, x0)

        def update(self, x, f, invs):
            return R.anderson_update(self.h, f, invs)
    a = run_case(n, RefMixer, iters=15)
    b = run_case(n, mirror_mixer(m, beta), iters=15)
    assert np.abs(np.stack(a) - np.stack(b)).max() <= 1e-10 * np.abs(np.stack(a)).max()
