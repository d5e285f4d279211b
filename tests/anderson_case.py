"""The iteration shared by the Anderson golden generator and the tests: power-
like iteration for the lowest eigenvector of diag(0, 1/(n-1), ..., 1) with the
residual lambda x - A x, as tests/Anderson/testAndersonMix.cc of the reference
does (residual with the Rayleigh quotient, scale by 1/max eigenvalue,
AndersonMix::update; the dot product carries the weight 1/<x,x> like
Solution::setInvS there)."""
import numpy as np


def run_case(n, make_mixer, iters):
    rng = np.random.default_rng(42)
    x = rng.uniform(0.0, 1.0, n)
    x[0], x[1] = 1.0, 2.0
    x /= np.linalg.norm(x)
    diag = np.arange(n) / (n - 1.0)
    mixer = make_mixer(x.copy())
    xs = []
    for _ in range(iters):
        invs = 1.0 / float(x @ x)
        r = diag * x
        lam = float(x @ r) / float(x @ x)
        r = (lam * x - r) / diag[-1]
        # the mixer owns the trial solution (AndersonMix::x_ is a reference to it)
        x, _ = mixer.update(x, r, invs)
        xs.append(x.copy())
    return xs
