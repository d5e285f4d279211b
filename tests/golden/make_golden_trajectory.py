#!/usr/bin/env python
"""Golden trajectories of the ABPG proxy (tests/abpg_proxy.py) computed from the COMPILED
REFERENCE's kernels (oracle/_ref): per case the 20 values of sum_i eps_i and of the residual
norm.  Run here (needs /root/reference compiled: make -C oracle ref):

    python tests/golden/make_golden_trajectory.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import abpg_proxy as ap  # noqa: E402
from oracle.oracle import Ref  # noqa: E402

ref = Ref()
out = {}
for name, c in ap.CASES.items():
    phi, v = ap.inputs(name)
    _, inv_diag, _ = ref.lap_constants(c["lap"], c["dims"], 1 if c["lap"] != 2 else 2, c["ll"])
    ops = ap.CpuOps(ref, c["lap"], v, c["ll"], c["levels"], c["dims"])
    t0 = time.time()
    se, nr, _ = ap.run(ops, phi.copy(), float(v.max()), inv_diag, c["levels"], c["K"])
    print(name, "%.1f s" % (time.time() - t0), "sum eps", se[0], "->", se[-1], "res", nr[0], "->", nr[-1])
    out[name + "_sum_eps"] = se
    out[name + "_norm_res"] = nr
np.savez(os.path.join(ROOT, "tests", "golden", "reference_trajectory.npz"), **out)
