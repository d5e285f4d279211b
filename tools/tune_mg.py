#!/usr/bin/env python
"""Sweep tile configurations of the fused Jacobi kernel (development tool)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mgmol_b200 import host as H
from tools.perf_probe import timeit

n, norb = 128, 256
cfgs = [None, "4,8,1,7,0", "4,8,1,9,0", "4,8,1,11,0", "4,4,3,7,0", "4,4,2,9,0", "4,4,2,12,0",
        "4,4,3,6,0", "4,8,2,6,0", "2,8,2,9,0", "2,16,1,9,0", "4,4,1,12,0", "4,2,4,9,0",
        "4,2,6,7,0", "2,4,4,9,0", "4,4,2,14,0"]
for lap in (2, 0):
    dims = (n, n, n)
    grid = H.Grid(dims, (0.1835 * n,) * 3, H.ghosts_for(lap))
    res = H.Orbitals(grid, norb, torch.float32)
    res.psi().copy_(torch.rand((norb,) + dims, device="cuda", dtype=torch.float32) - 0.5)
    pc = H.OrbitalsPreconditioning()
    pc.setup(res, 2, lap)
    pc.gamma_ = 0.3
    for cfg in cfgs:
        if cfg is None:
            os.environ.pop("MGB_MG_CFG", None)
        else:
            os.environ["MGB_MG_CFG"] = cfg
        try:
            med, best = timeit(lambda: pc.precond_mg(res), reps=5, warm=1)
            print(json.dumps({"lap": lap, "cfg": cfg, "ms": med, "best": best}), flush=True)
        except Exception as e:
            print(json.dumps({"lap": lap, "cfg": cfg, "error": str(e)[:100]}), flush=True)
    pc.close()
