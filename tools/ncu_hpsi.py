#!/usr/bin/env python
"""One H psi launch per (dtype, lap) on an n^3 x orb block for an ncu capture
(development tool):  ncu --set full -k regex:k_hpsi -o out python tools/ncu_hpsi.py --n 256 --orb 64"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mgmol_b200 import host as H  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--orb", type=int, default=64)
ap.add_argument("--dtypes", default="f64,f32")
ap.add_argument("--laps", default="2,0")
a = ap.parse_args()
n = a.n
dims = (n, n, n)
v = torch.rand(dims, device="cuda", dtype=torch.float64) - 0.7
for d in a.dtypes.split(","):
    dt = torch.float64 if d == "f64" else torch.float32
    phi = torch.rand((a.orb,) + dims, device="cuda", dtype=dt) - 0.5
    hp = torch.empty_like(phi)
    for lap in (int(x) for x in a.laps.split(",")):
        grid = H.Grid(dims, (0.1835 * n,) * 3, H.ghosts_for(lap))
        H.LapFactory.createLap(grid, lap).applyWithPot(phi, v, hp)
    torch.cuda.synchronize()
    del phi, hp
