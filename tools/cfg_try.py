#!/usr/bin/env python
"""Time explicit MGB_HPSI_CFG / MGB_HPSI_ORDER settings (development tool):
python tools/cfg_try.py --n 256 --orb 64 --dtype f64 --lap 2 "4,2,1,3,64:0" "4,2,1,3,0:0" ..."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mgmol_b200 import host as H  # noqa: E402
from mgmol_b200._lib import lib, check  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--orb", type=int, default=64)
ap.add_argument("--dtype", default="f64")
ap.add_argument("--lap", type=int, default=2)
ap.add_argument("cfgs", nargs="*")
a = ap.parse_args()
PEAK = 6550.4
n = a.n
dims = (n, n, n)
dt = torch.float64 if a.dtype == "f64" else torch.float32
S = 8 if a.dtype == "f64" else 4
v = torch.rand(dims, device="cuda", dtype=torch.float64) - 0.7
phi = torch.rand((a.orb,) + dims, device="cuda", dtype=dt) - 0.5
hp = torch.empty_like(phi)
grid = H.Grid(dims, (0.1835 * n,) * 3, H.ghosts_for(a.lap))
lapop = H.LapFactory.createLap(grid, a.lap)
check(lib().mgb_hpsi_force_path(1))
for c in a.cfgs + ["default"]:
    if c == "default":
        os.environ.pop("MGB_HPSI_CFG", None)
        os.environ.pop("MGB_HPSI_ORDER", None)
    else:
        cfg, order = c.split(":")
        os.environ["MGB_HPSI_CFG"] = cfg
        os.environ["MGB_HPSI_ORDER"] = order
    for _ in range(2):
        lapop.applyWithPot(phi, v, hp)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lapop.applyWithPot(phi, v, hp)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    gbs = 2 * S * float(n) ** 3 * a.orb / (ms * 1e-3) / 1e9
    print(json.dumps({"cfg": c, "dtype": a.dtype, "lap": a.lap, "orb": a.orb, "ms": ms, "gbs": gbs,
                      "frac": gbs / PEAK}), flush=True)
