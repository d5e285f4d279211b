#!/usr/bin/env python
"""One H psi launch per (dtype, lap) on an n^3 x orb block, for an ncu capture; writes a
sidecar (one JSON line per launch, in launch order: box, dtype, operator, orbital count and
the kernel signature mgb_hpsi_last_kernel reports) that tools/traffic_table.py pairs with
the kernels of the report.

    ncu --set full --clock-control none --import-source on -k regex:k_hpsi_tma \\
        -o gpurun_out/r02_hpsi256 -f python tools/ncu_hpsi.py --n 256 --orb 64 \\
        --sidecar gpurun_out/r02_hpsi256.launches.jsonl
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mgmol_b200 import host as H  # noqa: E402
from mgmol_b200._lib import lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--orb", type=int, default=64)
ap.add_argument("--dtypes", default="f64,f32")
ap.add_argument("--laps", default="0,2")
ap.add_argument("--sidecar", default=None)
a = ap.parse_args()
n = a.n
dims = (n, n, n)
v = torch.rand(dims, device="cuda", dtype=torch.float64) - 0.7
rows = []
for d in a.dtypes.split(","):
    dt = torch.float64 if d == "f64" else torch.float32
    phi = torch.rand((a.orb,) + dims, device="cuda", dtype=dt) - 0.5
    hp = torch.empty_like(phi)
    for lap in (int(x) for x in a.laps.split(",")):
        grid = H.Grid(dims, (0.1835 * n,) * 3, H.ghosts_for(lap))
        H.LapFactory.createLap(grid, lap).applyWithPot(phi, v, hp)
        rows.append({"grid": list(dims), "dtype": d, "lap_type": lap, "orbitals": a.orb,
                     "kernel": lib().mgb_hpsi_last_kernel().decode()})
    torch.cuda.synchronize()
    del phi, hp
if a.sidecar:
    with open(a.sidecar, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
