#!/usr/bin/env python
"""Tile-configuration search for the fused H psi kernel on one B200
(development tool): times every valid MGB_HPSI_CFG = RY,YG,NB,S,XC and both CTA
orders on an n^3 box, prints one JSON line per (dtype, lap, cfg, order).

    python tools/cfg_sweep.py [--n 256] [--orb 128] [--full 512]
"""
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

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def valid(ry, yg, nb, s, g, ny, nz, es):
    ty = ry * yg
    if ty > ny or ny % ty or ty < g:
        return False
    vec = 16 // es
    nzv = nz // vec
    tpo = -(-yg * nzv // 32) * 32
    if 32 + nb * tpo > 544:
        return False
    rb = nz * es
    r128 = lambda a: -(-a // 128) * 128  # noqa: E731
    tile = r128(g * rb) * 2 + r128(ty * rb)
    smem = 1024 + s * (nb + 1) * tile
    return smem <= 227 * 1024


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--orb", type=int, default=128)
    ap.add_argument("--full", type=int, default=0, help="re-time the best configs with this many orbitals")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    n = args.n
    dims = (n, n, n)
    v = torch.rand(dims, device="cuda", dtype=torch.float64) - 0.7
    best = {}
    for dt, S in ((torch.float64, 8), (torch.float32, 4)):
        phi = torch.rand((args.orb,) + dims, device="cuda", dtype=dt) - 0.5
        hp = torch.empty_like(phi)
        for lap in (2, 0):
            g = H.ghosts_for(lap)
            grid = H.Grid(dims, (0.1835 * n,) * 3, g)
            lapop = H.LapFactory.createLap(grid, lap)
            check(lib().mgb_hpsi_force_path(1))
            rows = []
            for ry in ((4, 2) if lap == 2 else (8, 4)):
                for yg in range(1, 33):
                    for nb in (1, 2, 3, 4):
                        for s in (2, 3, 4):
                            if not valid(ry, yg, nb, s, g, n, n, S):
                                continue
                            for order in (0, 1):
                                os.environ["MGB_HPSI_CFG"] = "%d,%d,%d,%d,0" % (ry, yg, nb, s)
                                os.environ["MGB_HPSI_ORDER"] = str(order)
                                try:
                                    ms = timeit(lambda: lapop.applyWithPot(phi, v, hp), reps=3, warm=1)
                                except Exception as e:  # noqa: BLE001
                                    print(json.dumps({"cfg": [ry, yg, nb, s], "error": str(e)}))
                                    continue
                                gbs = 2 * S * float(n) ** 3 * args.orb / (ms * 1e-3) / 1e9
                                rows.append((gbs, ry, yg, nb, s, order, ms))
            rows.sort(reverse=True)
            for r in rows[:8]:
                print(json.dumps({"dtype": str(dt), "lap": lap, "cfg": list(r[1:5]), "order": r[5],
                                  "ms": r[6], "gbs": r[0], "frac": r[0] / PEAK}), flush=True)
            best[(str(dt), lap)] = rows[:3]
            os.environ.pop("MGB_HPSI_CFG", None)
            os.environ.pop("MGB_HPSI_ORDER", None)
            ms = timeit(lambda: lapop.applyWithPot(phi, v, hp), reps=5, warm=1)
            gbs = 2 * S * float(n) ** 3 * args.orb / (ms * 1e-3) / 1e9
            print(json.dumps({"dtype": str(dt), "lap": lap, "cfg": "library default", "ms": ms,
                              "gbs": gbs, "frac": gbs / PEAK}), flush=True)
            lib().mgb_hpsi_force_path(0)
        del phi, hp
    if args.full:
        for dt, S in ((torch.float64, 8), (torch.float32, 4)):
            phi = torch.rand((args.full,) + dims, device="cuda", dtype=dt) - 0.5
            hp = torch.empty_like(phi)
            for lap in (2, 0):
                g = H.ghosts_for(lap)
                grid = H.Grid(dims, (0.1835 * n,) * 3, g)
                lapop = H.LapFactory.createLap(grid, lap)
                for r in best[(str(dt), lap)] + [None]:
                    if r is None:
                        os.environ.pop("MGB_HPSI_CFG", None)
                        os.environ.pop("MGB_HPSI_ORDER", None)
                    else:
                        os.environ["MGB_HPSI_CFG"] = "%d,%d,%d,%d,0" % r[1:5]
                        os.environ["MGB_HPSI_ORDER"] = str(r[5])
                    ms = timeit(lambda: lapop.applyWithPot(phi, v, hp), reps=5, warm=2)
                    gbs = 2 * S * float(n) ** 3 * args.full / (ms * 1e-3) / 1e9
                    print(json.dumps({"full": args.full, "dtype": str(dt), "lap": lap,
                                      "cfg": list(r[1:5]) if r else "library default",
                                      "order": r[5] if r else None, "ms": ms, "gbs": gbs,
                                      "frac": gbs / PEAK}), flush=True)
            del phi, hp


if __name__ == "__main__":
    main()
