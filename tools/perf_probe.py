#!/usr/bin/env python
"""Per-kernel-group timing on one B200 (development tool, not the bench):
H psi variants, multigrid preconditioner, contractions, cuBLAS FP64 reference.
Prints one JSON line per measurement.

    python tools/perf_probe.py [--what hpsi,mg,gemm,peak] [--n 128] [--orb 256]
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


def timeit(fn, reps=10, warm=3):
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
    return float(np.median(ts)), float(min(ts))


def out(**kw):
    print(json.dumps(kw), flush=True)


def probe_hpsi(n, norb):
    for dt, S in ((torch.float64, 8), (torch.float32, 4)):
        for lap in (2, 0):
            dims = (n, n, n)
            grid = H.Grid(dims, (0.1835 * n,) * 3, H.ghosts_for(lap))
            phi = torch.rand((norb,) + dims, device="cuda", dtype=dt) - 0.5
            v = torch.rand(dims, device="cuda", dtype=torch.float64) - 0.7
            hp = torch.empty_like(phi)
            lapop = H.LapFactory.createLap(grid, lap)
            for path in (1, 2):
                check(lib().mgb_hpsi_force_path(path))
                try:
                    med, best = timeit(lambda: lapop.applyWithPot(phi, v, hp))
                except Exception as e:  # noqa: BLE001
                    out(kernel="hpsi", path=path, error=str(e))
                    continue
                finally:
                    lib().mgb_hpsi_force_path(0)
                upd = float(n) ** 3 * norb
                gbs = 2 * S * upd / (med * 1e-3) / 1e9
                out(kernel="hpsi", path=path, dtype=str(dt), lap=lap, n=n, orb=norb,
                    ms=med, ms_best=best, upd_per_s=upd / (med * 1e-3), gbs=gbs,
                    frac=gbs / PEAK)
            del phi, hp


def probe_mg(n, norb):
    for dt, S in ((torch.float64, 8), (torch.float32, 4)):
        for lap in (2, 0):
            dims = (n, n, n)
            grid = H.Grid(dims, (0.1835 * n,) * 3, H.ghosts_for(lap))
            res = H.Orbitals(grid, norb, dt)
            res.psi().copy_(torch.rand((norb,) + dims, device="cuda", dtype=dt) - 0.5)
            pc = H.OrbitalsPreconditioning()
            pc.setup(res, 2, lap)
            pc.gamma_ = 0.3
            l0 = lib().mgb_launch_count()
            pc.precond_mg(res)
            nl = lib().mgb_launch_count() - l0
            med, best = timeit(lambda: pc.precond_mg(res), reps=5, warm=1)
            upd = float(n) ** 3 * norb
            model = (74.0 + 2 * S) * upd
            out(kernel="precond_mg", dtype=str(dt), lap=lap, n=n, orb=norb, ms=med,
                ms_best=best, launches=int(nl), upd_per_s=upd / (med * 1e-3),
                model_gbs=model / (med * 1e-3) / 1e9,
                model_frac=model / (med * 1e-3) / 1e9 / PEAK)
            pc.close()
            del res


def probe_gemm(n, norb):
    npt = n ** 3
    for dt in (torch.float64, torch.float32):
        dims = (n, n, n)
        grid = H.Grid(dims, (0.1835 * n,) * 3, 1)
        a = H.Orbitals(grid, norb, dt)
        b = H.Orbitals(grid, norb, dt)
        a.psi().copy_(torch.rand((norb,) + dims, device="cuda", dtype=dt) - 0.5)
        b.psi().copy_(torch.rand((norb,) + dims, device="cuda", dtype=dt) - 0.5)
        med, best = timeit(lambda: a.computeGram(), reps=5, warm=2)
        fl = float(norb) * norb * npt  # syrk: N^2 K useful flops
        out(kernel="gram_syrk", dtype=str(dt), n=n, orb=norb, ms=med, ms_best=best,
            tflops_useful=fl / (med * 1e-3) / 1e12,
            # off-diagonal tiles whole, diagonal tiles 10/16 of a tile (the
            # half-work path) -> N^2 K (1 + 1/(4 T)), T = tile rows
            tflops_executed=fl * (1 + 0.25 / max(1, (norb + 127) // 128)) / (med * 1e-3) / 1e12)
        med, best = timeit(lambda: a.computeLocalProduct(b), reps=5, warm=2)
        out(kernel="gemm_tn", dtype=str(dt), n=n, orb=norb, ms=med, ms_best=best,
            tflops=2 * fl / (med * 1e-3) / 1e12)
        M = torch.rand((norb, norb), device="cuda", dtype=torch.float64) - 0.5
        med, best = timeit(lambda: a.multiplyByMatrix(M, b), reps=5, warm=2)
        out(kernel="gemm_nn", dtype=str(dt), n=n, orb=norb, ms=med, ms_best=best,
            tflops=2 * fl / (med * 1e-3) / 1e12)
        # library references for the same shapes (cuBLAS through torch)
        A2 = a.psi().reshape(norb, npt)
        B2 = b.psi().reshape(norb, npt)
        med, best = timeit(lambda: torch.matmul(A2, B2.t()), reps=5, warm=2)
        out(kernel="cublas_gemm_tn", dtype=str(dt), n=n, orb=norb, ms=med, ms_best=best,
            tflops=2 * fl / (med * 1e-3) / 1e12)
        Mt = M.to(dt)
        med, best = timeit(lambda: torch.matmul(Mt, A2), reps=5, warm=2)
        out(kernel="cublas_gemm_nn", dtype=str(dt), n=n, orb=norb, ms=med, ms_best=best,
            tflops=2 * fl / (med * 1e-3) / 1e12)
        del a, b


def probe_peak():
    for dt in (torch.float64, torch.float32):
        for m in (4096, 8192):
            x = torch.rand((m, m), device="cuda", dtype=dt)
            y = torch.rand((m, m), device="cuda", dtype=dt)
            torch.backends.cuda.matmul.allow_tf32 = False
            med, best = timeit(lambda: torch.matmul(x, y), reps=5, warm=2)
            out(kernel="cublas_square", dtype=str(dt), m=m, ms=med, ms_best=best,
                tflops=2.0 * m ** 3 / (best * 1e-3) / 1e12)
    # plain copy (the HBM denominator) measured the driver's way
    a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda")
    b = torch.empty_like(a)
    med, best = timeit(lambda: b.copy_(a), reps=10, warm=3)
    out(kernel="copy", gbs=2.0 * a.numel() * 2 / (best * 1e-3) / 1e9, ms_best=best)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="peak,hpsi,mg,gemm")
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--orb", type=int, default=256)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    for w in args.what.split(","):
        {"hpsi": probe_hpsi, "mg": probe_mg, "gemm": probe_gemm}.get(
            w, lambda n, o: probe_peak())(args.n, args.orb)


if __name__ == "__main__":
    main()
