#!/usr/bin/env python
"""Development check of the tcgen05 Phi*M kernel (k_gemm_nn_umma) on one B200: error against
the exact (FP64) product and time, next to the mma.sync 3xTF32 kernel (mode 2) and the DMMA
kernel on widened operands (mode 1).  One JSON line per case.

    python tools/umma_nn_check.py [--big]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mgmol_b200._lib import lib, check  # noqa: E402


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


def run_case(m, n, k, positive, time_it, beta=0.0):
    g = torch.Generator(device="cuda").manual_seed(7 + n)
    a = torch.rand((k, m), generator=g, device="cuda", dtype=torch.float32)
    mc = torch.rand((n, k), generator=g, device="cuda", dtype=torch.float64)
    if not positive:
        a -= 0.5
        mc -= 0.5
    out0 = torch.rand((n, m), generator=g, device="cuda", dtype=torch.float32)
    alpha = 0.75
    ex = torch.empty((n, m), device="cuda", dtype=torch.float64)
    sc = torch.empty((n, m), device="cuda", dtype=torch.float64)
    step = 1 << 18
    for p0 in range(0, m, step):
        ad = a[:, p0:p0 + step].double()
        ex[:, p0:p0 + step] = alpha * (mc @ ad) + beta * out0[:, p0:p0 + step].double()
        sc[:, p0:p0 + step] = alpha * (mc.abs() @ ad.abs())
    for label, mode in (("tcgen05", 0), ("mma.sync 3xTF32", 2), ("DMMA widened", 1)):
        check(lib().mgb_set_f32_contraction(mode))
        try:
            out = out0.clone()

            def f():
                check(lib().mgb_gemm_nn(0, m, n, k, alpha, a.data_ptr(), m, mc.data_ptr(), k, beta,
                                        out.data_ptr(), m, None))
            f()
            torch.cuda.synchronize()
            err = float(((out.double() - ex).abs() / (sc + 1e-30)).max())
            errmax = float((out.double() - ex).abs().max() / ex.abs().max())
            rec = dict(case=label, m=m, n=n, k=k, positive=positive, beta=beta, err_vs_abs_sum=err,
                       err_vs_max=errmax)
            if time_it:
                out.copy_(out0)
                t = timeit(f)
                rec.update(ms=t, tflops=2.0 * m * n * k / t / 1e9,
                           hbm_GBps=(4.0 * m * k + 4.0 * m * n) / t / 1e6)
            print(json.dumps(rec), flush=True)
        finally:
            check(lib().mgb_set_f32_contraction(0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    args = ap.parse_args()
    for m, n, k in ((4096, 128, 128), (5000, 130, 130), (32768, 256, 256),
                    (8192, 70, 130), (128 * 9 + 7, 264, 200)):
        for positive in (False, True):
            run_case(m, n, k, positive, False)
    run_case(4096, 128, 128, False, False, beta=-0.5)
    if args.big:
        for m, n in ((128 ** 3, 256), (128 ** 3, 512), (96 ** 3, 1024)):
            for positive in (False, True):
                run_case(m, n, n, positive, True)


if __name__ == "__main__":
    main()
