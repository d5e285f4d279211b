#!/usr/bin/env python
"""Development check of the tcgen05 float contraction kernel (k_gemm_tn_umma) on one B200:
error against the exact (FP64) contraction and time, next to the mma.sync 3xTF32 kernel
(mode 2) and the DMMA kernel on widened operands (mode 1).  One JSON line per case.

    python tools/umma_tn_check.py [--big] [--trunc 1,0] [--ch 8] [--fold 16]
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


def run_case(N, K, positive, modes, time_it):
    g = torch.Generator(device="cuda").manual_seed(5 + N)
    a = torch.rand((N, K), generator=g, device="cuda", dtype=torch.float32)
    b = torch.rand((N, K), generator=g, device="cuda", dtype=torch.float32)
    if not positive:
        a -= 0.5
        b -= 0.5
    # exact references in double, chunked over K to bound memory
    exg = torch.zeros((N, N), device="cuda", dtype=torch.float64)
    exp = torch.zeros((N, N), device="cuda", dtype=torch.float64)
    step = 1 << 20
    for k0 in range(0, K, step):
        ad = a[:, k0:k0 + step].double()
        bd = b[:, k0:k0 + step].double()
        exg += ad @ ad.t()
        exp += ad @ bd.t()
    na = torch.sqrt(torch.diag(exg))
    nb = torch.sqrt((b.double() ** 2).sum(1)) if K <= (1 << 22) else None
    if nb is None:
        nb = torch.zeros(N, device="cuda", dtype=torch.float64)
        for k0 in range(0, K, step):
            nb += (b[:, k0:k0 + step].double() ** 2).sum(1)
        nb = torch.sqrt(nb)
    S = torch.empty((N, N), device="cuda", dtype=torch.float64)
    P = torch.empty((N, N), device="cuda", dtype=torch.float64)
    for label, mode, env in modes:
        for k_, v_ in env.items():
            os.environ[k_] = v_
        check(lib().mgb_set_f32_contraction(mode))
        try:
            def gram():
                check(lib().mgb_syrk_t(0, N, K, 1.0, a.data_ptr(), K, S.data_ptr(), N, None))

            def proj():
                check(lib().mgb_gemm_tn(0, N, N, K, 1.0, a.data_ptr(), K, b.data_ptr(), K, 0.0,
                                        P.data_ptr(), N, None))
            S.fill_(float("nan"))
            P.fill_(float("nan"))
            gram()
            proj()
            torch.cuda.synchronize()
            # C is column-major C(i, j) at i + j*ldc: as a torch (N, N) row-major view it is C^T
            eg = float(((S.t() - exg).abs() / torch.outer(na, na)).max())
            ep = float(((P.t() - exp).abs() / torch.outer(na, nb)).max())
            sym = bool(torch.equal(S, S.t()))
            rec = dict(case=label, N=N, K=K, positive=positive, gram_err=eg, proj_err=ep,
                       symmetric=sym)
            if time_it:
                tg, tp = timeit(gram), timeit(proj)
                rec.update(gram_ms=tg, gram_tflops=N * N * K / tg / 1e9, proj_ms=tp,
                           proj_tflops=2.0 * N * N * K / tp / 1e9)
            print(json.dumps(rec), flush=True)
        finally:
            check(lib().mgb_set_f32_contraction(0))
            for k_ in env:
                os.environ.pop(k_, None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--trunc", default="1,0")
    ap.add_argument("--ch", default="1")
    ap.add_argument("--fold", default="32")
    ap.add_argument("--ko", default="2")
    ap.add_argument("--mc", default="1")
    ap.add_argument("--legacy", action="store_true")
    ap.add_argument("--nosmall", action="store_true")
    ap.add_argument("--shapes", default="256x2097152,512x2097152,1024x884736")
    args = ap.parse_args()
    modes = []
    for ko, lk in [(k_, l_) for k_ in args.ko.split(",") for l_ in args.mc.split(",")]:
        for tr in args.trunc.split(","):
            for ch in args.ch.split(","):
                for fo in args.fold.split(","):
                    modes.append((f"umma ko={ko} mc={lk} chb={ch} trunc={tr} fold={fo}", 0,
                                  {"MGB_UMMA_TRUNC": tr, "MGB_UMMA_CHB": ch, "MGB_UMMA_FOLD": fo,
                                   "MGB_UMMA_KO": ko, "MGB_TN_MC": lk}))
    if args.legacy:
        modes += [("mma.sync 3xTF32", 2, {}), ("DMMA widened", 1, {})]
    small = [(128, 4096), (130, 6144), (256, 32768), (304, 65536), (37, 1680), (136, 4000)]
    for N, K in ([] if args.nosmall else small):
        for positive in (False, True):
            run_case(N, K, positive, modes, False)
    if args.big:
        for N, K in [tuple(int(x) for x in sh.split("x")) for sh in args.shapes.split(",")]:
            for positive in (False, True):
                run_case(N, K, positive, modes, True)


if __name__ == "__main__":
    main()
