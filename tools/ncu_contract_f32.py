#!/usr/bin/env python
"""One Gram and one Phi^T (H Phi)-shaped float contraction on an n^3 x orb block, for an ncu
capture of the tcgen05 kernels:

    ncu --set full --clock-control none --import-source on -k regex:k_gemm_tn_umma \\
        -o gpurun_out/r02_umma -f python tools/ncu_contract_f32.py --n 128 --orb 256
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mgmol_b200._lib import lib, check  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=128)
ap.add_argument("--orb", type=int, default=256)
ap.add_argument("--nn", action="store_true", help="also Phi M (mgb_gemm_nn)")
a = ap.parse_args()
K, N = a.n ** 3, a.orb
phi = torch.rand((N, K), device="cuda", dtype=torch.float32) - 0.5
hphi = torch.rand((N, K), device="cuda", dtype=torch.float32) - 0.5
S = torch.empty((N, N), device="cuda", dtype=torch.float64)
check(lib().mgb_syrk_t(0, N, K, 1.0, phi.data_ptr(), K, S.data_ptr(), N, None))
check(lib().mgb_gemm_tn(0, N, N, K, 1.0, phi.data_ptr(), K, hphi.data_ptr(), K, 0.0, S.data_ptr(), N,
                        None))
if a.nn:
    M = torch.randn((N, N), device="cuda", dtype=torch.float64) / N ** 0.5
    check(lib().mgb_gemm_nn(0, K, N, N, 1.0, phi.data_ptr(), K, M.data_ptr(), N, 0.0,
                            hphi.data_ptr(), K, None))
torch.cuda.synchronize()
