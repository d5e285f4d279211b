#!/usr/bin/env python
"""One fused V-cycle (precond_mg, 2 levels) and one FP64 Gram / Phi^T (H Phi) / Phi M on an
n^3 x orb block, for an ncu capture of those kernels:

    ncu --set full --clock-control none --import-source on -k regex:"k_mg_|k_gemm_|k_tn_fixup" \\
        -o gpurun_out/r02_pieces -f python tools/ncu_pieces.py --n 128 --orb 256
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mgmol_b200 import host as H  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=128)
ap.add_argument("--orb", type=int, default=256)
ap.add_argument("--lap", type=int, default=2)
ap.add_argument("--dtype", default="f64")
a = ap.parse_args()
n, norb, lap = a.n, a.orb, a.lap
tdt = torch.float64 if a.dtype == "f64" else torch.float32
dims = (n, n, n)
grid = H.Grid(dims, (0.1835 * n,) * 3, H.ghosts_for(lap))
phi = H.Orbitals(grid, norb, tdt, torch.rand((norb,) + dims, device="cuda", dtype=tdt) - 0.5)
res = H.Orbitals(grid, norb, tdt, torch.rand((norb,) + dims, device="cuda", dtype=tdt) - 0.5)
pc = H.OrbitalsPreconditioning()
pc.setup(res, 2, lap)
pc.gamma_ = 0.3
pc.precond_mg(res)
S = phi.computeGram()
P = phi.computeLocalProduct(res)
M = torch.rand((norb, norb), device="cuda", dtype=torch.float64) - 0.5
phi.multiplyByMatrix(M, res)
torch.cuda.synchronize()
pc.close()
