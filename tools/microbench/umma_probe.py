"""Driver of the tcgen05 kind::tf32 probe (tools/microbench/umma_tf32_probe.cu).
Development tool for the float contraction kernels' move from mma.sync to
tcgen05; not part of the library or the tests.  Needs a B200:

    gpurun -- python tools/microbench/umma_probe.py

Prints one JSON line per case: the error of C = A B^T against the float64
product, for TF32-exact operands (checks the descriptors: any layout mistake
shows as O(1) error), for general operands with one TF32 product (error ~1e-3,
the TF32 rounding) and with the 3xTF32 split (error ~1e-6, FP32 accumulation)."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "umma_tf32_probe.cu")
LIB = os.path.join(HERE, "umma_tf32_probe.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def build():
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-shared", "-Xcompiler", "-fPIC", "-o", LIB, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise SystemExit("nvcc failed")
    return LIB


def tf32_exact(x):
    """Round to 10 explicit mantissa bits (what kind::tf32 keeps)."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    u = (u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    return u.view(np.float32)


def main():
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("umma_probe.py: needs a GPU")
    L = ctypes.CDLL(build())
    L.umma_tf32_probe.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3
    rng = np.random.default_rng(0)
    ok = True
    for n in (32, 128, 256):
        for K in (32, 64, 512):
            a = rng.standard_normal((128, K)).astype(np.float32)
            b = rng.standard_normal((n, K)).astype(np.float32)
            for name, split, aa, bb, bar in (("tf32_exact_operands", 0, tf32_exact(a), tf32_exact(b), 1e-5),
                                             ("one_tf32_product", 0, a, b, 5e-3),
                                             ("3xtf32", 1, a, b, 1e-5)):
                da, db = torch.from_numpy(aa).cuda(), torch.from_numpy(bb).cuda()
                dc = torch.full((128, n), float("nan"), dtype=torch.float32, device="cuda")
                rc = L.umma_tf32_probe(da.data_ptr(), db.data_ptr(), dc.data_ptr(), n, K, split)
                ref = aa.astype(np.float64) @ bb.astype(np.float64).T
                scale = np.abs(aa).astype(np.float64) @ np.abs(bb).astype(np.float64).T
                got = dc.cpu().numpy().astype(np.float64)
                err = float(np.nanmax(np.abs(got - ref) / scale)) if rc == 0 else None
                good = rc == 0 and np.isfinite(got).all() and err <= bar
                ok = ok and good
                print(json.dumps({"case": name, "n": n, "K": K, "rc": rc, "rel_err": err,
                                  "bar": bar, "ok": bool(good)}))
    print("umma probe", "ok" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
