#!/usr/bin/env python
"""A/B timing of mgb_hpsi from two builds of the library on the same GPU
(development tool): python tools/ab_hpsi.py libA.so libB.so"""
import ctypes
import sys

import numpy as np
import torch


class G(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_int * 3), ("gdim", ctypes.c_int * 3), ("ghosts", ctypes.c_int),
                ("h", ctypes.c_double * 3), ("bc", ctypes.c_int * 3), ("nproc", ctypes.c_int * 3),
                ("coord", ctypes.c_int * 3)]


def main():
    n, N = 128, 256
    libs = [ctypes.CDLL(p) for p in sys.argv[1:]]
    for dt, code in ((torch.float64, 1), (torch.float32, 0)):
        phi = torch.rand((N, n, n, n), device="cuda", dtype=dt) - 0.5
        out = torch.empty_like(phi)
        v = torch.rand((n, n, n), device="cuda", dtype=torch.float64) - 0.7
        for lap, g in ((2, 2), (0, 1)):
            gr = G((n, n, n), (n, n, n), g, (0.1835,) * 3, (1, 1, 1), (1, 1, 1), (0, 0, 0))
            res = []
            for rep in range(3):
                for L in libs:
                    def run():
                        rc = L.mgb_hpsi(lap, code, ctypes.byref(gr), ctypes.c_void_p(phi.data_ptr()),
                                        ctypes.c_size_t(n ** 3), ctypes.c_void_p(v.data_ptr()),
                                        ctypes.c_void_p(out.data_ptr()), ctypes.c_size_t(n ** 3), N,
                                        None, None, None)
                        assert rc == 0
                    for _ in range(3):
                        run()
                    torch.cuda.synchronize()
                    ts = []
                    for _ in range(10):
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record()
                        run()
                        b.record()
                        torch.cuda.synchronize()
                        ts.append(a.elapsed_time(b))
                    res.append(float(np.median(ts)))
            print(dt, "lap", lap, " ".join("%.4f" % t for t in res), flush=True)


if __name__ == "__main__":
    main()
