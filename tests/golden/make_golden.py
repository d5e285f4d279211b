"""Generate the golden vectors under tests/golden/ from the reference's OWN
compiled sources (oracle/_ref/libmgmol_ref.so, built by `make -C oracle ref`
from /root/reference, unmodified).  Run in the build container only -- the GPU
box has no /root/reference; it checks against the committed .npz files.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.oracle import Ref, synthetic_orbitals, synthetic_potential  # noqa: E402


def main():
    R = Ref()
    out = {}
    dims, ll, N = (12, 8, 16), (3.0, 2.2, 4.1), 3
    v = synthetic_potential(dims)
    out["dims"] = np.array(dims)
    out["ll"] = np.array(ll)
    out["nfunc"] = np.array(N)
    for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
        phi = synthetic_orbitals(N, dims, dt)
        for lt in (0, 2):
            for bc in ((1, 1, 1), (0, 0, 0), (1, 0, 1)):
                key = "hpsi_lap%d_%s_bc%d%d%d" % ((lt, tag) + bc)
                out[key] = R.hpsi(lt, phi, v, ll, bc)
        for g in (1, 2):
            for bc in ((1, 1, 1), (0, 0, 0), (0, 1, 1)):
                out["trade_g%d_%s_bc%d%d%d" % ((g, tag) + bc)] = \
                    R.trade_boundaries(phi, g, bc, ll)
        h = tuple(l / d for l, d in zip(ll, dims))
        for kind, g in ((0, 1), (1, 1), (2, 2), (3, 3), (4, 4), (100, 1)):
            gv = R.trade_boundaries(phi, g, (1, 1, 1), ll)
            # 6th/8th: one function per call (the batched reference loop does
            # not reset its x offset between functions, FDkernels.cc:297-343)
            if kind in (3, 4):
                res = np.concatenate([R.fdkernel(kind, gv[i:i + 1], g, h) for i in range(N)])
            else:
                res = R.fdkernel(kind, gv, g, h)
            out["fd%d_%s" % (kind, tag)] = res
        gv = R.trade_boundaries(phi, 1, (1, 1, 1), ll)
        out["restrict_%s" % tag] = R.restrict3D(gv, 1, ll)
        out["extend_%s" % tag] = R.extend3D(out["restrict_%s" % tag], gv, 1, ll)
        for lt in (0, 2):
            for bc in ((1, 1, 1), (0, 0, 0)):
                for lev in (1, 2):
                    key = "precond_lap%d_lev%d_%s_bc%d%d%d" % ((lt, lev, tag) + bc)
                    out[key] = R.precond_mg(lt, lev, phi, ll, 0.37, bc)
        a = synthetic_orbitals(5, (6, 4, 8), dt)
        b = synthetic_orbitals(5, (6, 4, 8), dt, first=11)
        out["gemm_tn_%s" % tag] = R.gemm_tn(a, b, 0.37)
        out["syrk_%s" % tag] = R.syrk(a, 0.37)
        M = np.random.default_rng(7).standard_normal((5, 5))
        out["gemm_nn_%s" % tag] = R.gemm_nn(a, M)
    for lt in (0, 1, 2, 3, 4):
        g = {0: 1, 1: 1, 2: 2, 3: 3, 4: 4}[lt]
        # Laph6/Laph8 assert a near-isotropic mesh (src/pb/Laph6.h:33)
        d, l = (dims, ll) if lt < 3 else ((8, 8, 8), (2.0, 2.0, 2.0))
        out["lapconst_%d" % lt] = np.array(R.lap_constants(lt, d, g, l))
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote %d arrays" % len(out))


if __name__ == "__main__":
    main()
