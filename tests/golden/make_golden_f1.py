"""Golden vectors for the "next" rows from the reference's OWN
compiled sources: Lap::rhs per orbital (B of Laph4M, B2 of Laph4MP), the
residual sequence of MGmol::computeResidualUsingHPhi (rhs per orbital,
MPgemmNN with localT, axpy(-1, H phi)) and the density sequence of
Rho::computeRhoSubdomainUsingBlas3 (MPgemmNN, then the j-ordered accumulation,
restated in numpy because Rho.cc itself needs Control/Mesh).  Run in the build
container only.

    python tests/golden/make_golden_f1.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.oracle import Ref, synthetic_orbitals, synthetic_potential  # noqa: E402

DIMS, LL, N = (12, 8, 16), (3.0, 2.2, 4.1), 5


def theta():
    t = np.random.default_rng(3).standard_normal((N, N)) / np.sqrt(N)
    return 0.5 * (t + t.T)


def main():
    R = Ref()
    out = {"dims": np.array(DIMS), "ll": np.array(LL), "nfunc": np.array(N), "theta": theta()}
    v = synthetic_potential(DIMS)
    for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
        phi = synthetic_orbitals(N, DIMS, dt)
        for lt in (0, 10):
            for bc in ((1, 1, 1), (0, 0, 0)):
                out["rhs_lap%d_%s_bc%d%d%d" % ((lt, tag) + bc)] = R.lap_rhs(lt, phi, LL, bc)
        for lt in (0, 2):
            hphi = R.hpsi(lt, phi, v, LL)
            bphi = R.lap_rhs(lt, phi, LL) if lt == 0 else phi
            res = R.gemm_nn(bphi, theta())          # psi.multiplyByMatrix(localT, res)
            res = res - hphi                         # res.axpy(-1., hphi)
            out["residual_lap%d_%s" % (lt, tag)] = res.astype(dt)
        product = R.gemm_nn(phi, theta())
        rho = np.zeros(DIMS)
        for j in range(N):
            rho += (product[j] * phi[j]).astype(np.float64)
        out["rho_%s" % tag] = rho
    np.savez_compressed(os.path.join(HERE, "reference_f1.npz"), **out)
    print("wrote %d arrays" % len(out))


if __name__ == "__main__":
    main()
