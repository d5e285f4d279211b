"""The CUDA Kleinman-Bylander kernels on the REFERENCE's projectors: ions built by the
compiled KBprojectorSparse from real pseudopotential files (tests/golden/reference_kb.npz,
tests/golden/make_golden_kb.py) -- 8 projectors per ion for the ONCV files (the 16-wide
instantiation of k_kb_psi), one for the Troullier-Martins file, overlapping and wrapping
balls -- against the reference's own kbpsi and H phi += V_nl phi: float bit for bit in the
scatter, the projections to the reordering of the sums."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

from kb_cases import CENTERS, DIMS, LL, SPECIES, fields, key, vel  # noqa: E402

TDT = {np.float64: torch.float64, np.float32: torch.float32}


def check(tag, dt):
    from mgmol_b200 import host as H
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_kb.npz"))
    grid = H.Grid(DIMS, LL, 1)
    assert abs(grid.vel() - vel()) <= 1e-15 * vel()
    kbp = H.KBProjectors(grid, TDT[dt])
    for j in range(len(CENTERS)):
        kbp.add_ion(g[key(tag, dt, "nlindex%d" % j)], g[key(tag, dt, "proj%d" % j)],
                    g[key(tag, dt, "coeff%d" % j)])
    kbp.commit()
    psi, h0 = fields(dt)
    kref = g[key(tag, dt, "kbpsi")]
    kb = kbp.computeKBpsi(torch.from_numpy(psi).cuda()).cpu().numpy()
    kerr = float(np.abs(kb - kref).max() / np.abs(kref).max())
    dh = torch.from_numpy(np.ascontiguousarray(h0[:1])).cuda()
    kbp.computeHnlPhiAndAdd2HPhi(torch.from_numpy(np.ascontiguousarray(kref[:, :1])).cuda(), dh)
    got = dh.cpu().numpy()[0]
    want = g[key(tag, dt, "hphi0")]
    nbad = int((got != want).sum())
    kbp.close()
    return kerr, nbad


@pytest.mark.gpu
@pytest.mark.parametrize("tag", [s[0] for s in SPECIES])
@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_cuda_kb_on_reference_projectors(tag, dt):
    kerr, nbad = check(tag, dt)
    assert kerr <= (1e-13 if dt == np.float64 else 1e-6), kerr
    assert nbad == 0, "H phi += V_nl phi differs from the reference in %d values" % nbad


if __name__ == "__main__":
    for tag in ("si_oncv", "si_tm"):
        for dt in (np.float32, np.float64):
            print(tag, np.dtype(dt).name, check(tag, dt), flush=True)
