"""Row N1 on the GPU: K = 20 ABPG iterations (H psi -> theta -> residual -> precond_mg ->
update -> Loewdin, tests/abpg_proxy.py) through the C ABI against the golden trajectory of
the COMPILED REFERENCE (tests/golden/reference_trajectory.npz).  north_star's energy bar is
1e-8 Ha: asserted with the preconditioner in its reference-shaped (bit-identical) mode; the
fused float V-cycle differs from the reference's by ~1e-6 of its output, which moves the
intermediate iterates but not the fixed point -- its bar is stated below."""
import os

import numpy as np
import pytest
import torch

import abpg_proxy as ap

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_trajectory.npz")))


@pytest.mark.parametrize("name", ["sih4", "sih4_4th", "nanowire"])
@pytest.mark.parametrize("mode", [1, 0])
def test_abpg_trajectory_against_compiled_reference(name, mode):
    """mode 1: V-cycle in its reference-shaped (bit-identical) form; 0: automatic, i.e. the
    fused float kernels wherever the grid is eligible (the nanowire box; 40^3 coarsens to
    10^3, whose z rows are not whole 16-byte vectors, and stays reference-shaped)."""
    from mgmol_b200 import host as H
    c = ap.CASES[name]
    phi, v = ap.inputs(name)
    grid = H.Grid(c["dims"], c["ll"], H.ghosts_for(c["lap"]))
    inv_diag = H.LapFactory.createLap(grid, c["lap"]).invDiagEl()
    ops = ap.GpuOps(H, c["lap"], v, c["ll"], c["levels"], c["dims"], c["N"], mode, torch.float64)
    se, nr, _ = ap.run(ops, ops.to_dev(phi), float(v.max()), inv_diag, c["levels"], c["K"])
    ops.close()
    d_eps = np.abs(se - GOLD[name + "_sum_eps"])
    d_res = np.abs(nr - GOLD[name + "_norm_res"])
    print("%s mode %d: |d sum eps| first %.3e, max %.3e, final %.3e Ha; max |d res| = %.3e"
          % (name, mode, d_eps[0], d_eps.max(), d_eps[-1], d_res.max()))
    scale = max(1.0, c["N"] / 4.0)
    assert d_eps[0] <= 1e-11 * scale            # before any float rounding is amplified
    # along the way: the reference's own mixed-precision noise (abpg_proxy.py), and for the
    # fused V-cycle its 1e-6 relative difference from the reference's
    assert d_eps.max() <= (1e-6 if mode == 1 else 2e-5) * scale
    if c["converged"]:
        assert d_eps[-1] <= 1e-8                # north_star: 1e-8 Ha after the quench
        assert nr[-1] < 1e-6
    else:
        assert nr[-1] < 0.05 * nr[0]
