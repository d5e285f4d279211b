"""Restart layout from / into a DEVICE-resident orbital block: column `color` of the block
is dataset Function%04d as it lies in HBM (src/ExtendedGridOrbitals.cc:734-741: one
copy_view_to_host per function)."""
import numpy as np
import pytest
import torch

from test_restart_cpu import FakeFile

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("odt, info", [(torch.float64, 4), (torch.float64, 3), (torch.float32, 3)])
def test_device_block_round_trip(odt, info):
    from mgmol_b200 import host as H
    from mgmol_b200.restart import read_func_hdf5, write_func_hdf5
    dims, N = (16, 8, 32), 4
    grid = H.Grid(dims, (4.0, 2.0, 8.0), 1)
    gen = torch.Generator(device="cuda").manual_seed(5)
    orb = H.Orbitals(grid, N, odt, torch.rand((N,) + dims, generator=gen, device="cuda", dtype=odt))
    f = FakeFile()
    assert write_func_hdf5(f, "Function", orb, out_restart_info=info) == N
    ftype = np.float64 if info > 3 else np.float32
    for c in range(N):
        assert np.array_equal(f["Function%04d" % c].data, orb.psi()[c].cpu().numpy().astype(ftype))
    back = H.Orbitals(grid, N, odt)
    back.incrementIterativeIndex()
    assert read_func_hdf5(f, "Function", back) == N
    assert back.getIterativeIndex() == 0
    expect = orb.psi().cpu().numpy().astype(ftype).astype(orb.psi().cpu().numpy().dtype)
    assert np.array_equal(back.psi().cpu().numpy(), expect)
