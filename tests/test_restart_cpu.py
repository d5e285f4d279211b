"""Orbital restart I/O layout (src/ExtendedGridOrbitals.cc:641-884,
src/HDFrestart.cc:41-53, 777-789, 1654-1666) on a numpy stand-in for an h5py File: the
image has no HDF5 library, so what is pinned here is the LAYOUT the reference writes --
dataset names, shapes, element type, attributes, the hyperslab of a task, the stop at
the first missing function -- not the container's bytes."""
import numpy as np
import pytest
import torch

from mgmol_b200.host import Grid
from mgmol_b200.parallel import cart_coords, local_box
from mgmol_b200.restart import (getDatasetName, read_attributes, read_func_hdf5,
                                write_func_hdf5)


class FakeDataset:
    def __init__(self, shape, dtype, chunks=None):
        self.data = np.zeros(shape, dtype)
        self.shape, self.dtype, self.chunks = tuple(shape), np.dtype(dtype), chunks
        self.attrs = {}

    def __setitem__(self, key, value):
        assert np.asarray(value).dtype == self.dtype, "writeData hands over the file's type"
        self.data[key] = value

    def __getitem__(self, key):
        return self.data[key]


class FakeFile(dict):
    """The h5py.File subset the layout needs."""

    def create_dataset(self, name, shape, dtype, chunks=None):
        assert name not in self
        self[name] = FakeDataset(shape, dtype, chunks)
        return self[name]


class HostOrbitals:
    """host.Orbitals with the block in host memory (no GPU here): same members."""

    def __init__(self, grid, psi):
        self.grid_, self.psi_, self.iterative_index_ = grid, psi, 5

    def psi(self):
        return self.psi_

    def resetIterativeIndex(self):
        self.iterative_index_ = 0


def test_dataset_names_follow_the_reference():
    assert getDatasetName("Function", 0) == "Function0000"
    assert getDatasetName("Function", 7) == "Function0007"
    assert getDatasetName("Function", 42) == "Function0042"
    assert getDatasetName("Function", 999) == "Function0999"
    assert getDatasetName("Function", 1000) == "Function1000"
    assert getDatasetName("Function", 12345) == "Function12345"


@pytest.mark.parametrize("odt", [torch.float64, torch.float32])
@pytest.mark.parametrize("info", [3, 4])
def test_one_file_per_task_round_trip(odt, info):
    dims, ll, N = (8, 6, 10), (4.0, 3.0, 5.0), 5
    grid = Grid(dims, ll, 1)
    rng = np.random.default_rng(3)
    psi = torch.from_numpy(rng.standard_normal((N,) + dims)).to(odt)
    f = FakeFile()
    assert write_func_hdf5(f, "Function", HostOrbitals(grid, psi), out_restart_info=info,
                           origin=(-2.0, -1.5, -2.5)) == N
    assert sorted(f) == ["Function%04d" % c for c in range(N)]
    ftype = np.float64 if info > 3 else np.float32
    for c in range(N):
        d = f["Function%04d" % c]
        assert d.shape == dims and d.dtype == ftype and d.chunks is None
        # column c of the block in its own point order, cast like HDFrestart::writeData
        assert np.array_equal(d.data, psi[c].numpy().astype(ftype))
        a = read_attributes(f, "Function", c)
        assert a == {"gids": [c], "lattice": list(ll), "origin": [-2.0, -1.5, -2.5]}
        assert d.attrs["List of gids"].dtype == np.int32
        assert d.attrs["Lattice parameters"].dtype == np.float64
    back = HostOrbitals(grid, torch.zeros((N,) + dims, dtype=odt))
    assert read_func_hdf5(f, "Function", back) == N
    assert back.iterative_index_ == 0
    expect = psi.numpy().astype(ftype).astype(psi.numpy().dtype)
    assert np.array_equal(back.psi().numpy(), expect)


def test_read_stops_at_the_first_missing_function():
    dims, N = (4, 4, 4), 6
    grid = Grid(dims, (1.0, 1.0, 1.0), 1)
    psi = torch.arange(N * 64, dtype=torch.float64).reshape((N,) + dims)
    f = FakeFile()
    write_func_hdf5(f, "Function", HostOrbitals(grid, psi), out_restart_info=4)
    del f["Function0003"]
    back = HostOrbitals(grid, torch.full((N,) + dims, -1.0, dtype=torch.float64))
    assert read_func_hdf5(f, "Function", back) == N      # the reference returns numst_
    got = back.psi().numpy()
    assert np.array_equal(got[:3], psi.numpy()[:3])
    assert (got[3:] == -1.0).all()                        # 4 and 5 exist but are never read
    with pytest.raises(ValueError):
        read_func_hdf5(f, "Function", HostOrbitals(Grid((4, 4, 8), (1.0, 1.0, 1.0), 1),
                                                   torch.zeros((N, 4, 4, 8), dtype=torch.float64)))


def test_parallel_file_holds_the_global_mesh_in_task_blocks():
    """MGMOL_USE_HDF5P: datasets of the global mesh, chunked by the task block, each task
    writing (and reading back) its hyperslab at coord * dim."""
    gdims, ll, N, nproc = (8, 12, 4), (2.0, 3.0, 1.0), 3, (2, 3, 1)
    rng = np.random.default_rng(9)
    full = rng.standard_normal((N,) + gdims)
    f = FakeFile()
    world = nproc[0] * nproc[1] * nproc[2]
    for rank in range(world):
        coord = cart_coords(rank, nproc)
        grid = Grid(gdims, ll, 1, (1, 1, 1), nproc, coord)
        mine = torch.from_numpy(np.ascontiguousarray(full[(slice(None),) + local_box(gdims, nproc, coord)]))
        write_func_hdf5(f, "Function", HostOrbitals(grid, mine), out_restart_info=4, parallel=True)
    for c in range(N):
        d = f["Function%04d" % c]
        assert d.shape == gdims and d.chunks == (4, 4, 4)
        assert np.array_equal(d.data, full[c])
    for rank in (0, world - 1):
        coord = cart_coords(rank, nproc)
        grid = Grid(gdims, ll, 1, (1, 1, 1), nproc, coord)
        back = HostOrbitals(grid, torch.zeros((N, 4, 4, 4), dtype=torch.float64))
        read_func_hdf5(f, "Function", back, parallel=True)
        assert np.array_equal(back.psi().numpy(),
                              full[(slice(None),) + local_box(gdims, nproc, coord)])


def test_localized_orbitals_carry_their_gids_and_regions():
    """LocGridOrbitals::write_func_hdf5 (src/LocGridOrbitals.cc:1082-1117): per color the
    gids of its regions and 4 doubles (centre, radius) per region."""
    dims = (4, 4, 4)
    grid = Grid(dims, (1.0, 1.0, 1.0), 1)
    psi = torch.zeros((2,) + dims, dtype=torch.float64)
    f = FakeFile()
    write_func_hdf5(f, "Function", HostOrbitals(grid, psi), gids=[[3, 11], []],
                    centers_and_radii=[[0.1, 0.2, 0.3, 4.0, 0.5, 0.6, 0.7, 4.5], []])
    assert read_attributes(f, "Function", 0)["gids"] == [3, 11]
    assert f["Function0000"].attrs["List of centers and radii"].shape == (8,)
    assert read_attributes(f, "Function", 1)["gids"] == []          # writeGids skips natt = 0
    assert "List of centers and radii" not in f["Function0001"].attrs
