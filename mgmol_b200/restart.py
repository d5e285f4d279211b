"""Orbital restart I/O layout: what ExtendedGridOrbitals /
LocGridOrbitals::write_func_hdf5 and read_func_hdf5 put into and take out of a
restart file (src/ExtendedGridOrbitals.cc:641-777, 780-884;
src/LocGridOrbitals.cc:1040-1200), for device-resident orbital blocks.

One dataset per storage function ("color"):

    name        getDatasetName(name, color) = name + "%04d" (src/HDFrestart.cc:41-53),
                e.g. Function0007
    shape       3-D.  One file per task (or HDF5 serial): the task's block
                (dim0, dim1, dim2).  Parallel HDF5 (MGMOL_USE_HDF5P): the global
                mesh (gdim0, gdim1, gdim2), chunked by the task block, every task
                writing the hyperslab at offset coord * dim
                (HDFrestart::createFilespace / createMemspace / createPlist,
                src/HDFrestart.h:196-296)
    type        float when out_restart_info <= 3, double above (Orbitals::
                outHdfDataType, src/Orbitals.h:72-77; HDFrestart::writeData converts
                with a (float) / (double) cast, src/HDFrestart.cc:1654-1666)
    attributes  "List of gids" (int), "Lattice parameters" and "Cell origin"
                (3 doubles each); LocGridOrbitals adds "List of centers and radii"
                (4 doubles per region)
    values      the function on the block in the orbital block's own point order,
                x slowest -- i.e. column `color` of the resident block as it is:
                the reference's copy_view_to_host(psi(color)) is one device-to-host
                copy here, no repacking.

The container is any object with h5py's File interface (create_dataset, `in`,
[], dataset slicing, .attrs).  `open_restart` opens a real HDF5 file through h5py
when that is installed; this image has no HDF5 library, so the tests drive the
layout with a numpy stand-in for the File object and a real file is never written
here (stated in DESIGN.md: the container format itself is the HDF5 library's)."""
import numpy as np
import torch


def getDatasetName(name, color):
    """src/HDFrestart.cc:41-53."""
    zeros = int(color < 10) + int(color < 100) + int(color < 1000)
    return name + "0" * zeros + str(color)


def open_restart(path, mode="r"):
    try:
        import h5py
    except ImportError as e:
        raise ImportError("mgmol_b200.restart.open_restart needs h5py (HDF5); pass any "
                          "h5py-compatible file object to write_func_hdf5 / read_func_hdf5 "
                          "instead") from e
    return h5py.File(path, mode)


def _block_slices(grid, parallel):
    """Where the task's block lies in the dataset."""
    if not parallel:
        return tuple(slice(0, n) for n in grid.shape())
    return tuple(slice(c * n, (c + 1) * n) for c, n in zip(grid.coord, grid.shape()))


def write_func_hdf5(h5file, name, orbitals, out_restart_info=3, origin=(0., 0., 0.),
                    parallel=False, gids=None, centers_and_radii=None):
    """ExtendedGridOrbitals::write_func_hdf5 (src/ExtendedGridOrbitals.cc:641-777).
    orbitals: host.Orbitals (device block) or a (numst, nx, ny, nz) tensor / array.
    gids: per color the list of global ids (default [color], the extended-orbitals
    case; LocGridOrbitals passes its ColoredRegions lists).  parallel: one shared file,
    datasets of the global mesh, this task's hyperslab (the caller's h5py file is then
    opened with the mpio driver, as MGmol opens it with H5Pset_fapl_mpio).
    Returns the number of datasets written."""
    grid = orbitals.grid_
    psi = orbitals.psi()
    numst = psi.shape[0]
    precision = 2 if out_restart_info > 3 else 1
    ftype = np.float64 if precision == 2 else np.float32
    shape = tuple(grid.gdim_) if parallel else tuple(grid.shape())
    chunks = tuple(grid.shape()) if parallel else None
    sl = _block_slices(grid, parallel)
    for color in range(numst):
        dsname = getDatasetName(name, color)
        if dsname in h5file:
            dset = h5file[dsname]
        else:
            kw = {"chunks": chunks} if chunks else {}
            dset = h5file.create_dataset(dsname, shape=shape, dtype=ftype, **kw)
            glist = [color] if gids is None else list(gids[color])
            if centers_and_radii is not None and len(centers_and_radii[color]):
                dset.attrs["List of centers and radii"] = np.asarray(
                    centers_and_radii[color], np.float64).reshape(-1)
            if glist:
                dset.attrs["List of gids"] = np.asarray(glist, np.int32)
            dset.attrs["Lattice parameters"] = np.asarray(grid.ll_, np.float64)
            dset.attrs["Cell origin"] = np.asarray(origin, np.float64)
        # copy_view_to_host(psi(color)) + the cast of HDFrestart::writeData
        col = psi[color]
        host = col.detach().cpu().numpy() if isinstance(col, torch.Tensor) else np.asarray(col)
        dset[sl] = host.astype(ftype)
    return numst


def read_func_hdf5(h5file, name, orbitals, parallel=False):
    """ExtendedGridOrbitals::read_func_hdf5 (src/ExtendedGridOrbitals.cc:780-884): reads
    Function0000, Function0001, ... into the columns of the resident block until a
    dataset is missing (the loop breaks there, :832-834), converts to ORBDTYPE, resets the
    iterative index.  Returns numst like the reference."""
    grid = orbitals.grid_
    psi = orbitals.psi()
    numst = psi.shape[0]
    sl = _block_slices(grid, parallel)
    for color in range(numst):
        dsname = getDatasetName(name, color)
        if dsname not in h5file:
            break
        dset = h5file[dsname]
        want = tuple(grid.gdim_) if parallel else tuple(grid.shape())
        if tuple(dset.shape) != want:
            raise ValueError("read_func_hdf5: dataset %s has shape %s, the mesh is %s"
                             % (dsname, tuple(dset.shape), want))
        buf = np.ascontiguousarray(dset[sl])
        if isinstance(psi, torch.Tensor):
            psi[color].copy_(torch.from_numpy(buf).to(psi.dtype))
        else:
            psi[color] = buf.astype(psi.dtype)
    orbitals.resetIterativeIndex()
    return numst


def read_attributes(h5file, name, color):
    """gids, lattice and origin of one function (HDFrestart::read_att and readGids,
    src/HDFrestart.cc:791-870)."""
    a = h5file[getDatasetName(name, color)].attrs
    return {"gids": [int(g) for g in a["List of gids"]] if "List of gids" in a else [],
            "lattice": [float(x) for x in a["Lattice parameters"]],
            "origin": [float(x) for x in a["Cell origin"]]}
