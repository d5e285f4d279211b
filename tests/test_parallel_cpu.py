"""World-size-2 tests of the host-side multi-GPU logic on CPU (gloo): the
PEenv decomposition, the x-halo plan that mgb_halo_exchange_x implements, and
the partial-matrix all-reduce.  The transport here is gloo instead of NCCL;
the plan (who sends which planes to whom, in which order) is the code under
test and is shared with the GPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mgmol_b200.parallel import (cart_coords, cart_rank, geom, local_box, neighbours,
                                 x_halo_plan)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_geom_matches_reference_heuristic():
    # PEenv::geom on cubic meshes: z first, then y, then x (src/pb/PEenv.cc:335-598)
    assert geom(128, 128, 128, 1) == (1, 1, 1)
    assert geom(128, 128, 128, 2) == (1, 1, 2)
    assert geom(128, 128, 128, 4) == (1, 2, 2)
    assert geom(128, 128, 128, 8) == (2, 2, 2)
    assert geom(256, 128, 64, 8) == (4, 2, 1)
    # two factors of 2 per direction stay reserved: 8^3 cannot be split 4 ways in one direction
    assert geom(8, 8, 8, 8) == (2, 2, 2)
    assert geom(8, 8, 8, 16) is None
    # a direction not divisible by 4 is refused ("Poisson Solver Requires ...")
    assert geom(30, 32, 32, 2) is None


def test_geom_against_the_compiled_reference():
    """parallel.geom against pb::PEenv::geom itself (compiled unmodified, called through
    oracle/ref_shim_geom.cc): 15360 meshes x task counts from the committed table
    (tests/golden/make_golden_geom.py), and live where the compiled reference exists.  The
    port returns None where the reference refuses the mesh (0 tasks placed) or places
    fewer tasks than it was given (it would then shrink the communicator)."""
    import ctypes
    table = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                 "reference_geom.npz"))["table"]
    assert len(table) == 15360
    refused = shrunk = 0
    for nx, ny, nz, nt, placed, px, py, pz in table.tolist():
        want = (px, py, pz) if placed == nt else None
        refused += placed == 0
        shrunk += 0 < placed < nt
        assert geom(nx, ny, nz, nt) == want, (nx, ny, nz, nt, placed, (px, py, pz))
    assert refused > 0 and shrunk > 0          # both kinds of refusal are in the table
    from oracle.oracle import REF_SO, Ref
    if Ref.available():
        lib = ctypes.CDLL(REF_SO)
        out = (ctypes.c_int * 3)()
        for nx, ny, nz, nt, placed, px, py, pz in table[::37].tolist():
            assert lib.ref_geom(nx, ny, nz, nt, 1, out) == placed
            assert tuple(out) == (px, py, pz)


def test_cartesian_topology():
    nproc = (2, 2, 2)
    for r in range(8):
        assert cart_rank(cart_coords(r, nproc), nproc) == r
    # rank 3 = (0,1,1): x neighbours (1,1,1)=7 both ways on a periodic ring of 2
    assert neighbours(3, nproc) == [(7, 7), (1, 1), (2, 2)]
    assert local_box((8, 8, 8), nproc, (1, 0, 1)) == (slice(4, 8), slice(0, 4), slice(4, 8))
    with pytest.raises(ValueError):
        local_box((9, 8, 8), nproc, (0, 0, 0))


def _worker(rank, world, port, bc_x, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g, nfunc = 2, 3
        gdims = (8 * world, 4, 6)
        rng = np.random.default_rng(5)
        full = torch.from_numpy(rng.standard_normal((nfunc,) + gdims))
        nproc = (world, 1, 1)
        box = local_box(gdims, nproc, cart_coords(rank, nproc))
        mine = full[(slice(None),) + box].contiguous()
        # --- x halo, following the plan ---------------------------------------
        xhalo = torch.zeros((nfunc, 2 * g) + gdims[1:], dtype=torch.float64)
        sends, recvs = x_halo_plan(rank, nproc, g, bc_x)
        reqs, bufs = [], []
        for peer, planes in sends:
            reqs.append(dist.isend(mine[:, planes].contiguous(), peer))
        for peer, slot in recvs:
            b = torch.empty((nfunc, g) + gdims[1:], dtype=torch.float64)
            bufs.append((slot, b))
            reqs.append(dist.irecv(b, peer))
        for r in reqs:
            r.wait()
        for slot, b in bufs:
            xhalo[:, slot * g:(slot + 1) * g] = b
        x0 = box[0].start
        nxg = gdims[0]
        expect = torch.zeros_like(xhalo)
        for k in range(g):
            lo = x0 - g + k
            hi = x0 + (box[0].stop - box[0].start) + k
            if bc_x == 1 or lo >= 0:
                expect[:, k] = full[:, lo % nxg]
            if bc_x == 1 or hi < nxg:
                expect[:, g + k] = full[:, hi % nxg]
        ok_halo = bool(torch.equal(xhalo, expect))
        # --- partial Gram + all-reduce = global Gram ------------------------------
        a = mine.reshape(nfunc, -1)
        part = a @ a.t()
        dist.all_reduce(part)
        af = full.reshape(nfunc, -1)
        ok_gram = bool(torch.allclose(part, af @ af.t(), rtol=1e-13, atol=1e-13))
        out[rank] = (ok_halo, ok_gram)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bc_x", [1, 0])
def test_x_halo_plan_and_allreduce_world2(bc_x):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), bc_x, out), nprocs=world, join=True)
    assert dict(out) == {0: (True, True), 1: (True, True)}


def test_color_maps_follow_gid_addressed_packets():
    """color_maps = which slot of the neighbour the reference's receiver would
    take a face from (src/pb/GridFuncVector.cc:1225-1246,1374-1419)."""
    from mgmol_b200.parallel import color_maps
    mine = [[4, 7, -1, 2], [4, 9, 3, 2]]       # slabs 0 .. last, 4 colors
    west = [[1, 4, 7, 0], [7, 4, 5, -1]]       # its LAST slab faces me
    east = [[2, -1, 9, 6], [2, 8, 9, 6]]       # its FIRST slab faces me
    mw, me = color_maps(mine, west, east)
    # my first-slab gids 4, 7, -1, 2 in the west rank's last slab [7, 4, 5, -1]
    assert mw == [1, 0, -1, -1]
    # my last-slab gids 4, 9, 3, 2 in the east rank's first slab [2, -1, 9, 6]
    assert me == [-1, 2, -1, 0]
    assert color_maps(mine, None, east)[0] == [-1, -1, -1, -1]
    ident = [[0, 1, 2]]
    assert color_maps(ident, ident, ident) == ([0, 1, 2], [0, 1, 2])


def test_ghosted_color_maps():
    """Per-slab gid addressing of the packed exchange."""
    import numpy as np
    from mgmol_b200.parallel import ghosted_color_maps
    mine = np.array([[4, 7, -1], [4, 9, 3]])
    south = np.array([[7, 4, 2], [3, -1, 9]])
    west = np.array([[1, 1, 1], [7, 5, 4]])
    nb = [[west, None], [south, south], [None, None]]
    m = ghosted_color_maps(mine, nb)
    assert m.shape == (3, 2, 2, 3)
    # y, both sides: same slab of the sender
    assert m[1, 0].tolist() == [[1, 0, -1], [-1, 2, 0]]
    assert m[1, 1].tolist() == m[1, 0].tolist()
    # x low side: my slab 0 against the west rank's LAST slab; other slabs unused
    assert m[0, 0].tolist() == [[2, 0, -1], [-1, -1, -1]]
    assert (m[0, 1] == -1).all() and (m[2] == -1).all()


def _gid_worker(rank, world, port, out):
    """The integration recipe for localized orbitals on a split domain: each rank
    sends its overlapping_gids table to its x neighbours (MPI_Sendrecv in MGmol,
    gloo here), builds the slot translations, and the halo of every color then
    comes from the neighbour's slot holding the same global orbital."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mgmol_b200.parallel import color_maps
        subdivx, ncol, ngid = 2, 5, 6
        rs = np.random.RandomState(100 + rank)
        mine = np.stack([np.append(rs.permutation(ngid)[:ncol - 1], -1) for _ in range(subdivx)])
        west, east = (rank - 1) % world, (rank + 1) % world
        t = torch.from_numpy(mine.copy())
        tw, te = torch.empty_like(t), torch.empty_like(t)
        reqs = [dist.isend(t, west), dist.isend(t.clone(), east), dist.irecv(te, east),
                dist.irecv(tw, west)]
        for r in reqs:
            r.wait()
        mw, me = color_maps(mine, tw.numpy(), te.numpy())
        # the data of color c in the boundary slab is just its gid: fetch through the map
        ok = True
        for c in range(ncol):
            g_lo, g_hi = mine[0][c], mine[-1][c]
            got_w = tw.numpy()[-1][mw[c]] if mw[c] >= 0 else None
            got_e = te.numpy()[0][me[c]] if me[c] >= 0 else None
            ok = ok and (got_w == g_lo if got_w is not None else
                         (g_lo < 0 or g_lo not in tw.numpy()[-1]))
            ok = ok and (got_e == g_hi if got_e is not None else
                         (g_hi < 0 or g_hi not in te.numpy()[0]))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gid_tables_exchange_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gid_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
