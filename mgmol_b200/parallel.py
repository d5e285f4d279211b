"""One process per GPU: the NCCL communicator of the C ABI, bootstrapped
through torch.distributed (which plays the role MPI_Bcast plays inside MGmol:
it only carries the 128-byte NCCL id; no tensor data goes through it)."""
import ctypes

import torch
import torch.distributed as dist

from ._lib import check, lib
from .host import _dt, _p, _stream


class Communicator:
    """pb::PEenv's communicator for the hot path (src/pb/PEenv.h:34-203)."""

    def __init__(self, rank, nranks):
        self.rank, self.nranks = rank, nranks
        idbuf = (ctypes.c_ubyte * 128)()
        if rank == 0:
            check(lib().mgb_comm_unique_id(idbuf))
        t = torch.tensor(list(idbuf), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, 0)
        raw = bytes(t.cpu().tolist())
        self.handle = ctypes.c_void_p()
        check(lib().mgb_comm_create(ctypes.byref(self.handle), raw, rank, nranks))

    def allreduce(self, t):
        """MGmol_MPI::allreduce(double*, n, MPI_SUM) (src/ExtendedGridOrbitals.cc:1746)."""
        assert t.dtype == torch.float64 and t.is_contiguous()
        check(lib().mgb_allreduce_sum_f64(self.handle, _p(t), t.numel(), _stream()))
        return t

    def barrier(self):
        """Stream-ordered barrier over the ranks."""
        check(lib().mgb_comm_barrier(self.handle, _stream()))

    def check(self):
        """Device sync + did any neighbour barrier time out?"""
        check(lib().mgb_comm_check(self.handle))

    def register(self, t):
        """Collective: publish this rank's array so that the neighbours'
        kernels can read its boundary planes in place over NVLink (CUDA IPC).
        Raises MgbError(-2) when the memory cannot be exported."""
        check(lib().mgb_peer_register(self.handle, _p(t), _stream()))

    def unregister(self, t):
        check(lib().mgb_peer_unregister(self.handle, _p(t)))

    def set_color_maps(self, map_west, map_east):
        """LocGridOrbitals on an x-split domain: color slot of my colors' orbitals
        on the west / east rank (see color_maps); None, None = identity."""
        if map_west is None:
            check(lib().mgb_peer_set_color_maps(self.handle, None, None, 0))
            return
        import numpy as np
        mw = np.ascontiguousarray(map_west, dtype=np.int32)
        me = np.ascontiguousarray(map_east, dtype=np.int32)
        check(lib().mgb_peer_set_color_maps(
            self.handle, mw.ctypes.data_as(ctypes.c_void_p),
            me.ctypes.data_as(ctypes.c_void_p), len(mw)))

    def set_ghosted_color_maps(self, maps):
        """gid-addressed packed exchange: maps[dir][side][iloc][color] (see
        ghosted_color_maps); None = slot-for-slot."""
        if maps is None:
            check(lib().mgb_halo_set_color_maps(self.handle, 0, 0, None))
            return
        import numpy as np
        m = np.ascontiguousarray(maps, dtype=np.int32)
        assert m.ndim == 4 and m.shape[:2] == (3, 2)
        check(lib().mgb_halo_set_color_maps(self.handle, m.shape[2], m.shape[3],
                                            m.ctypes.data_as(ctypes.c_void_p)))

    def halo_exchange_x(self, grid, g, noghost, xhalo):
        nfunc = noghost.shape[0]
        check(lib().mgb_halo_exchange_x(self.handle, _dt(noghost), grid.ref(), g,
                                        _p(noghost), grid.size(), _p(xhalo), nfunc,
                                        _stream()))

    def trade_boundaries(self, gfv):
        """GridFuncVector::trade_boundaries on a px x py x pz decomposition."""
        if gfv.updated_boundaries_:
            return
        check(lib().mgb_halo_exchange_ghosted(self.handle, _dt(gfv.data), gfv.grid_.ref(),
                                              _p(gfv.data), gfv.nfunc_, _stream()))
        gfv.updated_boundaries_ = True

    def close(self):
        if self.handle:
            lib().mgb_comm_destroy(self.handle)
            self.handle = None


# ---------------------------------------------------------------------------
# Host-side decomposition logic (no device needed): pb::PEenv
# ---------------------------------------------------------------------------
_PRIMES = (2, 3, 5, 7, 11, 13, 17, 19, 23)


def geom(nx, ny, nz, ntasks, bias=1):
    """PEenv::geom (src/pb/PEenv.cc:335-598): split `ntasks` ranks over x, y, z
    by handing the largest prime factors of the rank count to the currently
    largest grid direction; two factors of 2 per direction stay reserved for
    the Poisson multigrid.  Returns (px, py, pz) or None when the reference
    refuses the mesh (a direction not divisible by 4, or ranks left over)."""
    n = [nx, ny, nz, ntasks]
    fac = [[0] * len(_PRIMES) for _ in range(4)]
    for i in range(4):
        m = n[i]
        for p, prime in enumerate(_PRIMES):
            while m % prime == 0 and fac[i][p] < 20:
                m //= prime
                fac[i][p] += 1
    for i in range(3):
        if fac[i][0] > 1:
            fac[i][0] -= 2
        else:
            return None
    n = [nx, ny, nz, ntasks]
    ndir = [1, 1, 1]
    div = True
    while div:
        div = False
        if n[2] >= n[1] and bias * n[1] >= n[0]:
            order = (2, 1, 0)
        elif n[1] >= n[2] and bias * n[2] >= n[0]:
            order = (1, 2, 0)
        elif bias * n[2] >= n[0] and n[0] >= bias * n[1]:
            order = (2, 0, 1)
        elif bias * n[1] >= n[0] and n[0] >= bias * n[2]:
            order = (1, 0, 2)
        elif n[0] >= bias * n[2] and n[2] >= n[1]:
            order = (0, 2, 1)
        elif n[0] >= bias * n[1] and n[1] >= n[2]:
            order = (0, 1, 2)
        else:
            order = (1, 2, 0)
        for d in order:
            if div:
                break
            for k in range(len(_PRIMES) - 1, -1, -1):
                if fac[d][k] > 0 and fac[3][k] > 0:
                    ndir[d] *= _PRIMES[k]
                    fac[d][k] -= 1
                    n[d] //= _PRIMES[k]
                    fac[3][k] -= 1
                    n[3] //= _PRIMES[k]
                    div = True
                    break
    if ndir[0] * ndir[1] * ndir[2] != ntasks:
        return None
    return tuple(ndir)


def geom_b200(nx, ny, nz, ntasks):
    """The decomposition used on an NVLink box.  PEenv::geom (above) hands the prime factors
    of the rank count to the currently largest direction, z first on ties: 1x1x2, 1x2x2,
    2x2x2 for a cubic grid.  Here the same factors go to x and y only -- the largest local
    extent first, x on ties -- and to z only when neither can take one (local extents stay
    multiples of 4 for the two multigrid levels): 2x1x1, 2x2x1, 4x2x1.  Reason, measured on
    8 x B200 (profiles/r02_decomposition.md): x planes and y rows of a neighbour are
    contiguous in its memory and the fused kernels read them in place at no measurable
    cost, while the z halo is 16 bytes out of every row -- 134 M DRAM page activations per
    step on the 256^3 x 4096 block, 7 ms whether the columns are pushed into dense buffers
    (comm_zhalo) or fetched in place.  Every px x py x pz decomposition, 2x2x2 included, is
    served by the same kernels (bench.py --decomp)."""
    local = [nx, ny, nz]
    out = [1, 1, 1]
    m = ntasks
    primes = []
    for p in _PRIMES:
        while m % p == 0:
            primes.append(p)
            m //= p
    if m != 1:
        return None
    for p in sorted(primes, reverse=True):
        cands = [d for d in (0, 1) if local[d] % (4 * p) == 0]
        if not cands and local[2] % (4 * p) == 0:
            cands = [2]
        if not cands:
            return None
        d = max(cands, key=lambda q: (local[q], -q))
        out[d] *= p
        local[d] //= p
    return tuple(out)


def cart_coords(rank, nproc):
    """MPI_Cart_coords on the row-major communicator of MPI_Cart_create
    (src/pb/PEenv.cc:89): rank = (cx * py + cy) * pz + cz."""
    px, py, pz = nproc
    return (rank // (py * pz), (rank // pz) % py, rank % pz)


def cart_rank(coord, nproc):
    px, py, pz = nproc
    return ((coord[0] % px) * py + (coord[1] % py)) * pz + (coord[2] % pz)


def neighbours(rank, nproc):
    """PEenv neighbours (src/pb/PEenv.cc:300-312): (low, high) rank per
    direction on the always-periodic Cartesian topology."""
    c = cart_coords(rank, nproc)
    out = []
    for d in range(3):
        lo = list(c)
        hi = list(c)
        lo[d] -= 1
        hi[d] += 1
        out.append((cart_rank(lo, nproc), cart_rank(hi, nproc)))
    return out


def local_box(gdims, nproc, coord):
    """Grid::dim / the slice of the global grid owned by `coord`
    (src/pb/Grid.cc:49-54): equal blocks, global dims must divide."""
    for n, p in zip(gdims, nproc):
        if n % p:
            raise ValueError("global dims must divide by ranks (Grid.cc:52-54)")
    dims = tuple(n // p for n, p in zip(gdims, nproc))
    return tuple(slice(c * d, (c + 1) * d) for c, d in zip(coord, dims))


def color_maps(my_gids, west_gids, east_gids):
    """The gid-addressed halo of GridFuncVector (src/pb/GridFuncVector.cc:
    1225-1246,1374-1419) as index translations for in-place peer reads.
    *_gids: (subdivx, ncolors) tables overlapping_gids_[iloc][color] of this
    rank and of its west / east neighbours (None where the domain ends).

    The west rank tags the planes it sends east with the gid of its LAST slab;
    I store them in my color whose FIRST slab holds that gid.  Hence
    map_west[c] = c' with west_gids[-1][c'] == my_gids[0][c], or -1; and
    map_east[c] = c' with east_gids[0][c'] == my_gids[-1][c], or -1."""
    def one(mine, theirs):
        out = []
        lookup = {} if theirs is None else {int(g): i for i, g in enumerate(theirs) if g >= 0}
        for g in mine:
            out.append(lookup.get(int(g), -1) if g >= 0 else -1)
        return out
    return (one(my_gids[0], None if west_gids is None else west_gids[-1]),
            one(my_gids[-1], None if east_gids is None else east_gids[0]))


def ghosted_color_maps(my_gids, neighbour_gids):
    """maps[dir][side][iloc][color] for Communicator.set_ghosted_color_maps.
    neighbour_gids[dir][side]: the (subdivx, ncolors) gid table of the rank that
    fills my low (side 0) / high (side 1) ghosts in direction dir, or None.
    y / z: slab by slab, the sender's color holding the same gid in the same slab
    (src/pb/GridFuncVector.cc:461-513, 958-998); x: the sender's boundary slab
    against my boundary slab (:1374-1419)."""
    import numpy as np
    my = np.asarray(my_gids)
    subdivx, ncolors = my.shape
    out = -np.ones((3, 2, subdivx, ncolors), np.int32)
    for d in range(3):
        for side in range(2):
            nb = neighbour_gids[d][side]
            if nb is None:
                continue
            nb = np.asarray(nb)
            for iloc in range(subdivx):
                if d == 0:
                    if iloc != (0 if side == 0 else subdivx - 1):
                        continue
                    theirs = nb[-1] if side == 0 else nb[0]
                else:
                    theirs = nb[iloc]
                lookup = {int(g): i for i, g in enumerate(theirs) if g >= 0}
                for c in range(ncolors):
                    g = int(my[iloc, c])
                    out[d, side, iloc, c] = lookup.get(g, -1) if g >= 0 else -1
    return out


def x_halo_plan(rank, nproc, g, bc_x=1):
    """What mgb_halo_exchange_x moves for one rank of an x-split domain:
    a list of (peer, send_planes, recv_slot) with send_planes a slice of the
    local x planes and recv_slot 0 (planes below the box) or 1 (above).
    Sends are ordered west, east and receives east, west, so that with two
    ranks (both neighbours the same peer) the k-th send pairs with the peer's
    k-th receive."""
    coord = cart_coords(rank, nproc)
    (west, east) = neighbours(rank, nproc)[0]
    have_w = bc_x == 1 or coord[0] > 0
    have_e = bc_x == 1 or coord[0] < nproc[0] - 1
    sends, recvs = [], []
    if have_w:
        sends.append((west, slice(0, g)))
    if have_e:
        sends.append((east, slice(-g, None)))
    if have_e:
        recvs.append((east, 1))
    if have_w:
        recvs.append((west, 0))
    return sends, recvs
