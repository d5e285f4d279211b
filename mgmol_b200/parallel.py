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
