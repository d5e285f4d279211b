"""Host-side mirror of the reference's operator API for the hot path, over the C
ABI.  Same class and method names as MGmol (file:line in the docstrings) so
the parity tests read like the reference's own unit tests.  PyTorch appears
only as the owner of device memory and streams; every number is produced by a
kernel of libmgmol_b200.so.

The production host for a C++ caller is include/mgmol_b200.hpp (same mirror in
C++); this module is what tests/ and bench.py drive.
"""
import ctypes

import torch

from . import _lib
from ._lib import MGB_F32, MGB_F64, MgbError, MgbGrid, check, lib  # noqa: F401


def _dt(t):
    if t.dtype == torch.float64:
        return MGB_F64
    if t.dtype == torch.float32:
        return MGB_F32
    raise TypeError("ORBDTYPE must be float32 or float64")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def ghosts_for(lap_type):
    """GridFactory (src/GridFactory.h:23-51): ghost width per operator."""
    return {0: 1, 10: 1, 1: 1, 2: 2, 3: 3, 4: 4}[lap_type]


class Grid:
    """pb::Grid + pb::PEenv for one rank (src/pb/Grid.h:24-121)."""

    def __init__(self, gdim, lattice, ghosts=1, bc=(1, 1, 1), nproc=(1, 1, 1),
                 coord=(0, 0, 0), level=0):
        self.gdim_ = tuple(int(n) for n in gdim)
        self.ll_ = tuple(float(l) for l in lattice)
        self.nproc = tuple(nproc)
        self.coord = tuple(coord)
        for n, p in zip(self.gdim_, self.nproc):
            if n % p:
                raise ValueError("global dims must divide by ranks (Grid.cc:52-54)")
        self.dim_ = tuple(n // p for n, p in zip(self.gdim_, self.nproc))
        self.ghost_pt_ = int(ghosts)
        self.bc = tuple(bc)
        self.level_ = level
        self.c = MgbGrid()
        for d in range(3):
            self.c.dim[d] = self.dim_[d]
            self.c.gdim[d] = self.gdim_[d]
            self.c.h[d] = self.ll_[d] / self.gdim_[d]
            self.c.bc[d] = self.bc[d]
            self.c.nproc[d] = self.nproc[d]
            self.c.coord[d] = self.coord[d]
        self.c.ghosts = self.ghost_pt_

    def dim(self, i):
        return self.dim_[i]

    def gdim(self, i):
        return self.gdim_[i]

    def ghost_pt(self):
        return self.ghost_pt_

    def hgrid(self, i):
        return self.ll_[i] / self.gdim_[i]

    def vel(self):
        return self.hgrid(0) * self.hgrid(1) * self.hgrid(2)

    def size(self):
        return self.dim_[0] * self.dim_[1] * self.dim_[2]

    def sizeg(self):
        g = self.ghost_pt_
        return (self.dim_[0] + 2 * g) * (self.dim_[1] + 2 * g) * (self.dim_[2] + 2 * g)

    def shape(self):
        return self.dim_

    def shapeg(self):
        g = self.ghost_pt_
        return tuple(d + 2 * g for d in self.dim_)

    def coarse_grid(self):
        """Grid::coarse_grid (src/pb/Grid.cc:214-231): same ghosts, half dims."""
        return Grid(tuple(n // 2 for n in self.gdim_), self.ll_, self.ghost_pt_,
                    self.bc, self.nproc, self.coord, self.level_ - 1)

    def with_ghosts(self, g):
        return Grid(self.gdim_, self.ll_, g, self.bc, self.nproc, self.coord,
                    self.level_)

    def ref(self):
        return ctypes.byref(self.c)


class GridFuncVector:
    """pb::GridFuncVector<T, MemorySpace::Device> (src/pb/GridFuncVector.h):
    nfunc ghosted functions in one device allocation."""

    def __init__(self, grid, nfunc, dtype=torch.float64, data=None):
        self.grid_ = grid
        self.nfunc_ = nfunc
        shape = (nfunc,) + grid.shapeg()
        self.data = (torch.zeros(shape, dtype=dtype, device="cuda")
                     if data is None else data)
        assert self.data.shape == shape and self.data.is_contiguous()
        self.updated_boundaries_ = False

    def grid(self):
        return self.grid_

    def size(self):
        return self.nfunc_

    def resetData(self):
        self.data.zero_()
        self.updated_boundaries_ = True

    def set_updated_boundaries(self, flag):
        self.updated_boundaries_ = bool(flag)

    def assign(self, noghost):
        """BlockVector::setDataWithGhosts (src/BlockVector.cc:489-517)."""
        assert noghost.is_contiguous() and noghost.shape[0] == self.nfunc_
        check(lib().mgb_gfv_set_with_ghosts(
            _dt(noghost), _dt(self.data), self.grid_.ref(), _p(noghost),
            self.grid_.size(), _p(self.data), self.nfunc_, _stream()))
        self.updated_boundaries_ = False

    def getValues(self, out):
        """BlockVector::assign(GridFuncVector) (src/BlockVector.cc:303-311)."""
        check(lib().mgb_gfv_get_values(
            _dt(self.data), _dt(out), self.grid_.ref(), _p(self.data), _p(out),
            self.grid_.size(), self.nfunc_, _stream()))
        return out

    def values(self):
        """getValues into a new no-ghost block."""
        return self.getValues(torch.empty((self.nfunc_,) + tuple(self.grid_.shape()),
                                          dtype=self.data.dtype, device="cuda"))

    def trade_boundaries(self):
        """src/pb/GridFuncVector.cc:1544-1622."""
        if self.updated_boundaries_:
            return
        check(lib().mgb_gfv_trade_boundaries(
            _dt(self.data), self.grid_.ref(), _p(self.data), self.nfunc_, _stream()))
        self.updated_boundaries_ = True

    def _fd(self, kind, rhs, rhs_ghosts=0):
        self.trade_boundaries()
        check(lib().mgb_fd_apply(kind, _dt(self.data), self.grid_.ref(),
                                 _p(self.data), _p(rhs.data), self.nfunc_,
                                 rhs_ghosts, _stream()))
        rhs.set_updated_boundaries(False)

    def applyLap(self, lap_type, rhs):
        """src/pb/GridFuncVector.cc:2370-2397."""
        kinds = {0: _lib.FD_DEL2_4TH_MEHR, 1: _lib.FD_DEL2_2ND, 2: _lib.FD_DEL2_4TH,
                 3: _lib.FD_DEL2_6TH, 4: _lib.FD_DEL2_8TH}
        if lap_type not in kinds:
            raise ValueError("LapFactory::createLap() --- option invalid: %d" % lap_type)
        self._fd(kinds[lap_type], rhs)

    def applyRHS(self, lap_type, rhs):
        """src/pb/GridFuncVector.cc:2400-2413."""
        if lap_type == 0:
            self._fd(_lib.FD_RHS_4TH_MEHR1, rhs, self.grid_.ghost_pt())
        else:
            self.data.copy_(rhs.data)

    def pointwiseProduct(self, A, Vghost):
        """src/pb/GridFuncVector.cc:90-136; Vghost: ghosted double field."""
        check(lib().mgb_gfv_pointwise_product(
            _dt(self.data), self.grid_.ref(), _p(A.data), _p(Vghost), _p(self.data),
            self.nfunc_, _stream()))
        self.updated_boundaries_ = A.updated_boundaries_

    def axpy(self, alpha, other):
        """src/pb/GridFuncVector.cc:1660-1668."""
        check(lib().mgb_axpy(_dt(self.data), self.data.numel(), float(alpha),
                             _p(other.data), _p(self.data), _stream()))
        self.updated_boundaries_ = other.updated_boundaries_ and self.updated_boundaries_

    def __isub__(self, other):
        self.axpy(-1.0, other)
        return self

    def scal(self, alpha):
        """GridFunc::operator*=(double) = MPscal over every stored value
        (src/pb/GridFunc.cc:508-514)."""
        check(lib().mgb_scal(_dt(self.data), self.data.numel(), float(alpha), _p(self.data),
                             _stream()))

    def jacobi(self, lap_type, B, w, jacobi_factor):
        """src/pb/GridFuncVector.cc:2416-2425."""
        self.trade_boundaries()
        check(lib().mgb_gfv_jacobi(lap_type, self.grid_.ref(), _p(self.data),
                                   _p(B.data), _p(w.data), self.nfunc_,
                                   float(jacobi_factor), _stream()))
        self.updated_boundaries_ = False
        w.set_updated_boundaries(False)

    def restrict3D(self, ucoarse):
        """src/pb/GridFuncVector.cc:1624-1631."""
        self.trade_boundaries()
        check(lib().mgb_gfv_restrict3D(_dt(self.data), self.grid_.ref(),
                                       _p(self.data), _p(ucoarse.data),
                                       self.nfunc_, _stream()))

    def extend3D(self, ucoarse):
        """src/pb/GridFuncVector.cc:1633-1641."""
        ucoarse.trade_boundaries()
        check(lib().mgb_gfv_extend3D(_dt(self.data), self.grid_.ref(),
                                     _p(ucoarse.data), _p(self.data), self.nfunc_,
                                     _stream()))
        self.updated_boundaries_ = False


    # -- pb::GridFunc operations of the Poisson multigrid (one function) -------
    def copy_from(self, other):
        """GridFunc copy constructor / operator=: all values, ghosts included."""
        assert other.data.shape == self.data.shape and other.data.dtype == self.data.dtype
        check(lib().mgb_copy_dev(_p(self.data), _p(other.data),
                                 self.data.numel() * self.data.element_size(), _stream()))
        self.updated_boundaries_ = other.updated_boundaries_

    def _interior(self, drop_bc_layers):
        g = self.grid_.ghost_pt()
        nx, ny, nz = self.grid_.shape()
        lo = [g + (1 if (drop_bc_layers and self.grid_.bc[d] != 1
                         and self.grid_.coord[d] == 0) else 0) for d in range(3)]
        return self.data[:, lo[0]:g + nx, lo[1]:g + ny, lo[2]:g + nz].contiguous()

    def gdot(self, other, comm=None):
        """GridFunc::gdot (src/pb/GridFunc.cc:2745-2798): interior points, minus
        the first layer in every non-periodic direction; double sum."""
        a, b = self._interior(True), other._interior(True)
        out = torch.empty(1, dtype=torch.float64, device="cuda")
        check(lib().mgb_dot(_dt(a), a.numel(), _p(a), _p(b), _p(out), _stream()))
        if comm is not None:
            comm.allreduce(out)
        return float(out.item())

    def norm2(self, comm=None):
        """GridFunc::norm2 (src/pb/GridFunc.cc:2856-2861)."""
        return (self.gdot(self, comm) * self.grid_.vel()) ** 0.5

    def get_average(self):
        """GridFunc::get_average (src/pb/GridFunc.cc:2888-2927), one rank."""
        a = self._interior(False)
        ones = torch.ones_like(a)
        out = torch.empty(1, dtype=torch.float64, device="cuda")
        check(lib().mgb_dot(_dt(a), a.numel(), _p(a), _p(ones), _p(out), _stream()))
        return float(out.item()) / float(self.grid_.size())

    def add_scalar(self, alpha):
        """GridFunc::operator+=(T) / operator-=(T): every stored value."""
        ones = torch.ones_like(self.data)
        check(lib().mgb_axpy(_dt(self.data), self.data.numel(), float(alpha), _p(ones),
                             _p(self.data), _stream()))


class Lap:
    """pb::Lap<T> family (src/pb/Lap.h:19-54) as created by LapFactory."""

    def __init__(self, grid, lap_type):
        self.grid_ = grid
        self.type_ = lap_type
        h = (ctypes.c_double * 3)(*(grid.hgrid(i) for i in range(3)))
        out = (ctypes.c_double * 3)()
        check(lib().mgb_lap_constants(lap_type, h, out))
        self.diagEl_, self.invDiagEl_, self.jacobiFactor_ = out[0], out[1], out[2]

    def diagEl(self):
        return self.diagEl_

    def invDiagEl(self):
        return self.invDiagEl_

    def jacobiFactor(self):
        return self.jacobiFactor_

    def minNumberGhosts(self):
        return ghosts_for(self.type_)

    def apply(self, A, B):
        """Lap::apply on a block (Laph4::apply(grid, A, B, nfunc),
        src/pb/Laph4.h:99-102, and the GridFuncVector analogue)."""
        A.applyLap(0 if self.type_ == 10 else self.type_, B)

    def applyWithPot(self, phi, vtot, hphi, xhalo_phi=None, xhalo_v=None):
        """Lap<T>::applyWithPot (src/pb/Lap.h:35; Laph4.h:95-98) generalised
        to the whole block = Hamiltonian::applyLocal body."""
        nfunc = phi.shape[0]
        check(lib().mgb_hpsi(
            self.type_, _dt(phi), self.grid_.ref(), _p(phi), self.grid_.size(),
            _p(vtot), _p(hphi), self.grid_.size(), nfunc,
            _p(xhalo_phi) if xhalo_phi is not None else None,
            _p(xhalo_v) if xhalo_v is not None else None, _stream()))
        return hphi


    def rhs(self, phi, bphi, xhalo_phi=None):
        """Lap<T>::rhs (src/pb/Lap.h:32) for every orbital of a no-ghost block:
        B phi (Mehrstellen) or a copy (B = 1)."""
        check(lib().mgb_apply_b(
            self.type_, _dt(phi), self.grid_.ref(), _p(phi), self.grid_.size(), _p(bphi),
            self.grid_.size(), phi.shape[0],
            _p(xhalo_phi) if xhalo_phi is not None else None, _stream()))
        return bphi

    def applyWithPotPeer(self, comm, phi, vtot, hphi, xhalo_v=None, vghost=None):
        """applyWithPot on a decomposed domain, the neighbours' boundary layers of
        phi read directly from their (registered) blocks over NVLink.  x slabs take
        the potential's 2g packed x planes (xhalo_v); any px x py x pz decomposition
        takes a ghosted copy of V whose boundaries were traded (vghost)."""
        if vghost is not None:
            check(lib().mgb_hpsi_peer3d(
                comm.handle, self.type_, _dt(phi), self.grid_.ref(), _p(phi), self.grid_.size(),
                _p(vtot), _p(vghost), _p(hphi), self.grid_.size(), phi.shape[0], _stream()))
            return hphi
        check(lib().mgb_hpsi_peer(
            comm.handle, self.type_, _dt(phi), self.grid_.ref(), _p(phi), self.grid_.size(),
            _p(vtot), _p(hphi), self.grid_.size(), phi.shape[0], _p(xhalo_v), _stream()))
        return hphi

    def applyWithPotHost(self, phi_host, vtot_host, hphi_host, chunk=0):
        """The same operator on HOST blocks (MGmol's MemorySpace::Host build):
        pinned CPU tensors in, pinned CPU tensor out, copies and kernel
        pipelined inside the library (mgb_hpsi_host)."""
        assert not phi_host.is_cuda and not hphi_host.is_cuda and not vtot_host.is_cuda
        assert phi_host.is_contiguous() and hphi_host.is_contiguous()
        nfunc = phi_host.shape[0]
        check(lib().mgb_hpsi_host(
            self.type_, _dt(phi_host), self.grid_.ref(), _p(phi_host), self.grid_.size(),
            _p(vtot_host), _p(hphi_host), self.grid_.size(), nfunc, int(chunk)))
        return hphi_host


    def applyWithPotHostPeer(self, comm, phi_host, vtot_host, hphi_host, chunk=0):
        """applyWithPotHost for one rank of an x-split domain (collective): every
        rank streams its sub-box through its own GPU; halos are read in place
        from the neighbours' input rings (mgb_hpsi_host_peer)."""
        assert not phi_host.is_cuda and not hphi_host.is_cuda and not vtot_host.is_cuda
        assert phi_host.is_contiguous() and hphi_host.is_contiguous()
        check(lib().mgb_hpsi_host_peer(
            comm.handle, self.type_, _dt(phi_host), self.grid_.ref(), _p(phi_host),
            self.grid_.size(), _p(vtot_host), _p(hphi_host), self.grid_.size(),
            phi_host.shape[0], int(chunk)))
        return hphi_host


class LapFactory:
    """src/LapFactory.h:26-56."""

    @staticmethod
    def createLap(grid, lap_type):
        if lap_type not in (0, 1, 2, 3, 4, 10):
            raise ValueError("LapFactory::createLap() --- option invalid:%d" % lap_type)
        return Lap(grid, lap_type)


class Potentials:
    """The slice of Potentials the path reads: vtot (POTDTYPE double, no
    ghosts, src/Potentials.h:141) and its iterative index."""

    def __init__(self, vtot):
        assert vtot.dtype == torch.float64 and vtot.is_contiguous()
        self.vtot_ = vtot
        self.itindex_ = 0

    def vtot(self):
        return self.vtot_

    def getIterativeIndex(self):
        return self.itindex_

    def update(self, vtot):
        self.vtot_.copy_(vtot)
        self.itindex_ += 1

    def max(self):
        return float(self.vtot_.max())


class Orbitals:
    """ExtendedGridOrbitals (src/ExtendedGridOrbitals.h:43-404) reduced to the
    hot path: psi as a no-ghost column-major npt x numst block (here a
    contiguous (numst, nx, ny, nz) tensor, lda = npt) plus the iterative index
    that keys Hamiltonian's cache (src/Orbitals.h:43-70)."""

    def __init__(self, grid, numst, dtype=torch.float64, psi=None):
        self.grid_ = grid
        self.numst_ = numst
        shape = (numst,) + grid.shape()
        self.psi_ = (torch.zeros(shape, dtype=dtype, device="cuda")
                     if psi is None else psi)
        assert self.psi_.shape == shape and self.psi_.is_contiguous()
        self.iterative_index_ = 0

    def chromatic_number(self):
        return self.numst_

    def getIterativeIndex(self):
        return self.iterative_index_

    def incrementIterativeIndex(self):
        self.iterative_index_ += 1

    def resetIterativeIndex(self):
        """src/Orbitals.h:43 (after read_func_hdf5)."""
        self.iterative_index_ = 0

    def psi(self):
        return self.psi_

    def getNumpt(self):
        return self.grid_.size()

    # -- BLAS-1 (src/ExtendedGridOrbitals.cc:202-212, BlockVector.h:112-126) --
    def axpy(self, alpha, other):
        check(lib().mgb_axpy(_dt(self.psi_), self.psi_.numel(), float(alpha),
                             _p(other.psi_), _p(self.psi_), _stream()))
        self.incrementIterativeIndex()

    def scal(self, alpha):
        check(lib().mgb_scal(_dt(self.psi_), self.psi_.numel(), float(alpha),
                             _p(self.psi_), _stream()))
        self.incrementIterativeIndex()

    def assign(self, other):
        """BlockVector::assign / Orbitals::assign: copy the block."""
        check(lib().mgb_copy_dev(_p(self.psi_), _p(other.psi_),
                                 self.psi_.numel() * self.psi_.element_size(), _stream()))
        self.incrementIterativeIndex()

    def __isub__(self, other):
        """operator-= (src/BlockVector.cc:289-299)."""
        check(lib().mgb_axpy(_dt(self.psi_), self.psi_.numel(), -1.0, _p(other.psi_),
                             _p(self.psi_), _stream()))
        self.incrementIterativeIndex()
        return self

    def clone(self):
        return type(self).__new__(type(self))._init_like(self)

    def _init_like(self, other):
        self.__dict__.update(other.__dict__)
        self.psi_ = other.psi_.clone()
        return self

    def computeDiagonalElementsDotProduct(self, other, comm=None):
        """ss[i] = vel * <phi_i, psi_i> for every orbital
        (src/ExtendedGridOrbitals.cc:1085-1106), one launch."""
        ss = torch.empty(self.numst_, dtype=torch.float64, device="cuda")
        check(lib().mgb_dot_cols(_dt(self.psi_), self.grid_.size(), self.numst_,
                                 self.grid_.vel(), _p(self.psi_), self.grid_.size(),
                                 _p(other.psi_), self.grid_.size(), _p(ss), _stream()))
        if comm is not None:
            comm.allreduce(ss)
        return ss

    def dotProduct(self, other, inv_s_diag=None, comm=None):
        """dotProductDiagonal (src/ExtendedGridOrbitals.cc:1205-1213): the trace
        of diag(S^-1) * diag(Phi^T Psi); inv_s_diag defaults to ones."""
        ss = self.computeDiagonalElementsDotProduct(other, comm)
        if inv_s_diag is not None:
            ss = ss * inv_s_diag
        return float(ss.sum())

    # -- contractions ---------------------------------------------------------
    def computeLocalProduct(self, other, comm=None):
        """vel * Phi^T A  (src/ExtendedGridOrbitals.cc:1049-1083); with a
        communicator the partial matrices are summed over ranks
        (addDotWithNcol2Matrix, :1704-1752)."""
        a = other.psi_ if isinstance(other, Orbitals) else other
        n = self.numst_
        ss = torch.empty((n, n), dtype=torch.float64, device="cuda")
        # column-major C(i,j) = <phi_i, a_j>; the row-major view is C^T
        check(lib().mgb_gemm_tn(_dt(self.psi_), n, n, self.grid_.size(),
                                self.grid_.vel(), _p(self.psi_), self.grid_.size(),
                                _p(a), self.grid_.size(), 0.0, _p(ss), n, _stream()))
        if comm is not None:
            comm.allreduce(ss)
        return ss.t()

    def addDotWithNcol2Matrix(self, other, mat, comm=None):
        """mat += vel * Phi^T A summed over the ranks
        (src/ExtendedGridOrbitals.cc:1704-1752); mat indexed [i, j] = <phi_i, a_j>."""
        mat += self.computeLocalProduct(other, comm)
        return mat

    def computeMatB(self, lapOper, orbitals=None, comm=None, work=None, xhalo_phi=None,
                    bcolor=32):
        """matB[i, j] = vel * <orbitals_i | B | this_j> (src/ExtendedGridOrbitals.cc:901-967),
        B = Lap::rhs (the Mehrstellen right-hand-side stencil, the identity otherwise), in
        blocks of `bcolor` columns through a work block like the reference's (32 there).
        work: an Orbitals whose storage may be used (>= bcolor orbitals)."""
        o = self if orbitals is None else orbitals
        n, m = self.numst_, o.numst_
        npt = self.grid_.size()
        if work is not None:
            bcolor = work.psi_.shape[0]  # the whole work block: one B pass, one contraction
            w = work.psi_
        else:
            bcolor = min(bcolor, n)
            w = torch.empty((bcolor,) + self.grid_.shape(), dtype=self.psi_.dtype, device="cuda")
        ss = torch.empty((n, m), dtype=torch.float64, device="cuda")  # column-major m x n
        for j0 in range(0, n, bcolor):
            nf = min(bcolor, n - j0)
            xh = xhalo_phi[j0:j0 + nf] if xhalo_phi is not None else None
            lapOper.rhs(self.psi_[j0:j0 + nf], w[:nf], xh)
            check(lib().mgb_gemm_tn(_dt(self.psi_), m, nf, npt, self.grid_.vel(), _p(o.psi_), npt,
                                    _p(w), npt, 0.0, _p(ss[j0:]), m, _stream()))
        if comm is not None:
            comm.allreduce(ss)
        return ss.t()

    def computeGram(self, comm=None):
        """S = vel * Phi^T Phi (src/ExtendedGridOrbitals.cc:985-1010,1138-1162)."""
        n = self.numst_
        ss = torch.empty((n, n), dtype=torch.float64, device="cuda")
        check(lib().mgb_syrk_t(_dt(self.psi_), n, self.grid_.size(), self.grid_.vel(),
                               _p(self.psi_), self.grid_.size(), _p(ss), n, _stream()))
        if comm is not None:
            comm.allreduce(ss)
        return ss

    def multiplyByMatrix(self, matrix, product=None):
        """Phi * M (src/ExtendedGridOrbitals.cc:448-498).  matrix[l, j] is a
        (numst, n) double tensor (row-major = the transpose of the reference's
        column-major storage is handled here)."""
        n = matrix.shape[1]
        inplace = product is None
        out = torch.empty((n,) + self.grid_.shape(), dtype=self.psi_.dtype,
                          device="cuda") if inplace else product.psi_
        mcol = matrix.t().contiguous()  # column-major k x n
        check(lib().mgb_gemm_nn(_dt(self.psi_), self.grid_.size(), n, self.numst_,
                                1.0, _p(self.psi_), self.grid_.size(), _p(mcol),
                                self.numst_, 0.0, _p(out), self.grid_.size(),
                                _stream()))
        if inplace:
            self.psi_.copy_(out)
            self.incrementIterativeIndex()
        return out

    def orthonormalizeLoewdin(self, gram=None, comm=None):
        """Phi <- Phi S^-1/2 (src/ExtendedGridOrbitals.cc:1304-1358).  The
        transform is GramMatrix::computeLoewdinTransform (src/GramMatrix.cc:
        267-285): S = V diag(l) V^T by dsyev, P = V diag(1/sqrt(l)) V^T.  The
        N x N eigenproblem belongs to the reference's dense-matrix layer
        (ProjectedMatrices, not part of the grid path), so it is a library call
        (cuSOLVER through torch.linalg.eigh) on the device: S never leaves HBM.
        The two grid-sized steps, the Gram matrix and Phi P, are this library's
        kernels.  Returns P (the reference's matrixTransform)."""
        S = self.computeGram(comm) if gram is None else gram
        lam, V = torch.linalg.eigh(S)
        if not bool(lam[0] > 0):
            raise ArithmeticError("orthonormalizeLoewdin: Gram matrix is not positive definite")
        P = (V * lam.rsqrt()) @ V.t()
        self.multiplyByMatrix(P)
        return P


class Masks:
    """The localization masks one rank holds for its colors: what MasksSet /
    Map2Masks / GridMask provide to the path (src/Map2Masks.cc:25-61,
    src/GridMask.h:41-52), as plain tables.  `op` 0 = GridMaskMult (u *= m),
    1 = GridMaskMax (clip |u| <= m), the default orbital masks
    (src/MasksSet.cc:15,158-182)."""

    MULT, MAX = 0, 1

    def __init__(self, grid, mg_levels, subdivx, ncolors, op):
        h = ctypes.c_void_p()
        check(lib().mgb_masks_create(ctypes.byref(h), grid.ref(), mg_levels, subdivx,
                                     ncolors, op))
        self.handle_ = h
        self.mg_levels_, self.subdivx_, self.ncolors_, self.op_ = (
            mg_levels, subdivx, ncolors, op)

    @classmethod
    def from_tables(cls, grid, tables):
        """tables: .mg_levels .subdivx .ncolors .op, .state[level][iloc, color],
        .values[level][(iloc, color)] -> numpy slab (float32 / float64)."""
        m = cls(grid, tables.mg_levels, tables.subdivx, tables.ncolors, tables.op)
        for level in range(tables.mg_levels + 1):
            for iloc in range(tables.subdivx):
                for color in range(tables.ncolors):
                    m.set(level, iloc, color, int(tables.state[level][iloc, color]),
                          tables.values[level].get((iloc, color)))
        m.commit()
        return m

    def set(self, level, iloc, color, state, values=None):
        ptr, dt = None, 1
        if state == 2 and values is not None:
            import numpy as np
            values = np.ascontiguousarray(values)
            dt = {np.dtype(np.float32): 0, np.dtype(np.float64): 1}[values.dtype]
            ptr = values.ctypes.data_as(ctypes.c_void_p)
        check(lib().mgb_masks_set(self.handle_, level, iloc, color, state, dt, ptr))

    def commit(self):
        check(lib().mgb_masks_commit(self.handle_))

    def apply_ghosted(self, gfv, level=0):
        """GridFuncVector::app_mask(level) (src/pb/GridFuncVector.cc:2428-2438)."""
        check(lib().mgb_gfv_app_mask(_dt(gfv.data), self.handle_, level,
                                     gfv.grid().ghost_pt(), _p(gfv.data), gfv.size(),
                                     _stream()))
        gfv.set_updated_boundaries(False)

    def close(self):
        if self.handle_ is not None:
            lib().mgb_masks_destroy(self.handle_)
            self.handle_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LocGridOrbitals(Orbitals):
    """LocGridOrbitals (src/LocGridOrbitals.h:60-, .cc) reduced to the hot
    path: storage npt x chromatic_number "colors"; in x-slab iloc (subdivx slabs
    of the local box) color c holds global orbital overlapping_gids[iloc][c] or
    nothing (-1).  The stencils, the V-cycle and the BLAS-1 run on the color
    block exactly as for ExtendedGridOrbitals; the contractions run per slab
    (LocalMatrices with subdivx sub-matrices) and the masks localize."""

    def __init__(self, grid, numst, overlapping_gids, dtype=torch.float64, psi=None,
                 masks=None):
        import numpy as np
        self.overlapping_gids_ = np.asarray(overlapping_gids, dtype=np.int64)
        self.subdivx_, ncolors = self.overlapping_gids_.shape
        super().__init__(grid, ncolors, dtype, psi)
        self.numst_global_ = numst            # numst_ of the reference
        self.chromatic_number_ = ncolors
        assert grid.dim(0) % self.subdivx_ == 0
        self.loc_numpt_ = grid.size() // self.subdivx_
        self.masks_ = masks

    def getOverlappingGids(self):
        return self.overlapping_gids_

    def _slab(self, t, iloc):
        return t.view(t.shape[0], -1)[:, iloc * self.loc_numpt_:(iloc + 1) * self.loc_numpt_]

    def applyMask(self, first_time=False):
        """src/LocGridOrbitals.cc:427-452."""
        if self.masks_ is not None:
            check(lib().mgb_app_mask(_dt(self.psi_), self.masks_.handle_, 0, _p(self.psi_),
                                     self.grid_.size(), self.chromatic_number_, _stream()))
        self.incrementIterativeIndex()

    def getLocalOverlap(self):
        """ss[iloc] = vel * Phi_iloc^T Phi_iloc (src/LocGridOrbitals.cc:1504-1530:
        LocalMatrices::syrk per slab + fillUpperWithLower + scal)."""
        n = self.chromatic_number_
        ss = torch.empty((self.subdivx_, n, n), dtype=torch.float64, device="cuda")
        check(lib().mgb_syrk_t_slabs(_dt(self.psi_), n, self.loc_numpt_, self.subdivx_,
                                     self.grid_.vel(), _p(self.psi_), self.grid_.size(),
                                     _p(ss), n, _stream()))
        return ss

    def computeLocalProduct(self, other, transpose=False):
        """ss[iloc] = vel * Phi_iloc^T A_iloc, or its transpose
        (src/LocGridOrbitals.cc:1554-1604); returned as row-major
        ss[iloc][i, j] = <a_i, b_j>."""
        arr = other.psi_ if isinstance(other, Orbitals) else other
        a, b = (arr, self.psi_) if transpose else (self.psi_, arr)
        m, n = a.shape[0], b.shape[0]
        ss = torch.empty((self.subdivx_, n, m), dtype=torch.float64, device="cuda")
        check(lib().mgb_gemm_tn_slabs(_dt(self.psi_), m, n, self.loc_numpt_, self.subdivx_,
                                      self.grid_.vel(), _p(a), self.grid_.size(), _p(b),
                                      self.grid_.size(), 0.0, _p(ss), m, _stream()))
        return ss.transpose(1, 2)

    def matrixToLocalMatrix(self, iloc, matrix):
        """lmatrix[icolor, jcolor] = matrix[gid_i, gid_j] where both slots are
        occupied, else 0 (src/LocGridOrbitals.cc:1400-1424).  matrix: (numst,
        numst) tensor indexed [i, j]."""
        g = torch.as_tensor(self.overlapping_gids_[iloc], device=matrix.device)
        ok = g >= 0
        gi = torch.where(ok, g, torch.zeros_like(g))
        lm = matrix[gi][:, gi]
        return lm * (ok[:, None] & ok[None, :]).to(lm.dtype)

    def multiplyByMatrix(self, local_matrices, product=None):
        """Per slab Phi_iloc <- Phi_iloc * M_iloc (src/LocGridOrbitals.cc:793-898);
        local_matrices: (subdivx, n, n) with [iloc][l, j]."""
        n = self.chromatic_number_
        inplace = product is None
        out = torch.empty_like(self.psi_) if inplace else product.psi_
        for iloc in range(self.subdivx_):
            mcol = local_matrices[iloc].t().contiguous()
            check(lib().mgb_gemm_nn(_dt(self.psi_), self.loc_numpt_, n, n, 1.0,
                                    self._slab(self.psi_, iloc).data_ptr(),
                                    self.grid_.size(), _p(mcol), n, 0.0,
                                    self._slab(out, iloc).data_ptr(), self.grid_.size(),
                                    _stream()))
        if inplace:
            self.psi_.copy_(out)
            self.incrementIterativeIndex()
        return out

    def multiply_by_matrix(self, matrix, product=None):
        """Phi * M for a global numst x numst matrix: per slab the block of M
        over the gids present there (src/LocGridOrbitals.cc:750-791)."""
        lms = torch.stack([self.matrixToLocalMatrix(i, matrix)
                           for i in range(self.subdivx_)])
        return self.multiplyByMatrix(lms, product)


class Hamiltonian:
    """Hamiltonian<T> (src/Hamiltonian.h:20-52, .cc:43-159): caches
    hlphi_ = H_loc * phi keyed by 100*phi.index + pot.index."""

    def __init__(self):
        self.lapOper_ = None
        self.pot_ = None
        self.hlphi_ = None
        self.itindex_ = -1

    def setup(self, grid, lap_type):
        self.lapOper_ = LapFactory.createLap(grid, lap_type)
        self.grid_ = grid

    def potential(self, pot=None):
        if pot is not None:
            self.pot_ = pot
        return self.pot_

    def lapOper(self):
        return self.lapOper_

    def applyLocal(self, phi, force=False, xhalo_phi=None, xhalo_v=None, peer_comm=None,
                   vghost=None):
        """src/Hamiltonian.cc:43-83.  peer_comm: decomposed domain whose orbital
        block is registered for direct peer reads (Communicator.register); the
        potential's halo comes as xhalo_v (x slabs: the 2g packed planes) or as vghost
        (any decomposition: a ghosted copy of V with traded boundaries)."""
        assert phi.getIterativeIndex() >= 0 and self.pot_.getIterativeIndex() >= 0
        if (self.hlphi_ is None or self.hlphi_.psi_.shape != phi.psi_.shape
                or self.hlphi_.psi_.dtype != phi.psi_.dtype):
            self.hlphi_ = Orbitals(phi.grid_, phi.numst_, phi.psi_.dtype)
            self.itindex_ = -1
        new_index = 100 * phi.getIterativeIndex() + self.pot_.getIterativeIndex()
        if force or new_index != self.itindex_:
            if peer_comm is not None:
                self.lapOper_.applyWithPotPeer(peer_comm, phi.psi_, self.pot_.vtot(),
                                               self.hlphi_.psi_, xhalo_v, vghost)
            else:
                self.lapOper_.applyWithPot(phi.psi_, self.pot_.vtot(), self.hlphi_.psi_,
                                           xhalo_phi, xhalo_v)
            self.itindex_ = new_index
        return self.hlphi_

    def addHlocal2matrix(self, phi1, phi2, hij, force=False, comm=None):
        """hij += Phi1^T H_loc Phi2 (src/Hamiltonian.cc:163-212)."""
        self.applyLocal(phi2, force)
        return phi1.addDotWithNcol2Matrix(self.hlphi_, hij, comm)

    def addHlocalij(self, phi1, phi2=None, comm=None):
        """Phi1^T H_loc Phi2 (src/Hamiltonian.cc:214-239)."""
        if phi2 is not None:
            self.applyLocal(phi2)
        return phi1.computeLocalProduct(self.hlphi_, comm)


class KBProjectors:
    """The non-local Kleinman-Bylander projectors overlapping the local box, as the sparse
    vectors KBprojectorSparse holds (src/KBprojectorSparse.h:39-52), with the two grid-sized
    steps that follow applyLocal in MGmol::getHpsiAndTheta (src/computeHij.cc:404-455):
    computeKBpsi (src/KBPsiMatrixSparse.cc:136-212) and computeHnlPhiAndAdd2HPhi
    (src/computeHij.cc:294-375)."""

    def __init__(self, grid, dtype=torch.float64):
        self.grid_, self.dtype_ = grid, dtype
        h = ctypes.c_void_p()
        check(lib().mgb_kb_create(ctypes.byref(h), MGB_F64 if dtype == torch.float64 else MGB_F32,
                                  grid.size()))
        self.handle_ = h

    def add_ion(self, nlindex, proj, coeff):
        """nlindex: node positions in the no-ghost storage; proj: (nproj, size_nl) values;
        coeff[p] = kbcoeff_p * sign_p.  Returns the row of the ion's first projector."""
        import numpy as np
        npdt = np.float64 if self.dtype_ == torch.float64 else np.float32
        idx = np.ascontiguousarray(nlindex, dtype=np.int32)
        pr = np.ascontiguousarray(proj, dtype=npdt).reshape(-1, len(idx))
        cf = np.ascontiguousarray(coeff, dtype=np.float64)
        assert len(cf) == pr.shape[0]
        row = ctypes.c_int(0)
        check(lib().mgb_kb_add_ion(self.handle_, len(idx), idx.ctypes.data_as(ctypes.c_void_p),
                                   pr.shape[0], pr.ctypes.data_as(ctypes.c_void_p),
                                   cf.ctypes.data_as(ctypes.c_void_p), ctypes.byref(row)))
        return row.value

    def commit(self):
        check(lib().mgb_kb_commit(self.handle_))

    def nrows(self):
        return lib().mgb_kb_nrows(self.handle_)

    def computeKBpsi(self, orbitals, comm=None, lapOper=None):
        """kbpsi[row, f] = vel <beta_row | psi_f>, summed over the ranks (globalSumKBpsi).
        lapOper: the `flag` of the reference -- project B phi instead of phi (Mehrstellen,
        kbBpsi)."""
        psi = orbitals.psi_ if isinstance(orbitals, Orbitals) else orbitals
        if lapOper is not None:
            psi = lapOper.rhs(psi, torch.empty_like(psi))
        nf = psi.shape[0]
        kbpsi = torch.zeros((self.nrows(), nf), dtype=torch.float64, device="cuda")
        check(lib().mgb_kb_psi(self.handle_, _dt(psi), self.grid_.vel(), _p(psi),
                               self.grid_.size(), nf, _p(kbpsi), _stream()))
        if comm is not None:
            comm.allreduce(kbpsi)
        return kbpsi

    def getVnlPsi(self, kbpsi, nfunc):
        """get_vnlpsi (src/get_vnlpsi.cc:24-87) for every function."""
        out = torch.empty((nfunc,) + tuple(self.grid_.shape()), dtype=self.dtype_, device="cuda")
        check(lib().mgb_kb_vnlpsi(self.handle_, _dt(out), _p(kbpsi), _p(out), self.grid_.size(),
                                  nfunc, 0, _stream()))
        return out

    def computeHnlPhiAndAdd2HPhi(self, kbpsi, hphi, lapOper=None):
        """src/computeHij.cc:294-375: H phi += V_nl phi (one pass over the touched points), or
        with the Mehrstellen operator H phi += B (V_nl phi)."""
        h = hphi.psi_ if isinstance(hphi, Orbitals) else hphi
        nf = h.shape[0]
        if lapOper is not None and lapOper.type_ in (0, 10):
            work = self.getVnlPsi(kbpsi, nf)
            bw = lapOper.rhs(work, torch.empty_like(work))
            check(lib().mgb_axpy(_dt(h), h.numel(), 1.0, _p(bw), _p(h), _stream()))
        else:
            check(lib().mgb_kb_vnlpsi(self.handle_, _dt(h), _p(kbpsi), _p(h), self.grid_.size(),
                                      nf, 1, _stream()))
        if isinstance(hphi, Orbitals):
            hphi.incrementIterativeIndex()
        return hphi

    def close(self):
        if self.handle_ is not None:
            lib().mgb_kb_destroy(self.handle_)
            self.handle_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def computeRhoUsingBlas3(orbitals1, localX, rho, orbitals2=None):
    """Rho::computeRhoSubdomainUsingBlas3 (src/Rho.cc:359-448) on the whole local
    box: rho (double tensor of the grid's shape) += sum_ij X_ij phi1_i phi2_j."""
    o2 = orbitals1 if orbitals2 is None else orbitals2
    n = orbitals1.chromatic_number()
    g = orbitals1.grid_
    xcol = localX.t().contiguous()
    check(lib().mgb_rho_blas3(_dt(orbitals1.psi_), g.size(), n, _p(orbitals1.psi_), g.size(),
                              _p(xcol), n, _p(o2.psi_), g.size(), _p(rho), _stream()))
    return rho


def computeResidualUsingHPhi(lapOper, psi, hphi, localT, res, xhalo_phi=None):
    """MGmol::computeResidualUsingHPhi (src/MGmol.cc:1227-1287):
    res = (B psi) theta - hphi in [Ry]; localT[l, j] = theta (numst x numst double
    tensor).  B psi goes into a library workspace, the contraction's epilogue
    subtracts hphi."""
    n = psi.chromatic_number()
    tcol = localT.t().contiguous()  # column-major
    g = psi.grid_
    check(lib().mgb_residual(lapOper.type_, _dt(psi.psi_), g.ref(), _p(psi.psi_), g.size(),
                             _p(hphi.psi_), g.size(), _p(tcol), n, _p(res.psi_), g.size(), n,
                             _p(xhalo_phi) if xhalo_phi is not None else None, _stream()))
    res.incrementIterativeIndex()
    return res


class OrbitalsPreconditioning:
    """OrbitalsPreconditioning<T> (src/OrbitalsPreconditioning.h:27-70,
    .cc:44-145) over Preconditioning<float> (src/Preconditioning.cc)."""

    def __init__(self):
        self.handle_ = None
        self.gamma_ = -1.0
        self.is_set_ = False

    def setup(self, orbitals, mg_levels, lap_type, masks=None):
        """masks: the `currentMasks` argument of the reference (a Masks set, or
        None for ExtendedGridOrbitals)."""
        assert not self.is_set_
        grid = orbitals.grid_.with_ghosts(ghosts_for(lap_type))
        h = ctypes.c_void_p()
        check(lib().mgb_precond_create(ctypes.byref(h), lap_type, mg_levels,
                                       grid.ref(), orbitals.chromatic_number()))
        if masks is not None:
            check(lib().mgb_precond_set_masks(h, masks.handle_))
        self.masks_ = masks
        self.handle_ = h
        self.lap_type_ = lap_type
        self.mg_levels_ = mg_levels
        self.grid_ = grid
        self.is_set_ = True

    def set_comm(self, comm):
        """Multi-rank boxes: the communicator the V-cycle exchanges ghosts over."""
        check(lib().mgb_precond_set_comm(self.handle_, comm.handle if comm else None))
        self.comm_ = comm

    def set_mode(self, mode):
        """0 automatic, 1 literal (bit-identical to the reference), 2 fused."""
        check(lib().mgb_precond_set_mode(self.handle_, int(mode)))

    def last_mode(self):
        return lib().mgb_precond_last_mode(self.handle_)

    def setGamma(self, lapOper, pot, mg_levels, small_eig):
        """src/OrbitalsPreconditioning.cc:120-145."""
        self.gamma_ = lib().mgb_gamma(lapOper.invDiagEl(), mg_levels, pot.max(),
                                      float(small_eig))
        return self.gamma_

    def precond_mg(self, orbitals):
        """src/OrbitalsPreconditioning.cc:87-117: res <- M^-1 res."""
        assert self.is_set_ and self.gamma_ > 0.0
        psi = orbitals.psi_
        check(lib().mgb_precond_mg(self.handle_, _dt(psi), _p(psi),
                                   orbitals.getNumpt(), orbitals.chromatic_number(),
                                   self.gamma_, _stream()))
        orbitals.incrementIterativeIndex()

    def vcycle(self, gfv_v, gfv_f):
        """Preconditioning<float>::mg (src/Preconditioning.cc:155-216)."""
        check(lib().mgb_precond_vcycle(self.handle_, _p(gfv_v.data), _p(gfv_f.data),
                                       gfv_v.size(), _stream()))

    def close(self):
        if self.handle_ is not None:
            lib().mgb_precond_destroy(self.handle_)
            self.handle_ = None
            self.is_set_ = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
