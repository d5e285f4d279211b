"""Poisson multigrid for the Hartree potential: the host
control flow of pb::SolverLap<Lap<T>, T>::solve (src/pb/SolverLap.cc:62-72) =
pb::Mgm (src/pb/Mgm.h:21-112) over pb::Vcycle (src/pb/Vcycle.h:29-250), on one
scalar field, composed from the same C-ABI operations the orbital path uses
with nfunc = 1: boundary trade, the Laplacians and the Mehrstellen right-hand
side, full-weighting restriction, trilinear prolongation, axpy and dot.

Operators: Laph4M (0), Laph2 (1), Laph4 (2); boundary conditions 0 (zero
Dirichlet) and 1 (periodic) per direction.  Multipole boundary values (bc 2)
and the dielectric (PB) operators are not part of it.

Decomposed boxes (px x py x pz ranks, one GPU each): pass the rank's
`Communicator`.  Fields then trade their boundaries with the neighbour ranks
(mgb_halo_exchange_ghosted: Y, Z, X faces as src/pb/GridFuncVector.cc:1544-1622),
dot products, norms and the average are all-reduced, and the level the local
boxes cannot coarsen any more is gathered on every rank and solved replicated,
as src/pb/Vcycle.h:66-143 does ("gather and solve on all PEs").

The solver works on "fields": objects with the GridFuncVector interface of
host.py.  The default is the device class; the CPU tests drive this very
control flow with a numpy stand-in to pin it against the compiled reference."""
import math
import os

import torch

from .host import GridFuncVector, Lap

# ghost layers the operator needs (Laph4M / Laph2 / Laph4 ::minNumberGhosts)
_MIN_GHOSTS = {0: 1, 1: 1, 2: 2}
# operator used one level down: Laph4 hands the coarse levels to Laph2
# (USE_LOWER_ORDER, src/pb/Vcycle.h:14,181-200; Laph4::getLowerOrderOp)
_LOWER_ORDER = {0: 0, 1: 1, 2: 1}


def _nranks(grid):
    return grid.nproc[0] * grid.nproc[1] * grid.nproc[2]


class DecomposedField(GridFuncVector):
    """pb::GridFunc<T> of a rank of a decomposed box: the boundary trade goes to
    the neighbour ranks, the reductions over the grid are all-reduced
    (GridFunc::gdot / norm2 / get_average with n_mpi_tasks() > 1,
    src/pb/GridFunc.cc:2745-2798, 2856-2861, 2888-2927)."""

    def __init__(self, grid, dtype, comm):
        super().__init__(grid, 1, dtype)
        self.comm_ = comm

    def trade_boundaries(self):
        self.comm_.trade_boundaries(self)

    def gdot(self, other, comm=None):
        return super().gdot(other, self.comm_)

    def get_average(self):
        # local sum / local size, summed over the ranks, / n_mpi_tasks
        t = torch.tensor([super().get_average()], dtype=torch.float64, device="cuda")
        self.comm_.allreduce(t)
        return float(t.item()) / _nranks(self.grid_)

    def gather(self):
        """GridFunc::init_vect(global, 'g'): the function on the whole box, on
        every rank, without ghosts (shape (1, gx, gy, gz)).  Every point is owned
        by one rank, so the sum over zero-padded contributions is exact."""
        from .parallel import local_box
        gr = self.grid_
        glob = torch.zeros((1,) + tuple(gr.gdim_), dtype=torch.float64, device="cuda")
        glob[(slice(None),) + local_box(gr.gdim_, gr.nproc, gr.coord)] = self.values().double()
        self.comm_.allreduce(glob)
        return glob.to(self.data.dtype)


def _device_field(comm):
    def make(grid, dtype):
        if _nranks(grid) > 1:
            return DecomposedField(grid, dtype, comm)
        return GridFuncVector(grid, 1, dtype)
    return make


def _bind(field, dtype, comm=None):
    """field(grid, dtype) -> field of the solver's precision.  A field factory
    serves single-rank grids (the replicated coarse levels) and decomposed ones."""
    make = field if field is not None else _device_field(comm)
    return lambda grid, dt=dtype: make(grid, dt)


def _check(grid, lap_type, who, decomposed=False):
    if lap_type not in _MIN_GHOSTS:
        raise ValueError("%s: operator %d not available" % (who, lap_type))
    if any(b not in (0, 1) for b in grid.bc):
        raise ValueError("%s: boundary conditions 0 and 1 only" % who)
    if _nranks(grid) > 1 and not decomposed:
        raise ValueError("%s: a decomposed box needs the ranks' communicator" % who)


# Lap::jacobi (src/pb/Lap.cc:26-37): W = A x - B; x += scale * W
def _jacobi(lap_type, x, rhs, w, scale):
    x.applyLap(lap_type, w)
    w.axpy(-1.0, rhs)
    x.axpy(scale, w)
    x.set_updated_boundaries(False)
    w.set_updated_boundaries(False)


class PoissonMG:
    """pb::SolverLap<T, T2> (src/pb/SolverLap.h:18-77)."""

    def __init__(self, grid, lap_type, dtype=torch.float64, field=None, comm=None):
        _check(grid, lap_type, "PoissonMG", comm is not None or field is not None)
        self.grid_ = grid.with_ghosts(_MIN_GHOSTS[lap_type])
        self.type_ = lap_type
        self.field_ = _bind(field, dtype, comm)
        self.gather_coarse_level_ = True
        # device fields: the V-cycle -- a fixed sequence of ~200 small kernels down to the
        # 1^3 level -- is captured once in a CUDA graph and replayed every sweep, so the
        # host launches one graph instead of ~200 kernels (MGB_POISSON_GRAPH=0: eager).
        # Decomposed boxes run it eagerly (the exchanges are NCCL calls).
        self.use_graph_ = (field is None and _nranks(grid) == 1
                           and os.environ.get("MGB_POISSON_GRAPH", "1") != "0")
        self.graph_ = None
        self.graph_replays = 0
        self.fully_periodic_ = tuple(grid.bc) == (1, 1, 1)
        self.setup(2, 2, 10, 1.e-16, 10)
        self.nb_sweeps_ = 0
        self.final_residual_ = -1.
        self.final_relative_residual_ = -1.
        self.residual_reduction_ = -1.

    def setup(self, nu1, nu2, max_sweeps, tol, max_nlevels, gather_coarse_level=True):
        self.nu1_, self.nu2_ = int(nu1), int(nu2)
        self.max_sweeps_ = int(max_sweeps)
        self.tol_ = float(tol)
        self.max_nlevels_ = int(max_nlevels)
        self.gather_coarse_level_ = bool(gather_coarse_level)
        self.graph_ = None  # the captured cycle depends on nu1, nu2 and the level count

    def _cycle(self, work1, res):
        """work1 = Vcycle(0, res): eager the first time (warm-up), then a captured graph
        over two static fields."""
        if not self.use_graph_ or not torch.cuda.is_available():
            work1.resetData()
            self._vcycle(self.type_, work1, res, self.max_nlevels_)
            return work1
        if self.graph_ is None:
            F, grid = self.field_, self.grid_
            self.g_res_, self.g_work_ = F(grid), F(grid)
            # the residual always arrives with stale ghosts: the captured sequence trades them
            self.g_res_.copy_from(res)
            self.g_res_.set_updated_boundaries(False)
            self.g_work_.resetData()
            self._vcycle(self.type_, self.g_work_, self.g_res_, self.max_nlevels_)  # warm-up
            torch.cuda.synchronize()
            self.graph_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_):
                self.g_res_.set_updated_boundaries(False)
                self.g_work_.resetData()
                self._vcycle(self.type_, self.g_work_, self.g_res_, self.max_nlevels_)
        self.g_res_.copy_from(res)
        self.g_res_.set_updated_boundaries(False)
        self.graph_.replay()
        self.graph_replays += 1
        self.g_work_.set_updated_boundaries(False)
        return self.g_work_

    def getNbSweeps(self):
        return self.nb_sweeps_

    def getFinalResidual(self):
        return self.final_residual_

    def getFinalRelativeResidual(self):
        return self.final_relative_residual_

    def getResidualReduction(self):
        return self.residual_reduction_

    def _regrid(self, src, grid):
        """The same function on a grid with another ghost width."""
        out = self.field_(grid)
        out.assign(src.values())
        return out

    # -- pb::Vcycle (src/pb/Vcycle.h:29-250), x = 0 on entry --------------------
    def _vcycle(self, lap_type, x, rhs, cogr, gather=None):
        grid = x.grid()
        g = grid.ghost_pt()
        flag_coarsen = all(grid.dim(d) % 2 == 0 and grid.dim(d) >= 2 * g for d in range(3))
        if gather is None:
            gather = self.gather_coarse_level_
        if (not flag_coarsen or grid.level_ <= -cogr) and gather and _nranks(grid) > 1:
            # src/pb/Vcycle.h:66-143: the local boxes cannot be coarsened any further --
            # gather this level on every rank, run the rest of the cycle replicated on
            # the whole box, keep the local part ("gather and solve on all PEs")
            from .host import Grid
            from .parallel import local_box
            rgrid = Grid(grid.gdim_, grid.ll_, g, grid.bc, (1, 1, 1), (0, 0, 0), grid.level_)
            rrhs, rx = self.field_(rgrid), self.field_(rgrid)
            rrhs.assign(rhs.gather())
            self._vcycle(lap_type, rx, rrhs, cogr - grid.level_, gather=False)
            box = (slice(None),) + local_box(grid.gdim_, grid.nproc, grid.coord)
            mine = rx.values()[box]
            x.assign(mine.contiguous() if hasattr(mine, "contiguous") else mine.copy())
            return
        lap = Lap(grid, lap_type)
        scale = -1. * lap.jacobiFactor()
        res = self.field_(grid)
        for _ in range(self.nu1_):
            _jacobi(lap_type, x, rhs, res, scale)
        if grid.level_ > -cogr and flag_coarsen:
            x.applyLap(lap_type, res)
            res.axpy(-1.0, rhs)
            gc = max(g - 1, 1)
            coarse_type = _LOWER_ORDER[lap_type]
            coarse_same = grid.coarse_grid()               # ghosts of this level
            coarse_grid = coarse_same.with_ghosts(gc)
            if gc == g:
                rcoarse = self.field_(coarse_grid)
                res.restrict3D(rcoarse)
            else:
                tmp = self.field_(coarse_same)
                res.restrict3D(tmp)
                rcoarse = self._regrid(tmp, coarse_grid)
            rcoarse.set_updated_boundaries(False)
            ucoarse = self.field_(coarse_grid)
            self._vcycle(coarse_type, ucoarse, rcoarse, cogr, gather)
            if gc == g:
                res.extend3D(ucoarse)
            else:
                res.extend3D(self._regrid(ucoarse, coarse_same))
            x.axpy(-1.0, res)
        for _ in range(self.nu2_):
            _jacobi(lap_type, x, rhs, res, scale)

    # -- pb::Mgm (src/pb/Mgm.h:21-112) + SolverLap::solve --------------------------
    def solve(self, vh, rho):
        """vh (in: initial guess, out: solution) and rho: no-ghost fields of
        the solver's precision, shape (nx, ny, nz).  Returns `converged`."""
        F, grid, lt = self.field_, self.grid_, self.type_
        shape1 = (1,) + tuple(grid.shape())
        gf_vh, gf_rho = F(grid), F(grid)
        gf_vh.assign(vh.reshape(shape1))
        gf_rho.assign(rho.reshape(shape1))
        res = F(grid)
        res.copy_from(gf_rho)
        rhs = F(grid)
        if lt == 0:
            res.applyRHS(0, rhs)         # Laph4M::rhs -> rhs_4th_Mehr1
        else:
            rhs.copy_from(res)           # FDoper::rhs: B = A
        lhs, work1 = F(grid), F(grid)
        inv_rhs_norm = 1. / rhs.norm2()
        init_residual_norm = 1.
        converged = False
        self.nb_sweeps_ = 0
        for i in range(self.max_sweeps_):
            gf_vh.applyLap(lt, lhs)
            res.copy_from(rhs)           # res.diff(rhs, lhs)
            res.axpy(-1.0, lhs)
            res_norm = res.norm2()
            if i == 0:
                init_residual_norm = res_norm
            if res_norm * inv_rhs_norm < self.tol_:
                self.final_residual_ = res_norm
                self.final_relative_residual_ = res_norm * inv_rhs_norm
                converged = True
                break
            corr = self._cycle(work1, res)
            self.nb_sweeps_ += 1
            gf_vh.axpy(1.0, corr)
        if not converged:
            gf_vh.applyLap(lt, lhs)
            lhs.axpy(-1.0, rhs)
            self.final_residual_ = lhs.norm2()
            self.final_relative_residual_ = self.final_residual_ * inv_rhs_norm
        self.residual_reduction_ = self.final_residual_ / init_residual_norm
        if self.fully_periodic_:
            gf_vh.add_scalar(-gf_vh.get_average())   # GridFunc::average0
        gf_vh.getValues(vh.reshape(shape1))
        return converged


class PoissonPCG:
    """PCGSolver<T, ScalarType> (src/PCGSolver.h:20-106, src/PCGSolver.cc): left
    preconditioned conjugate gradient on A u = rhs, the preconditioner one
    multigrid V-cycle in POISSONPRECONDTYPE = float (preconSolve,
    src/PCGSolver.cc:112-162): the solver's operator on the fine level
    (Control::lap_type in the reference), Laph2 below, the coarse grids keeping
    the fine grid's ghost width."""

    def __init__(self, grid, lap_type, dtype=torch.float64, field=None, precond_dtype=None,
                 comm=None):
        _check(grid, lap_type, "PoissonPCG", comm is not None or field is not None)
        self.grid_ = grid.with_ghosts(_MIN_GHOSTS[lap_type])
        self.type_ = lap_type
        self.field_ = _bind(field, dtype, comm)
        if precond_dtype is None:
            precond_dtype = torch.float32
        # the preconditioner's levels are the local boxes' own (setupPrecon coarsens
        # mygrid without a gather, src/PCGSolver.cc:50-110)
        self.pfield_ = _bind(field, precond_dtype, comm)
        self.fully_periodic_ = tuple(grid.bc) == (1, 1, 1)
        self.final_residual_ = -1.
        self.residual_reduction_ = -1.
        self.grids_ = None
        self.setup(2, 2, 10, 1.e-16, 10)

    def setup(self, nu1, nu2, max_sweeps, tol, max_nlevels):
        self.nu1_, self.nu2_ = int(nu1), int(nu2)
        self.maxiters_ = int(max_sweeps)
        self.tol_ = float(tol)
        self.max_nlevels_ = int(max_nlevels)
        self._setup_precon()

    def getFinalResidual(self):
        return self.final_residual_

    def getResidualReduction(self):
        return self.residual_reduction_

    # -- setupPrecon (src/PCGSolver.cc:50-110) ------------------------------------
    def _setup_precon(self):
        g = self.grid_.ghost_pt()
        grids, mygrid = [self.grid_], self.grid_
        self.nlevels_ = self.max_nlevels_
        for ln in range(1, self.max_nlevels_ + 1):
            if not all(mygrid.dim(d) % 2 == 0 and mygrid.dim(d) >= 2 * g for d in range(3)):
                self.nlevels_ = ln - 1
                break
            mygrid = mygrid.coarse_grid()
            grids.append(mygrid)
        self.grids_ = grids
        self.ptype_ = [self.type_] + [1] * (len(grids) - 1)
        self.pscale_ = [-1. * Lap(gr, t).jacobiFactor() for gr, t in zip(grids, self.ptype_)]
        self.work_ = [self.pfield_(gr) for gr in grids]
        self.rcoarse_ = [self.pfield_(gr) for gr in grids[1:]]
        self.newv_ = [self.pfield_(gr) for gr in grids[1:]]

    # -- preconSolve (src/PCGSolver.cc:112-162) ------------------------------------
    def _precon_solve(self, v, f, level):
        last = level == self.nlevels_
        ncycl = max(4, self.nu1_ + self.nu2_) if last else self.nu1_
        lt, scale, work = self.ptype_[level], self.pscale_[level], self.work_[level]
        for _ in range(ncycl):
            _jacobi(lt, v, f, work, scale)
        if last:
            return
        rcoarse, newv = self.rcoarse_[level], self.newv_[level]
        work.restrict3D(rcoarse)
        rcoarse.set_updated_boundaries(False)
        newv.resetData()
        self._precon_solve(newv, rcoarse, level + 1)
        work.extend3D(newv)
        v.axpy(-1.0, work)
        for _ in range(self.nu2_):
            _jacobi(lt, v, f, work, scale)
        bc = self.grid_.bc
        if bc[0] != 1 or bc[2] != 1:      # as written at src/PCGSolver.cc:161
            v.trade_boundaries()

    # -- solve (src/PCGSolver.cc:165-252) ---------------------------------------------
    def solve(self, vh, rho):
        """vh (in: initial guess, out: solution) and rho (the right-hand side of
        A u = rho): no-ghost fields, shape (nx, ny, nz).  Returns `converged`."""
        F, P, grid, lt = self.field_, self.pfield_, self.grid_, self.type_
        shape1 = (1,) + tuple(grid.shape())
        gf_phi, gf_rhs = F(grid), F(grid)
        gf_phi.assign(vh.reshape(shape1))
        gf_rhs.assign(rho.reshape(shape1))
        lhs = F(grid)
        gf_phi.applyLap(lt, lhs)
        res = F(grid)
        res.copy_from(gf_rhs)
        res.axpy(-1.0, lhs)
        init_rnorm = res.norm2()
        if init_rnorm < 1.e-24:
            return True
        rnorm = init_rnorm
        prec_z, prec_res = P(grid), P(grid)
        prec_res.assign(res.values())         # GridFunc<float>(res)
        prec_z.resetData()
        self._precon_solve(prec_z, prec_res, 0)
        z, p, ap = F(grid), F(grid), F(grid)
        z.assign(prec_z.values())
        p.assign(prec_z.values())
        rtz = res.gdot(z)
        converged = False
        for _ in range(self.maxiters_):
            p.applyLap(lt, ap)
            ptap = p.gdot(ap)
            alp = rtz / ptap
            gf_phi.axpy(alp, p)
            res.axpy(-alp, ap)
            rnorm = res.norm2()
            if rnorm <= self.tol_ * init_rnorm:
                converged = True
                break
            prec_z.resetData()
            prec_res.assign(res.values())
            self._precon_solve(prec_z, prec_res, 0)
            z.assign(prec_z.values())
            rtz_new = res.gdot(z)
            bet = rtz_new / rtz
            p.scal(bet)
            p.axpy(1.0, z)
            rtz = rtz_new
        self.final_residual_ = rnorm
        self.residual_reduction_ = rnorm / init_rnorm
        if self.fully_periodic_:
            gf_phi.add_scalar(-gf_phi.get_average())
        gf_phi.getValues(vh.reshape(shape1))
        return converged


class Hartree:
    """Hartree<T> (src/Hartree.h:18-46, src/Hartree.cc:27-112) with the state of
    its Poisson base (src/Poisson.h:31-89): the Hartree potential vh, kept between
    calls as the next initial guess, and the integrals of vh against the charges.
    `solve(rho, rhoc)`: rhs = 4 pi (rho - rhoc) in the solver's precision, then the
    Poisson solver; boundary conditions 0 / 1 (no multipole boundary values)."""

    def __init__(self, grid, lap_type, dtype=torch.float64, field=None, pcg=False,
                 rho_dtype=torch.float64, precond_dtype=None, comm=None):
        self.field_ = _bind(field, dtype, comm)
        self.rfield_ = _bind(field, rho_dtype, comm)
        if pcg:
            self.poisson_solver_ = PoissonPCG(grid, lap_type, dtype, field, precond_dtype, comm)
        else:
            self.poisson_solver_ = PoissonMG(grid, lap_type, dtype, field, comm)
        self.grid_ = self.poisson_solver_.grid_
        zero = self.field_(self.grid_)
        zero.resetData()
        self.vh_ = zero.values()
        self.Int_vhrho_ = self.Int_vhrhoc_ = self.Int_vhrho_old_ = 0.

    def setup(self, nu1, nu2, max_sweeps, tol, max_nlevels, gather_coarse_level=True):
        if isinstance(self.poisson_solver_, PoissonMG):
            self.poisson_solver_.setup(nu1, nu2, max_sweeps, tol, max_nlevels, gather_coarse_level)
        else:  # Hartree_CG ignores the flag (src/Hartree_CG.h:35-37)
            self.poisson_solver_.setup(nu1, nu2, max_sweeps, tol, max_nlevels)

    def vh(self):
        return self.vh_

    def set_vh(self, vh):
        f = self.field_(self.grid_)
        f.assign(vh.reshape((1,) + tuple(self.grid_.shape())))
        self.vh_ = f.values()

    def resetVh(self):
        f = self.field_(self.grid_)
        f.resetData()
        self.vh_ = f.values()

    def IntVhRho(self):
        return self.Int_vhrho_

    def IntVhRhoc(self):
        return self.Int_vhrhoc_

    def IntVhRho_old(self):
        return self.Int_vhrho_old_

    def getResidualReduction(self):
        return self.poisson_solver_.getResidualReduction()

    def getFinalResidual(self):
        return self.poisson_solver_.getFinalResidual()

    def _vh_dot(self, charge):
        """vel * vh_->gdot(charge), the charge converted to vh's precision the
        way the mixed-type gdot reads both operands as double."""
        a, b = self.field_(self.grid_), self.field_(self.grid_)
        a.assign(self.vh_)
        b.assign(charge)
        return self.grid_.vel() * a.gdot(b)

    def solve(self, rho, rhoc):
        """rho, rhoc: no-ghost fields (nx, ny, nz) of the charge precision."""
        shape1 = (1,) + tuple(self.grid_.shape())
        rho1, rhoc1 = rho.reshape(shape1), rhoc.reshape(shape1)
        self.Int_vhrho_old_ = self._vh_dot(rho1)
        work_rho, gf_rhoc = self.rfield_(self.grid_), self.rfield_(self.grid_)
        work_rho.assign(rho1)
        gf_rhoc.assign(rhoc1)
        work_rho.axpy(-1.0, gf_rhoc)                 # work_rho -= rhoc
        rhs = self.field_(self.grid_)
        rhs.assign(work_rho.values())                # GridFunc<POTDTYPE> rhs(work_rho)
        rhs.scal(4. * math.pi)                       # Hartree units
        conv = self.poisson_solver_.solve(self.vh_.reshape(tuple(self.grid_.shape())),
                                          rhs.values().reshape(tuple(self.grid_.shape())))
        self.Int_vhrho_ = self._vh_dot(rho1)
        self.Int_vhrhoc_ = self._vh_dot(rhoc1)
        return conv
