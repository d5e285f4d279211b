// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// The reference's preconditioned conjugate gradient Poisson solver through its
// own class: PCGSolver<Lap<T>, T>::solve (src/PCGSolver.cc:165-252), multigrid
// preconditioner in POISSONPRECONDTYPE = float (preconSolve, :112-162).  The
// shim owns no numerics.  Compiled with control_stub/Control.h pre-included.
#include <iostream>
#include <sstream>

#include <mpi.h>

#include "Grid.h"
#include "GridFunc.h"
#include "Laph2.h"
#include "Laph4.h"
#include "Laph4M.h"
#include "MGmol_MPI.h"
#include "PCGSolver.h"
#include "PEenv.h"

namespace
{
template <class LapT, typename T>
int solve(const int* dims, int ghosts, const double* ll, const int* bc, T* vh, const T* rho,
    int nu1, int nu2, int max_sweeps, double tol, int max_nlevels, double* stats)
{
    const double origin[3] = { 0., 0., 0. };
    const unsigned ud[3]   = { (unsigned)dims[0], (unsigned)dims[1], (unsigned)dims[2] };
    pb::PEenv pe(MPI_COMM_WORLD, dims[0], dims[1], dims[2]);
    pb::Grid grid(origin, ll, ud, pe, (short)ghosts, 0);
    LapT oper(grid);
    PCGSolver<LapT, T> solver(oper, (short)bc[0], (short)bc[1], (short)bc[2]);
    solver.setup((short)nu1, (short)nu2, (short)max_sweeps, tol, (short)max_nlevels);
    pb::GridFunc<T> gf_vh(grid, (short)bc[0], (short)bc[1], (short)bc[2]);
    pb::GridFunc<T> gf_rho(grid, (short)bc[0], (short)bc[1], (short)bc[2]);
    gf_vh.assign(vh);
    gf_rho.assign(rho);
    const bool conv = solver.solve(gf_vh, gf_rho);
    gf_vh.init_vect(vh, 'd');
    stats[0] = solver.getFinalResidual();
    stats[1] = solver.getResidualReduction();
    return conv ? 1 : 0;
}
}

extern "C" void ref_poisson_ensure_mpi(void);

extern "C" int ref_pcg_solve(int lap_type, int dtype, const int* dims, const double* ll,
    const int* bc, void* vh, const void* rho, int nu1, int nu2, int max_sweeps, double tol,
    int max_nlevels, double* stats)
{
    ref_poisson_ensure_mpi();
    // the preconditioner's fine-level operator follows Control::lap_type
    Control::instance()->lap_type = (short)lap_type;
#define RUN(L, G)                                                                         \
    (dtype == 1 ? solve<pb::L<double>, double>(dims, G, ll, bc, (double*)vh,              \
                      (const double*)rho, nu1, nu2, max_sweeps, tol, max_nlevels, stats)  \
                : solve<pb::L<float>, float>(dims, G, ll, bc, (float*)vh,                 \
                      (const float*)rho, nu1, nu2, max_sweeps, tol, max_nlevels, stats))
    switch (lap_type)
    {
        case 0:
            return RUN(Laph4M, 1);
        case 1:
            return RUN(Laph2, 1);
        case 2:
            return RUN(Laph4, 2);
        default:
            std::cerr << "ref_pcg_solve: operator " << lap_type << " not wired" << std::endl;
            return -1;
    }
#undef RUN
}
