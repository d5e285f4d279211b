// C-ABI entry of the Hartree Poisson solvers.  The solvers
// themselves are the header templates of include/mgmol_b200_poisson.hpp -- host
// control flow of pb::SolverLap / pb::Mgm / pb::Vcycle and PCGSolver over this
// library's own grid operations with one function; this file instantiates them
// on the device field and gives them a plain-C face.
#include <cmath>
#include <cstring>

#include <new>

#include "common.cuh"
// The header templates end the run on a failed call, like the reference; behind
// a C entry that returns codes they must not: every failure throws
// mgmol_b200::Error here and becomes the return value below.
#define MGMOL_B200_ERRORS_THROW
#include "mgmol_b200_poisson.hpp"

using namespace mgb;

namespace
{
template <typename T>
int solve_t(int solver, int lap_type, const mgmol_b200::Grid& grid, T* vh, const T* rho, int nu1,
    int nu2, int max_sweeps, double tol, int max_nlevels, double* stats)
{
    using namespace mgmol_b200;
    bool conv;
    if (solver == MGB_POISSON_MG)
    {
        PoissonMG<GridFunc<T>> s(grid, lap_type);
        s.setup((short)nu1, (short)nu2, (short)max_sweeps, tol, (short)max_nlevels);
        conv = s.solve(vh, rho);
        if (stats)
        {
            stats[1] = s.getNbSweeps();
            stats[2] = s.getFinalResidual();
            stats[3] = s.getFinalRelativeResidual();
            stats[4] = s.getResidualReduction();
        }
    }
    else
    {
        PoissonPCG<GridFunc<T>, GridFunc<float>> s(grid, lap_type);
        s.setup((short)nu1, (short)nu2, (short)max_sweeps, tol, (short)max_nlevels);
        conv = s.solve(vh, rho);
        if (stats)
        {
            stats[1] = -1.;
            stats[2] = s.getFinalResidual();
            stats[3] = -1.;
            stats[4] = s.getResidualReduction();
        }
    }
    if (stats) stats[0] = conv ? 1. : 0.;
    return MGB_OK;
}
}

extern "C" int mgb_poisson_solve(int solver, int lap_type, int dtype, const mgb_grid* grid,
    void* vh, const void* rho, int nu1, int nu2, int max_sweeps, double tol, int max_nlevels,
    double* stats)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(grid)) return rc;
    MGB_REQUIRE(vh && rho, "mgb_poisson_solve: null pointer");
    MGB_REQUIRE(solver == MGB_POISSON_MG || solver == MGB_POISSON_PCG,
        "mgb_poisson_solve: solver %d (0 = multigrid, 1 = preconditioned CG)", solver);
    MGB_REQUIRE(lap_type == MGB_LAP_4M || lap_type == MGB_LAP_2 || lap_type == MGB_LAP_4,
        "mgb_poisson_solve: operator %d not available (Laph4M, Laph2, Laph4)", lap_type);
    MGB_REQUIRE(dtype == MGB_F32 || dtype == MGB_F64, "mgb_poisson_solve: bad dtype %d", dtype);
    MGB_REQUIRE(nu1 >= 0 && nu2 >= 0 && max_sweeps >= 0 && max_nlevels >= 0,
        "mgb_poisson_solve: negative parameter");
    for (int d = 0; d < 3; d++)
        MGB_REQUIRE(grid->nproc[d] == 1, "mgb_poisson_solve: single-rank boxes only");
    unsigned gdim[3];
    double ll[3];
    for (int d = 0; d < 3; d++)
    {
        gdim[d] = (unsigned)grid->gdim[d];
        // the lattice length whose quotient by gdim is exactly the caller's h
        // (the solver derives its coefficients from ll / gdim)
        const double n = (double)grid->gdim[d], h = grid->h[d];
        double cand[5] = { h * n, 0., 0., 0., 0. };
        cand[1]        = std::nextafter(cand[0], 0.);
        cand[2]        = std::nextafter(cand[0], 2. * cand[0]);
        cand[3]        = std::nextafter(cand[1], 0.);
        cand[4]        = std::nextafter(cand[2], 2. * cand[0]);
        ll[d]          = cand[0];
        for (int k = 0; k < 5; k++)
            if (cand[k] / n == h)
            {
                ll[d] = cand[k];
                break;
            }
    }
    try
    {
        const mgmol_b200::Grid g(
            gdim, ll, (short)grid->ghosts, grid->bc, grid->nproc, grid->coord);
        if (dtype == MGB_F64)
            return solve_t<double>(solver, lap_type, g, (double*)vh, (const double*)rho, nu1, nu2,
                max_sweeps, tol, max_nlevels, stats);
        return solve_t<float>(solver, lap_type, g, (float*)vh, (const float*)rho, nu1, nu2,
            max_sweeps, tol, max_nlevels, stats);
    }
    catch (const mgmol_b200::Error& e)
    {
        // mgb_last_error() already holds the failing call's message (e.g. the
        // cudaMalloc of a level's work field); the level fields allocated so
        // far were released while the exception unwound
        if (e.rc == MGB_EINVAL && std::strcmp(e.where, "precondition") == 0)
            set_error("mgb_poisson_solve: a precondition of the solver templates failed "
                      "(message on stderr)");
        return e.rc;
    }
    catch (const std::bad_alloc&)
    {
        set_error("mgb_poisson_solve: host allocation failed");
        return MGB_ECUDA;
    }
}
