// The Poisson solver templates of include/mgmol_b200_poisson.hpp on one problem
// read from stdin:
//   <cpu|gpu> as argv[1]: host fields over the oracle's kernels (no device), or
//   device fields over the C ABI
//   stdin: solver(0 = SolverLap/Mgm, 1 = PCGSolver, 2 = Hartree over Mgm, 3 = Hartree
//          over PCGSolver) lap_type dtype(0 f32, 1 f64)
//          nx ny nz  lx ly lz  bcx bcy bcz  nu1 nu2 max_sweeps tol max_nlevels
//          then vh0 (nx*ny*nz values) and rho (nx*ny*nz values); Hartree: rho then rhoc
//          (double charges; vh0 is the kept potential of the first solve)
//   stdout: converged nb_sweeps final_residual final_relative_residual
//           residual_reduction, then the solution, one value per line.
//           Hartree: two consecutive solves with the same charges, one line each
//           (converged IntVhRho_old IntVhRho IntVhRhoc final_residual), then vh.
#include <cstdio>
#include <cstring>
#include <vector>

#include "mgmol_b200_poisson.hpp"
#include "oracle_field.hpp"

using namespace mgmol_b200;

struct Problem
{
    int solver, lap, dtype, bc[3];
    unsigned dims[3];
    double ll[3], tol;
    int nu1, nu2, max_sweeps, max_nlevels;
    std::vector<double> vh, rho;
};

template <class Solver, typename T>
static void report(Solver& s, bool conv, int sweeps, double relres, const std::vector<T>& vh)
{
    std::printf("%d %d %.17g %.17g %.17g\n", (int)conv, sweeps, s.getFinalResidual(), relres,
        s.getResidualReduction());
    for (size_t i = 0; i < vh.size(); i++)
        std::printf("%.17g\n", (double)vh[i]);
}

// host pointers go straight in (cpu) or through device buffers (gpu)
template <class Solver, typename T>
static bool run(Solver& s, std::vector<T>& vh, const std::vector<T>& rho, bool device)
{
    if (!device) return s.solve(vh.data(), rho.data());
    DeviceMemory<T> dvh(vh.size()), drho(rho.size());
    dvh.copy_to_dev(vh.data(), vh.size());
    drho.copy_to_dev(rho.data(), rho.size());
    MGB_CHECK(mgb_stream_sync(nullptr));
    const bool conv = s.solve(dvh.data(), drho.data());
    dvh.copy_to_host(vh.data(), vh.size());
    return conv;
}

template <typename T, class Field, class PField>
static void solve(const Problem& p, bool device)
{
    Grid grid(p.dims, p.ll, 1, p.bc);
    std::vector<T> vh(p.vh.begin(), p.vh.end()), rho(p.rho.begin(), p.rho.end());
    if (p.solver == 0)
    {
        PoissonMG<Field> s(grid, p.lap);
        s.setup(p.nu1, p.nu2, p.max_sweeps, p.tol, p.max_nlevels);
        const bool conv = run(s, vh, rho, device);
        report(s, conv, s.getNbSweeps(), s.getFinalRelativeResidual(), vh);
    }
    else
    {
        PoissonPCG<Field, PField> s(grid, p.lap);
        s.setup(p.nu1, p.nu2, p.max_sweeps, p.tol, p.max_nlevels);
        const bool conv = run(s, vh, rho, device);
        report(s, conv, -1, -1., vh);
    }
}

// Hartree<T>::solve twice: the second solve starts from the first one's potential
template <typename T, class Field, class PField, class Solver>
static void solve_hartree(const Problem& p, const std::vector<double>& rhoc, bool device)
{
    Grid grid(p.dims, p.ll, 1, p.bc);
    typedef typename Field::template rebind<double>::type RField;
    Hartree<Field, RField, Solver> h(grid, p.lap);
    h.setup(p.nu1, p.nu2, p.max_sweeps, p.tol, p.max_nlevels);
    const size_t n = p.vh.size();
    std::vector<T> vh0(p.vh.begin(), p.vh.end()), vh(n);
    if (!device)
    {
        h.set_vh(vh0.data());
        for (int k = 0; k < 2; k++)
        {
            const bool conv = h.solve(p.rho.data(), rhoc.data());
            std::printf("%d %.17g %.17g %.17g %.17g\n", (int)conv, h.IntVhRho_old(), h.IntVhRho(),
                h.IntVhRhoc(), h.getFinalResidual());
        }
        std::copy(h.vh(), h.vh() + n, vh.begin());
    }
    else
    {
        DeviceMemory<T> dvh(n);
        DeviceMemory<double> drho(n), drhoc(n);
        dvh.copy_to_dev(vh0.data(), n);
        drho.copy_to_dev(p.rho.data(), n);
        drhoc.copy_to_dev(rhoc.data(), n);
        MGB_CHECK(mgb_stream_sync(nullptr));
        h.set_vh(dvh.data());
        for (int k = 0; k < 2; k++)
        {
            const bool conv = h.solve(drho.data(), drhoc.data());
            std::printf("%d %.17g %.17g %.17g %.17g\n", (int)conv, h.IntVhRho_old(), h.IntVhRho(),
                h.IntVhRhoc(), h.getFinalResidual());
        }
        MGB_CHECK(mgb_copy_to_host(vh.data(), h.vh(), n * sizeof(T), nullptr));
        MGB_CHECK(mgb_stream_sync(nullptr));
    }
    for (size_t i = 0; i < n; i++)
        std::printf("%.17g\n", (double)vh[i]);
}

int main(int argc, char** argv)
{
    if (argc != 2) return 2;
    const bool device = std::strcmp(argv[1], "gpu") == 0;
    if (device && mgb_device_count() < 1)
    {
        std::printf("no CUDA device\n");
        return 77;
    }
    Problem p;
    if (std::scanf("%d %d %d %u %u %u %lf %lf %lf %d %d %d %d %d %d %lf %d", &p.solver, &p.lap,
            &p.dtype, &p.dims[0], &p.dims[1], &p.dims[2], &p.ll[0], &p.ll[1], &p.ll[2], &p.bc[0],
            &p.bc[1], &p.bc[2], &p.nu1, &p.nu2, &p.max_sweeps, &p.tol, &p.max_nlevels)
        != 17)
        return 2;
    const size_t n = (size_t)p.dims[0] * p.dims[1] * p.dims[2];
    p.vh.resize(n);
    p.rho.resize(n);
    for (size_t i = 0; i < n; i++)
        if (std::scanf("%lf", &p.vh[i]) != 1) return 2;
    for (size_t i = 0; i < n; i++)
        if (std::scanf("%lf", &p.rho[i]) != 1) return 2;
    if (p.solver >= 2)
    {
        std::vector<double> rhoc(n);
        for (size_t i = 0; i < n; i++)
            if (std::scanf("%lf", &rhoc[i]) != 1) return 2;
        const bool pcg = p.solver == 3;
#define HARTREE(T, F, PF)                                                                      \
    (pcg ? solve_hartree<T, F<T>, PF<float>, PoissonPCG<F<T>, PF<float>>>(p, rhoc, device)      \
         : solve_hartree<T, F<T>, PF<float>, PoissonMG<F<T>>>(p, rhoc, device))
        if (device)
        {
            if (p.dtype == 1)
                HARTREE(double, GridFunc, GridFunc);
            else
                HARTREE(float, GridFunc, GridFunc);
        }
        else
        {
            if (p.dtype == 1)
                HARTREE(double, OracleField, OracleField);
            else
                HARTREE(float, OracleField, OracleField);
        }
#undef HARTREE
        return 0;
    }
    if (device)
    {
        if (p.dtype == 1)
            solve<double, GridFunc<double>, GridFunc<float>>(p, true);
        else
            solve<float, GridFunc<float>, GridFunc<float>>(p, true);
    }
    else
    {
        if (p.dtype == 1)
            solve<double, OracleField<double>, OracleField<float>>(p, false);
        else
            solve<float, OracleField<float>, OracleField<float>>(p, false);
    }
    return 0;
}
