// The error policy of include/mgmol_b200.hpp behind a C entry
// (MGMOL_B200_ERRORS_THROW, used by mgmol_b200/csrc/poisson.cu): failed
// preconditions and failed library calls throw mgmol_b200::Error with the
// MGB_E* code instead of ending the process, and what was allocated so far is
// released while the exception unwinds.  Runs without a GPU: on a box without a
// device every compute call fails with MGB_ENODEVICE, which is the failure used.
#define MGMOL_B200_ERRORS_THROW
#include "mgmol_b200_poisson.hpp"

#include <cstdio>

// Every member of the host mirror compiles (members of class templates are otherwise only
// checked when used): the decomposed-domain entry points, the KB projectors, the solvers.
template class mgmol_b200::GridFuncVector<double>;
template class mgmol_b200::GridFuncVector<float>;
template class mgmol_b200::ExtendedGridOrbitals<double>;
template class mgmol_b200::ExtendedGridOrbitals<float>;
template class mgmol_b200::LocGridOrbitals<double>;
template class mgmol_b200::Lap<double>;
template class mgmol_b200::Lap<float>;
template class mgmol_b200::Hamiltonian<double>;
template class mgmol_b200::Hamiltonian<float>;
template class mgmol_b200::OrbitalsPreconditioning<double>;
template class mgmol_b200::OrbitalsPreconditioning<float>;
template class mgmol_b200::KBProjectors<double>;
template class mgmol_b200::KBProjectors<float>;
template class mgmol_b200::PoissonMG<mgmol_b200::GridFunc<double>>;
template class mgmol_b200::PoissonPCG<mgmol_b200::GridFunc<double>, mgmol_b200::GridFunc<float>>;
template void mgmol_b200::Communicator::trade_boundaries<double>(
    mgmol_b200::GridFuncVector<double>&, void*);
template void mgmol_b200::Communicator::trade_boundaries<float>(
    mgmol_b200::GridFuncVector<float>&, void*);
template void mgmol_b200::Communicator::haloExchangeX<double>(
    const mgmol_b200::Grid&, int, const double*, size_t, double*, int, void*);

int main()
{
    using namespace mgmol_b200;
    int seen = 0;
    // (1) precondition of Grid (src/pb/Grid.cc:49-54: gdim divisible by the tasks)
    try
    {
        const unsigned gdim[3] = { 9, 8, 8 };
        const double ll[3]     = { 1., 1., 1. };
        const int bc[3] = { 1, 1, 1 }, nproc[3] = { 2, 1, 1 }, coord[3] = { 0, 0, 0 };
        Grid g(gdim, ll, 1, bc, nproc, coord);
        std::printf("no exception from Grid\n");
        return 1;
    }
    catch (const Error& e)
    {
        if (e.rc != MGB_EINVAL) return 2;
        seen++;
    }
    // (2) a failing library call: a null destination is MGB_EINVAL with or
    // without a device (or MGB_ENODEVICE, checked first, on a CPU box)
    try
    {
        MGB_CHECK(mgb_copy_to_dev(nullptr, nullptr, 8, nullptr));
        std::printf("no exception from MGB_CHECK\n");
        return 3;
    }
    catch (const Error& e)
    {
        if (e.rc != MGB_EINVAL && e.rc != MGB_ENODEVICE) return 4;
        if (!e.where || !*e.where) return 5;
        seen++;
    }
    // (3) the invalid Laplacian option of LapFactory (src/pb/LapFactory.h)
    try
    {
        const unsigned gdim[3] = { 8, 8, 8 };
        const double ll[3]     = { 1., 1., 1. };
        const int bc[3] = { 1, 1, 1 }, nproc[3] = { 1, 1, 1 }, coord[3] = { 0, 0, 0 };
        Grid g(gdim, ll, 1, bc, nproc, coord);
        Lap<double>* lap = LapFactory<double>::createLap(g, 7);
        delete lap;
        return 6;
    }
    catch (const Error& e)
    {
        if (e.rc != MGB_EINVAL) return 7;
        seen++;
    }
    std::printf("ok %d\n", seen);
    return seen == 3 ? 0 : 8;
}
