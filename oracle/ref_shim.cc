// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// extern "C" entry points around the UNMODIFIED reference sources (compiled
// in place from /root/reference by oracle/Makefile into oracle/_ref/).  The
// shim owns no numerics: every number it returns is produced by a reference
// function.  Where a reference class cannot be compiled here (Hamiltonian,
// BlockVector, the Orbitals classes and OrbitalsPreconditioning pull
// Control.h -> Boost), the shim replays that class's call sequence on the
// reference's own pb:: objects, citing the lines it replays.
#include <cstring>
#include <iostream>
#include <memory>
#include <vector>

#include <mpi.h>

#include "FDkernels.h"
#include "Grid.h"
#include "GridFunc.h"
#include "GridFuncVector.h"
#include "LapFactory.h"
#include "MGkernels.h"
#include "PEenv.h"
#include "Preconditioning.h"
#include "mputils.h"

using HostSpace = MemorySpace::Host;
using LAU       = LinearAlgebraUtils<MemorySpace::Host>;

namespace
{
struct Box
{
    pb::PEenv pe;
    pb::Grid grid;
    Box(const int dims[3], const double ll[3], int ghosts)
        : pe(MPI_COMM_WORLD, dims[0], dims[1], dims[2]),
          grid(zero3(), ll, udims(dims), pe, (short)ghosts, 0)
    {
    }
    static const double* zero3()
    {
        static const double o[3] = { 0., 0., 0. };
        return o;
    }
    static const unsigned* udims(const int d[3])
    {
        static thread_local unsigned u[3];
        u[0] = (unsigned)d[0];
        u[1] = (unsigned)d[1];
        u[2] = (unsigned)d[2];
        return u;
    }
};

std::vector<std::vector<int>> identity_gids(int nfunc)
{
    std::vector<std::vector<int>> gid(1);
    gid[0].resize(nfunc);
    for (int i = 0; i < nfunc; i++)
        gid[0][i] = i;
    return gid;
}

template <typename T>
void fdkernel(int kind, const int dims[3], int ghosts, const double ll[3],
    T* v, T* out, int nfunc, int rhs_ghosts)
{
    Box b(dims, ll, ghosts);
    HostSpace host;
    switch (kind)
    {
        case 0:
            pb::FDkernelDel2_4th_Mehr(b.grid, v, out, (size_t)nfunc, host);
            break;
        case 1:
            pb::FDkernelDel2_2nd(b.grid, v, out, (size_t)nfunc, host);
            break;
        case 2:
            pb::FDkernelDel2_4th(b.grid, v, out, (size_t)nfunc, host);
            break;
        case 3:
            pb::FDkernelDel2_6th(b.grid, v, out, (size_t)nfunc, host);
            break;
        case 4:
            pb::FDkernelDel2_8th(b.grid, v, out, (size_t)nfunc, host);
            break;
        case 100:
            pb::FDkernelRHS_4th_Mehr1(
                b.grid, v, out, (short)rhs_ghosts, (size_t)nfunc, host);
            break;
        default:
            std::cerr << "ref_shim: bad FD kernel kind " << kind << std::endl;
            abort();
    }
}

// BlockVector::setDataWithGhosts (src/BlockVector.cc:489-517): resetData,
// then per function GridFunc::assign of the no-ghost column.
template <typename T, typename T2>
void set_data_with_ghosts(
    pb::GridFuncVector<T, HostSpace>& gfv, const T2* phi, int nfunc, size_t npt)
{
    gfv.resetData();
    gfv.set_updated_boundaries(false);
    for (int i = 0; i < nfunc; i++)
        gfv.assign(i, phi + (size_t)i * npt);
}

// BlockVector::assign(GridFuncVector) (src/BlockVector.cc:303-311):
// per function GridFunc::getValues into the no-ghost column.
template <typename T, typename T2>
void strip_ghosts(
    pb::GridFuncVector<T, HostSpace>& gfv, T2* out, int nfunc, size_t npt)
{
    for (int i = 0; i < nfunc; i++)
        gfv.getValues(i, out + (size_t)i * npt);
}

// Hamiltonian<T>::applyLocal(ncolors, phi, hphi), src/Hamiltonian.cc:85-159.
template <typename T>
void hpsi(int lap_type, const int dims[3], const double ll[3], const int bc[3],
    const T* phi, const double* vtot, T* hphi, int nfunc)
{
    const bool mehrstellen = (lap_type == 0 || lap_type == 10);
    // src/GridFactory.h:23-51: Mehrstellen 1 ghost, 4th order 2 ghosts
    const int ghosts = mehrstellen ? 1 : 2;
    Box b(dims, ll, ghosts);
    const size_t npt = b.grid.size();
    const std::vector<std::vector<int>> gid = identity_gids(nfunc);

    // phi.setDataWithGhosts(); phi.trade_boundaries();   (:101-102)
    pb::GridFuncVector<T, HostSpace> gfvphi(b.grid, bc[0], bc[1], bc[2], gid);
    set_data_with_ghosts(gfvphi, phi, nfunc, npt);
    gfvphi.trade_boundaries();

    if (mehrstellen)
    {
        // :108-111
        pb::GridFunc<double> gfpot(b.grid, bc[0], bc[1], bc[2]);
        gfpot.assign(vtot);
        gfpot.trade_boundaries();
        // :113-117
        pb::GridFuncVector<T, HostSpace> gfvw1(
            b.grid, bc[0], bc[1], bc[2], gid);
        gfvw1.pointwiseProduct(gfvphi, gfpot);
        // :119-122  work1 = B*V*psi
        pb::GridFuncVector<T, HostSpace> gfv_work1(
            b.grid, bc[0], bc[1], bc[2], gid);
        gfvw1.applyRHS(0, gfv_work1);
        // :127  gfvw1 = -Lap*phi
        gfvphi.applyLap(0, gfvw1);
        // :129
        gfv_work1.axpy(1., gfvw1);
        // :131  hphi.setPsi(gfv_work1)
        strip_ghosts(gfv_work1, hphi, nfunc, npt);
    }
    else
    {
        // :138-155
        pb::Lap<T>* lap = LapFactory<T>::createLap(b.grid, lap_type);
        for (int i = 0; i < nfunc; i++)
            lap->applyWithPot(
                gfvphi.getGridFunc(i), vtot, hphi + (size_t)i * npt);
        delete lap;
    }
}
} // namespace

extern "C"
{

// dtype: 0 = float, 1 = double
void ref_fdkernel(int kind, int dtype, const int* dims, int ghosts,
    const double* ll, void* v, void* out, int nfunc, int rhs_ghosts)
{
    if (dtype == 1)
        fdkernel<double>(
            kind, dims, ghosts, ll, (double*)v, (double*)out, nfunc, rhs_ghosts);
    else
        fdkernel<float>(
            kind, dims, ghosts, ll, (float*)v, (float*)out, nfunc, rhs_ghosts);
}

void ref_hpsi(int lap_type, int dtype, const int* dims, const double* ll,
    const int* bc, const void* phi, const double* vtot, void* hphi, int nfunc)
{
    if (dtype == 1)
        hpsi<double>(lap_type, dims, ll, bc, (const double*)phi, vtot,
            (double*)hphi, nfunc);
    else
        hpsi<float>(lap_type, dims, ll, bc, (const float*)phi, vtot,
            (float*)hphi, nfunc);
}

// no-ghost block -> ghosted block with exchanged boundaries
// (BlockVector::setDataWithGhosts + GridFuncVector::trade_boundaries,
//  src/pb/GridFuncVector.cc:1544-1622)
void ref_trade_boundaries(int dtype, const int* dims, int ghosts,
    const double* ll, const int* bc, const void* noghost, void* ghosted,
    int nfunc)
{
    Box b(dims, ll, ghosts);
    const std::vector<std::vector<int>> gid = identity_gids(nfunc);
    const size_t npt                        = b.grid.size();
    const size_t ng                         = b.grid.sizeg();
    if (dtype == 1)
    {
        pb::GridFuncVector<double, HostSpace> g(
            b.grid, bc[0], bc[1], bc[2], gid);
        set_data_with_ghosts(g, (const double*)noghost, nfunc, npt);
        g.trade_boundaries();
        memcpy(ghosted, g.data(), sizeof(double) * ng * nfunc);
    }
    else
    {
        pb::GridFuncVector<float, HostSpace> g(
            b.grid, bc[0], bc[1], bc[2], gid);
        set_data_with_ghosts(g, (const float*)noghost, nfunc, npt);
        g.trade_boundaries();
        memcpy(ghosted, g.data(), sizeof(float) * ng * nfunc);
    }
}

// Operator constants as the reference's Lap classes report them
// (diagEl / invDiagEl / jacobiFactor; src/pb/Laph4M.h:80, Laph4.h:106,
//  Laph2.h:95).  out[0..2] = diagEl, invDiagEl, jacobiFactor.
void ref_lap_constants(
    int lap_type, const int* dims, int ghosts, const double* ll, double* out)
{
    Box b(dims, ll, ghosts);
    pb::Lap<float>* lap = LapFactory<float>::createLap(b.grid, lap_type);
    out[0]              = lap->diagEl();
    out[1]              = lap->invDiagEl();
    out[2]              = lap->jacobiFactor();
    delete lap;
}

// One GridFuncVector::jacobi sweep (src/pb/GridFuncVector.cc:2416-2425) on
// ghosted blocks v (in/out), f, w (out).
void ref_jacobi(int lap_type, const int* dims, int ghosts, const double* ll,
    const int* bc, float* v, const float* f, float* w, int nfunc,
    double jacobi_factor)
{
    Box b(dims, ll, ghosts);
    const std::vector<std::vector<int>> gid = identity_gids(nfunc);
    const size_t n                          = b.grid.sizeg() * (size_t)nfunc;
    pb::GridFuncVector<float, HostSpace> gv(b.grid, bc[0], bc[1], bc[2], gid);
    pb::GridFuncVector<float, HostSpace> gf(b.grid, bc[0], bc[1], bc[2], gid);
    pb::GridFuncVector<float, HostSpace> gw(b.grid, bc[0], bc[1], bc[2], gid);
    memcpy(gv.data(), v, n * sizeof(float));
    memcpy(gf.data(), f, n * sizeof(float));
    memcpy(gw.data(), w, n * sizeof(float));
    gv.set_updated_boundaries(false);
    gv.jacobi(lap_type, gf, gw, jacobi_factor);
    memcpy(v, gv.data(), n * sizeof(float));
    memcpy(w, gw.data(), n * sizeof(float));
}

// MGkernelRestrict3D / MGkernelExtend3D through GridFuncVector
// (src/pb/GridFuncVector.cc:1624-1641): fine/coarse are ghosted blocks; the
// source block's boundaries are traded first, as the reference does.
void ref_restrict3D(int dtype, const int* dims, int ghosts, const double* ll,
    const int* bc, const void* fine, void* coarse, int nfunc)
{
    Box b(dims, ll, ghosts);
    pb::Grid cgrid(b.grid.coarse_grid());
    const std::vector<std::vector<int>> gid = identity_gids(nfunc);
    if (dtype == 1)
    {
        pb::GridFuncVector<double, HostSpace> gf(
            b.grid, bc[0], bc[1], bc[2], gid);
        pb::GridFuncVector<double, HostSpace> gc(
            cgrid, bc[0], bc[1], bc[2], gid);
        memcpy(gf.data(), fine, sizeof(double) * b.grid.sizeg() * nfunc);
        gf.set_updated_boundaries(false);
        gf.restrict3D(gc);
        memcpy(coarse, gc.data(), sizeof(double) * cgrid.sizeg() * nfunc);
    }
    else
    {
        pb::GridFuncVector<float, HostSpace> gf(
            b.grid, bc[0], bc[1], bc[2], gid);
        pb::GridFuncVector<float, HostSpace> gc(
            cgrid, bc[0], bc[1], bc[2], gid);
        memcpy(gf.data(), fine, sizeof(float) * b.grid.sizeg() * nfunc);
        gf.set_updated_boundaries(false);
        gf.restrict3D(gc);
        memcpy(coarse, gc.data(), sizeof(float) * cgrid.sizeg() * nfunc);
    }
}

void ref_extend3D(int dtype, const int* dims, int ghosts, const double* ll,
    const int* bc, const void* coarse, void* fine, int nfunc)
{
    Box b(dims, ll, ghosts);
    pb::Grid cgrid(b.grid.coarse_grid());
    const std::vector<std::vector<int>> gid = identity_gids(nfunc);
    if (dtype == 1)
    {
        pb::GridFuncVector<double, HostSpace> gf(
            b.grid, bc[0], bc[1], bc[2], gid);
        pb::GridFuncVector<double, HostSpace> gc(
            cgrid, bc[0], bc[1], bc[2], gid);
        memcpy(gc.data(), coarse, sizeof(double) * cgrid.sizeg() * nfunc);
        memcpy(gf.data(), fine, sizeof(double) * b.grid.sizeg() * nfunc);
        gc.set_updated_boundaries(false);
        gf.extend3D(gc);
        memcpy(fine, gf.data(), sizeof(double) * b.grid.sizeg() * nfunc);
    }
    else
    {
        pb::GridFuncVector<float, HostSpace> gf(
            b.grid, bc[0], bc[1], bc[2], gid);
        pb::GridFuncVector<float, HostSpace> gc(
            cgrid, bc[0], bc[1], bc[2], gid);
        memcpy(gc.data(), coarse, sizeof(float) * cgrid.sizeg() * nfunc);
        memcpy(gf.data(), fine, sizeof(float) * b.grid.sizeg() * nfunc);
        gc.set_updated_boundaries(false);
        gf.extend3D(gc);
        memcpy(fine, gf.data(), sizeof(float) * b.grid.sizeg() * nfunc);
    }
}

// OrbitalsPreconditioning<T>::precond_mg (src/OrbitalsPreconditioning.cc:
// 87-117) without masks: res (no-ghost, ORBDTYPE) -> M^-1 res, through the
// reference's Preconditioning<float>::mg (src/Preconditioning.cc:155-216).
void ref_precond_mg(int lap_type, int mg_levels, int dtype, const int* dims,
    const double* ll, const int* bc, void* res, int nfunc, double gamma)
{
    const bool mehrstellen = (lap_type == 0 || lap_type == 10);
    const int ghosts       = mehrstellen ? 1 : 2;
    Box b(dims, ll, ghosts);
    const size_t npt                        = b.grid.size();
    const std::vector<std::vector<int>> gid = identity_gids(nfunc);
    const short sbc[3] = { (short)bc[0], (short)bc[1], (short)bc[2] };

    Preconditioning<float> precond(
        (short)lap_type, (short)mg_levels, b.grid, sbc);
    precond.setup(gid);

    pb::GridFuncVector<float, HostSpace> gfv_work(
        b.grid, bc[0], bc[1], bc[2], gid);
    pb::GridFuncVector<float, HostSpace> gfv_work2(
        b.grid, bc[0], bc[1], bc[2], gid);

    gfv_work.resetData();
    if (dtype == 1)
        set_data_with_ghosts(gfv_work2, (const double*)res, nfunc, npt);
    else
        set_data_with_ghosts(gfv_work2, (const float*)res, nfunc, npt);
    gfv_work.axpy(gamma, gfv_work2);

    precond.mg(gfv_work, gfv_work2, (short)lap_type, 0);

    if (dtype == 1)
        strip_ghosts(gfv_work, (double*)res, nfunc, npt);
    else
        strip_ghosts(gfv_work, (float*)res, nfunc, npt);
}

// mputils contractions (src/linear_algebra/mputils.cc).  Column-major, BLAS
// argument order.  dtype selects <double,double,double> (-> DGEMM/DSYRK) or
// the float-in / double-accumulate template paths.
// c_dtype: element type of C (0 float, 1 double).
void ref_mpgemm(int dtype, int c_dtype, char transa, char transb, int m, int n,
    int k, double alpha, const void* a, int lda, const void* b, int ldb,
    double beta, void* c, int ldc)
{
    if (dtype == 1 && c_dtype == 1)
        LAU::MPgemm(transa, transb, m, n, k, alpha, (const double*)a, lda,
            (const double*)b, ldb, beta, (double*)c, ldc);
    else if (dtype == 0 && c_dtype == 0)
        LAU::MPgemm(transa, transb, m, n, k, alpha, (const float*)a, lda,
            (const float*)b, ldb, beta, (float*)c, ldc);
    else if (dtype == 0 && c_dtype == 1)
        LAU::MPgemm(transa, transb, m, n, k, alpha, (const float*)a, lda,
            (const float*)b, ldb, beta, (double*)c, ldc);
    else
    {
        std::cerr << "ref_mpgemm: unsupported type combination" << std::endl;
        abort();
    }
}

// MPgemmNN with a double matrix B (ExtendedGridOrbitals::multiplyByMatrix,
// src/ExtendedGridOrbitals.cc:448-462: A = Phi (ORBDTYPE), B = MATDTYPE
// double, C = ORBDTYPE).
void ref_mpgemmNN_phiM(int dtype, int m, int n, int k, double alpha,
    const void* a, int lda, const double* b, int ldb, double beta, void* c,
    int ldc)
{
    if (dtype == 1)
        LAU::MPgemmNN(m, n, k, alpha, (const double*)a, lda, b, ldb, beta,
            (double*)c, ldc);
    else
        LAU::MPgemmNN(
            m, n, k, alpha, (const float*)a, lda, b, ldb, beta, (float*)c, ldc);
}

void ref_mpsyrk(int dtype, int c_dtype, char uplo, char trans, int n, int k,
    double alpha, const void* a, int lda, double beta, void* c, int ldc)
{
    if (dtype == 1 && c_dtype == 1)
        LAU::MPsyrk(uplo, trans, n, k, alpha, (const double*)a, lda, beta,
            (double*)c, ldc);
    else if (dtype == 0 && c_dtype == 0)
        LAU::MPsyrk(uplo, trans, n, k, alpha, (const float*)a, lda, beta,
            (float*)c, ldc);
    else if (dtype == 0 && c_dtype == 1)
        LAU::MPsyrk(uplo, trans, n, k, alpha, (const float*)a, lda, beta,
            (double*)c, ldc);
    else
    {
        std::cerr << "ref_mpsyrk: unsupported type combination" << std::endl;
        abort();
    }
}

double ref_mpdot(int dtype, int len, const void* x, const void* y)
{
    if (dtype == 1)
        return LAU::MPdot(len, (const double*)x, (const double*)y);
    return LAU::MPdot(len, (const float*)x, (const float*)y);
}

void ref_mpaxpy(int dtype, int len, double alpha, const void* x, void* y)
{
    if (dtype == 1)
        LAU::MPaxpy(len, alpha, (const double*)x, (double*)y);
    else
        LAU::MPaxpy(len, alpha, (const float*)x, (float*)y);
}

// pb::Lap<T>::rhs(GridFunc&, T*) per orbital, as MGmol::computeResidualUsingHPhi
// applies it (src/MGmol.cc:1252-1260): Laph4M -> FDoper::rhs_4th_Mehr1
// (src/pb/FDoper.cc:498-508), Laph4MP -> rhs_4th_Mehr2 (:573-637; asserts more
// than one ghost layer), others -> copy.  in / out: no-ghost blocks.
void ref_lap_rhs(int lap_type, int dtype, const int* dims, int ghosts, const double* ll,
    const int* bc, const void* in, void* out, int nfunc)
{
    Box b(dims, ll, ghosts);
    const size_t npt = b.grid.size();
    const std::vector<std::vector<int>> gid = identity_gids(nfunc);
    if (dtype == 1)
    {
        pb::GridFuncVector<double, HostSpace> gfv(b.grid, bc[0], bc[1], bc[2], gid);
        set_data_with_ghosts(gfv, (const double*)in, nfunc, npt);
        gfv.trade_boundaries();
        pb::Lap<double>* lap = LapFactory<double>::createLap(b.grid, lap_type);
        for (int i = 0; i < nfunc; i++)
            lap->rhs(gfv.getGridFunc(i), (double*)out + (size_t)i * npt);
        delete lap;
    }
    else
    {
        pb::GridFuncVector<float, HostSpace> gfv(b.grid, bc[0], bc[1], bc[2], gid);
        set_data_with_ghosts(gfv, (const float*)in, nfunc, npt);
        gfv.trade_boundaries();
        pb::Lap<float>* lap = LapFactory<float>::createLap(b.grid, lap_type);
        for (int i = 0; i < nfunc; i++)
            lap->rhs(gfv.getGridFunc(i), (float*)out + (size_t)i * npt);
        delete lap;
    }
}

int ref_shim_version() { return 2; }
}
