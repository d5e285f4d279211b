// mgmol_b200.hpp -- C++ host side of the B200 hot path, above the C ABI of
// mgmol_b200.h.  Header-only; needs only a C++11 compiler (no CUDA headers) and
// libmgmol_b200.so at link time.
//
// The classes mirror the reference's operator API for this path -- same names,
// same method names, same argument meaning, same error behaviour (the reference
// aborts: src/pb/Lap.h:37-38, src/pb/FDkernels.h:42-45) -- so MGmol's driver
// code (src/computeHij.cc:404-455, src/MGmol.cc:1191-1291, src/ABPG.cc:73-140)
// reads the same against either implementation.  INTEGRATION.md shows where
// each one slots into the MGmol tree.
//
//   reference class (file)                          here
//   MemorySpace::Memory<T,Device> (tools/memory_space.h:253-333)  DeviceMemory<T>
//   pb::Grid + pb::PEenv (pb/Grid.h:24-121)                       Grid
//   pb::GridFuncVector<T,Device> (pb/GridFuncVector.h:30-493)     GridFuncVector<T>
//   pb::Lap<T>, LapFactory<T> (pb/Lap.h:19-54, LapFactory.h:26-56) Lap<T>, LapFactory<T>
//   Potentials (Potentials.h:141: vtot, iterative index)          Potentials
//   ExtendedGridOrbitals (ExtendedGridOrbitals.h:43-404)          ExtendedGridOrbitals<T>
//   Hamiltonian<T> (Hamiltonian.h:20-52)                          Hamiltonian<T>
//   OrbitalsPreconditioning<T> (OrbitalsPreconditioning.h:27-70)  OrbitalsPreconditioning<T>
//   pb::PEenv / MGmol_MPI, the decomposed path (pb/PEenv.h:34-203)  Communicator
//   KBPsiMatrixSparse + get_vnlpsi (KBPsiMatrixSparse.cc:136-212,  KBProjectors<T>
//     get_vnlpsi.cc:24-87, computeHij.cc:294-375)
#ifndef MGMOL_B200_HPP
#define MGMOL_B200_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "mgmol_b200.h"

namespace mgmol_b200
{

// The reference has no error codes on this path: a failed precondition ends
// the run.  Same here, with the library's message.  A host that must survive a
// failure (the library's own C entry over these templates, csrc/poisson.cu)
// defines MGMOL_B200_ERRORS_THROW before including this header: every failed
// call then throws mgmol_b200::Error carrying the MGB_E* code instead.
#ifdef MGMOL_B200_ERRORS_THROW
struct Error
{
    int rc;
    const char* where;
};
inline void check(int rc, const char* where)
{
    if (rc != MGB_OK) throw Error{ rc, where };
}
[[noreturn]] inline void fail(const char* /*message already printed*/)
{
    throw Error{ MGB_EINVAL, "precondition" };
}
#else
inline void check(int rc, const char* where)
{
    if (rc != MGB_OK)
    {
        std::fprintf(stderr, "mgmol_b200: %s failed (%d): %s\n", where, rc, mgb_last_error());
        std::abort();
    }
}
[[noreturn]] inline void fail(const char*) { std::abort(); }
#endif
#define MGB_CHECK(call) ::mgmol_b200::check((call), #call)

template <typename T>
struct dtype_of;
template <>
struct dtype_of<float>
{
    static const int value = MGB_F32;
};
template <>
struct dtype_of<double>
{
    static const int value = MGB_F64;
};

// MemorySpace::Memory<T, MemorySpace::Device>: allocate / free / copy / set.
template <typename T>
class DeviceMemory
{
public:
    DeviceMemory() : ptr_(nullptr), size_(0) {}
    explicit DeviceMemory(size_t n) : ptr_(nullptr), size_(0) { allocate(n); }
    ~DeviceMemory() { free(); }
    DeviceMemory(const DeviceMemory&) = delete;
    DeviceMemory& operator=(const DeviceMemory&) = delete;
    void allocate(size_t n)
    {
        free();
        void* p = nullptr;
        MGB_CHECK(mgb_malloc(&p, n * sizeof(T)));
        ptr_  = static_cast<T*>(p);
        size_ = n;
    }
    void free()
    {
        // runs in destructors (possibly while an Error unwinds): never throws
        if (ptr_ && mgb_free(ptr_) != MGB_OK)
            std::fprintf(stderr, "mgmol_b200: mgb_free failed: %s\n", mgb_last_error());
        ptr_  = nullptr;
        size_ = 0;
    }
    void set(int value, void* stream = nullptr)
    {
        MGB_CHECK(mgb_memset(ptr_, value, size_ * sizeof(T), stream));
    }
    void copy_to_dev(const T* host, size_t n, void* stream = nullptr)
    {
        assert(n <= size_);
        MGB_CHECK(mgb_copy_to_dev(ptr_, host, n * sizeof(T), stream));
    }
    void copy_to_host(T* host, size_t n, void* stream = nullptr) const
    {
        assert(n <= size_);
        MGB_CHECK(mgb_copy_to_host(host, ptr_, n * sizeof(T), stream));
        MGB_CHECK(mgb_stream_sync(stream));
    }
    T* data() { return ptr_; }
    const T* data() const { return ptr_; }
    size_t size() const { return size_; }

private:
    T* ptr_;
    size_t size_;
};

// GramMatrix::computeLoewdinTransform (src/GramMatrix.cc:267-285) for a
// replicated n x n matrix on the host: S = V diag(l) V^T, P = V diag(1/sqrt(l)) V^T
// and, when asked for, sqrtS = V diag(sqrt(l)) V^T (the reference's invLoewdin).
// The reference diagonalises with dsyev; this header needs no LAPACK, so cyclic
// Jacobi rotations do it (n is the number of orbitals; O(n^3) per sweep, a
// handful of sweeps).  A caller that links LAPACK can pass its own P to
// ExtendedGridOrbitals::orthonormalizeLoewdin instead.  Matrices column-major,
// symmetric.  Returns false when S is not positive definite.
inline bool loewdinTransform(const int n, const double* S, double* P, double* sqrtS = nullptr)
{
    const size_t N = (size_t)n;
    std::vector<double> A(S, S + N * N), V(N * N, 0.);
    for (size_t i = 0; i < N; i++)
        V[i + i * N] = 1.;
    for (int sweep = 0; sweep < 60; sweep++)
    {
        double off = 0., diag = 0.;
        for (size_t q = 0; q < N; q++)
        {
            diag += A[q + q * N] * A[q + q * N];
            for (size_t p = 0; p < q; p++)
                off += A[p + q * N] * A[p + q * N];
        }
        if (off <= 1e-32 * diag) break;
        for (size_t q = 1; q < N; q++)
            for (size_t p = 0; p < q; p++)
            {
                const double apq = A[p + q * N];
                if (apq == 0.) continue;
                const double theta = (A[q + q * N] - A[p + p * N]) / (2. * apq);
                const double t
                    = (theta >= 0. ? 1. : -1.) / (std::fabs(theta) + std::sqrt(theta * theta + 1.));
                const double c = 1. / std::sqrt(t * t + 1.), sn = t * c;
                for (size_t k = 0; k < N; k++) // columns p, q
                {
                    const double akp = A[k + p * N], akq = A[k + q * N];
                    A[k + p * N] = c * akp - sn * akq;
                    A[k + q * N] = sn * akp + c * akq;
                }
                for (size_t k = 0; k < N; k++) // rows p, q
                {
                    const double apk = A[p + k * N], aqk = A[q + k * N];
                    A[p + k * N] = c * apk - sn * aqk;
                    A[q + k * N] = sn * apk + c * aqk;
                }
                for (size_t k = 0; k < N; k++)
                {
                    const double vkp = V[k + p * N], vkq = V[k + q * N];
                    V[k + p * N] = c * vkp - sn * vkq;
                    V[k + q * N] = sn * vkp + c * vkq;
                }
            }
    }
    // The accumulated rotations drift from orthogonality by ~n eps; the
    // reference's own unit test asks B^-1/2 B^1/2 = 1 to 1e-14
    // (tests/testGramMatrix.cc:92-108).  One Newton-Schulz step
    // V <- V (3 - V^T V) / 2 squares that error; the eigenvalues are then the
    // Rayleigh quotients of the polished vectors with the original matrix.
    {
        std::vector<double> W(N * N), V2(N * N);
        for (size_t j = 0; j < N; j++)
            for (size_t i = 0; i < N; i++)
            {
                double w = 0.;
                for (size_t k = 0; k < N; k++)
                    w += V[k + i * N] * V[k + j * N];
                W[i + j * N] = (i == j ? 1.5 : 0.) - 0.5 * w;
            }
        for (size_t j = 0; j < N; j++)
            for (size_t i = 0; i < N; i++)
            {
                double v = 0.;
                for (size_t k = 0; k < N; k++)
                    v += V[i + k * N] * W[k + j * N];
                V2[i + j * N] = v;
            }
        V.swap(V2);
        for (size_t k = 0; k < N; k++)
        {
            double lam = 0.;
            for (size_t j = 0; j < N; j++)
            {
                double sv = 0.;
                for (size_t i = 0; i < N; i++)
                    sv += S[i + j * N] * V[i + k * N];
                lam += sv * V[j + k * N];
            }
            A[k + k * N] = lam;
        }
    }
    for (size_t k = 0; k < N; k++)
        if (!(A[k + k * N] > 0.)) return false;
    for (size_t j = 0; j < N; j++)
        for (size_t i = 0; i < N; i++)
        {
            double a = 0., b = 0.;
            for (size_t k = 0; k < N; k++)
            {
                const double vv = V[i + k * N] * V[j + k * N];
                const double r  = std::sqrt(A[k + k * N]);
                a += vv / r;
                b += vv * r;
            }
            P[i + j * N] = a;
            if (sqrtS) sqrtS[i + j * N] = b;
        }
    return true;
}

// GridFactory (src/GridFactory.h:23-51): ghost width per operator
inline short ghostsFor(const int lap_type)
{
    switch (lap_type)
    {
        case MGB_LAP_4: return 2;
        case MGB_LAP_6: return 3;
        case MGB_LAP_8: return 4;
        default: return 1;
    }
}

// pb::Grid + the slice of pb::PEenv the path needs
class Grid
{
public:
    Grid(const unsigned gdim[3], const double lattice[3], const short nghosts,
        const int bc[3] = nullptr, const int nproc[3] = nullptr,
        const int coord[3] = nullptr)
    {
        for (int d = 0; d < 3; d++)
        {
            c_.nproc[d] = nproc ? nproc[d] : 1;
            c_.coord[d] = coord ? coord[d] : 0;
            c_.bc[d]    = bc ? bc[d] : 1;
            c_.gdim[d]  = (int)gdim[d];
            // Grid.cc:49-54: dim = gdim / n_mpi_task, exact
            if (gdim[d] % c_.nproc[d] != 0)
            {
                std::fprintf(stderr, "Grid: gdim[%d]=%u not divisible by %d tasks\n", d,
                    gdim[d], c_.nproc[d]);
                ::mgmol_b200::fail("precondition");
            }
            c_.dim[d] = (int)gdim[d] / c_.nproc[d];
            ll_[d]    = lattice[d];
            c_.h[d]   = lattice[d] / gdim[d];
        }
        c_.ghosts = nghosts;
    }
    int dim(const short i) const { return c_.dim[i]; }
    int gdim(const short i) const { return c_.gdim[i]; }
    short ghost_pt() const { return (short)c_.ghosts; }
    double hgrid(const short i) const { return c_.h[i]; }
    double vel() const { return c_.h[0] * c_.h[1] * c_.h[2]; }
    size_t size() const { return (size_t)c_.dim[0] * c_.dim[1] * c_.dim[2]; }
    size_t sizeg() const
    {
        const size_t g = 2 * c_.ghosts;
        return (c_.dim[0] + g) * (c_.dim[1] + g) * (c_.dim[2] + g);
    }
    // Grid::coarse_grid (src/pb/Grid.cc:214-231): half the points, same ghosts
    Grid coarse_grid() const
    {
        unsigned gd[3] = { (unsigned)c_.gdim[0] / 2, (unsigned)c_.gdim[1] / 2,
            (unsigned)c_.gdim[2] / 2 };
        return Grid(gd, ll_, (short)c_.ghosts, c_.bc, c_.nproc, c_.coord);
    }
    Grid with_ghosts(const short g) const
    {
        Grid r(*this);
        r.c_.ghosts = g;
        return r;
    }
    const mgb_grid* c() const { return &c_; }

private:
    mgb_grid c_;
    double ll_[3];
};

// pb::GridFuncVector<T, MemorySpace::Device>: nfunc ghosted functions in one
// device allocation, with the updated_boundaries_ bookkeeping of the reference.
template <typename T>
class GridFuncVector
{
public:
    GridFuncVector(const Grid& grid, const int nfunc)
        : grid_(grid), nfunc_(nfunc), mem_(grid.sizeg() * (size_t)nfunc),
          updated_boundaries_(false)
    {
        resetData();
    }
    const Grid& grid() const { return grid_; }
    int size() const { return nfunc_; }
    T* data() { return mem_.data(); }
    void resetData(void* stream = nullptr)
    {
        mem_.set(0, stream);
        updated_boundaries_ = true;
    }
    void set_updated_boundaries(const bool f) { updated_boundaries_ = f; }
    bool updated_boundaries() const { return updated_boundaries_; }
    // GridFunc::scal = MPscal over every stored value (src/pb/GridFunc.cc:516-521)
    void scal(const double alpha, void* stream = nullptr)
    {
        MGB_CHECK(mgb_scal(
            dtype_of<T>::value, grid_.sizeg() * (size_t)nfunc_, alpha, mem_.data(), stream));
    }
    // BlockVector::setDataWithGhosts (src/BlockVector.cc:489-517)
    template <typename T2>
    void assign(const T2* noghost_dev, const size_t ld, void* stream = nullptr)
    {
        MGB_CHECK(mgb_gfv_set_with_ghosts(dtype_of<T2>::value, dtype_of<T>::value, grid_.c(),
            noghost_dev, ld, mem_.data(), nfunc_, stream));
        updated_boundaries_ = false;
    }
    // BlockVector::assign(GridFuncVector) (src/BlockVector.cc:303-311)
    template <typename T2>
    void getValues(T2* noghost_dev, const size_t ld, void* stream = nullptr)
    {
        MGB_CHECK(mgb_gfv_get_values(dtype_of<T>::value, dtype_of<T2>::value, grid_.c(),
            mem_.data(), noghost_dev, ld, nfunc_, stream));
    }
    // src/pb/GridFuncVector.cc:1544-1622 (directions owned by one rank)
    void trade_boundaries(void* stream = nullptr)
    {
        if (updated_boundaries_) return;
        MGB_CHECK(mgb_gfv_trade_boundaries(
            dtype_of<T>::value, grid_.c(), mem_.data(), nfunc_, stream));
        updated_boundaries_ = true;
    }
    // src/pb/GridFuncVector.cc:2370-2397
    void applyLap(const int type, GridFuncVector<T>& rhs, void* stream = nullptr)
    {
        static const int kinds[5] = { MGB_FD_DEL2_4TH_MEHR, MGB_FD_DEL2_2ND, MGB_FD_DEL2_4TH,
            MGB_FD_DEL2_6TH, MGB_FD_DEL2_8TH };
        if (type < 0 || type > 4)
        {
            std::fprintf(stderr, "GridFuncVector::applyLap: option invalid: %d\n", type);
            ::mgmol_b200::fail("precondition");
        }
        trade_boundaries(stream);
        MGB_CHECK(mgb_fd_apply(kinds[type], dtype_of<T>::value, grid_.c(), mem_.data(),
            rhs.data(), nfunc_, 0, stream));
        rhs.set_updated_boundaries(false);
    }
    // src/pb/GridFuncVector.cc:2400-2413
    void applyRHS(const int type, GridFuncVector<T>& rhs, void* stream = nullptr)
    {
        if (type == 0)
        {
            trade_boundaries(stream);
            MGB_CHECK(mgb_fd_apply(MGB_FD_RHS_4TH_MEHR1, dtype_of<T>::value, grid_.c(),
                mem_.data(), rhs.data(), nfunc_, grid_.ghost_pt(), stream));
            rhs.set_updated_boundaries(false);
        }
        else // as the reference has it: a memcpy of rhs INTO this (:2408-2411)
            MGB_CHECK(mgb_copy_dev(
                mem_.data(), rhs.data(), grid_.sizeg() * nfunc_ * sizeof(T), stream));
    }
    // src/pb/GridFuncVector.cc:90-136; V: ghosted double field on the device
    void pointwiseProduct(GridFuncVector<T>& A, const double* Vghost, void* stream = nullptr)
    {
        MGB_CHECK(mgb_gfv_pointwise_product(dtype_of<T>::value, grid_.c(), A.data(), Vghost,
            mem_.data(), nfunc_, stream));
        updated_boundaries_ = A.updated_boundaries_;
    }
    // src/pb/GridFuncVector.cc:1645-1668
    void axpy(const double alpha, GridFuncVector<T>& x, void* stream = nullptr)
    {
        MGB_CHECK(mgb_axpy(dtype_of<T>::value, grid_.sizeg() * (size_t)nfunc_, alpha, x.data(),
            mem_.data(), stream));
        updated_boundaries_ = updated_boundaries_ && x.updated_boundaries_;
    }
    GridFuncVector<T>& operator-=(GridFuncVector<T>& x)
    {
        axpy(-1., x);
        return *this;
    }
    // src/pb/GridFuncVector.cc:1624-1641
    void restrict3D(GridFuncVector<T>& coarse, void* stream = nullptr)
    {
        trade_boundaries(stream);
        MGB_CHECK(mgb_gfv_restrict3D(
            dtype_of<T>::value, grid_.c(), mem_.data(), coarse.data(), nfunc_, stream));
        coarse.set_updated_boundaries(false);
    }
    void extend3D(GridFuncVector<T>& coarse, void* stream = nullptr)
    {
        coarse.trade_boundaries(stream);
        MGB_CHECK(mgb_gfv_extend3D(
            dtype_of<T>::value, grid_.c(), coarse.data(), mem_.data(), nfunc_, stream));
        updated_boundaries_ = false;
    }

private:
    Grid grid_;
    int nfunc_;
    DeviceMemory<T> mem_;
    bool updated_boundaries_;
};

// Potentials: the slice the path reads -- vtot (POTDTYPE double, no ghosts,
// src/Potentials.h:141) on the device and its iterative index
class Potentials
{
public:
    explicit Potentials(const size_t npt) : vtot_(npt), itindex_(0) {}
    void setVtot(const double* host, void* stream = nullptr)
    {
        vtot_.copy_to_dev(host, vtot_.size(), stream);
        itindex_++;
    }
    const double* vtot() const { return vtot_.data(); }
    double* vtot() { return vtot_.data(); }
    int getIterativeIndex() const { return itindex_; }
    void incrementIterativeIndex() { itindex_++; }

private:
    DeviceMemory<double> vtot_;
    int itindex_;
};

// ExtendedGridOrbitals reduced to the hot path: psi is the no-ghost
// column-major numpt x numst block (lda = numpt, src/BlockVector.cc:79) in
// device memory; the iterative index keys Hamiltonian's cache
// (src/Orbitals.h:43-70).
template <typename T>
class ExtendedGridOrbitals
{
public:
    ExtendedGridOrbitals(const Grid& grid, const int numst)
        : grid_(grid), numst_(numst), numpt_(grid.size()), lda_(grid.size()),
          psi_(grid.size() * (size_t)numst), iterative_index_(0)
    {
        psi_.set(0);
    }
    int numst() const { return numst_; }
    int chromatic_number() const { return numst_; }
    size_t getNumpt() const { return numpt_; }
    size_t getLda() const { return lda_; }
    const Grid& grid() const { return grid_; }
    T* getPsi(const int i = 0) { return psi_.data() + (size_t)i * lda_; }
    const T* getPsi(const int i = 0) const { return psi_.data() + (size_t)i * lda_; }
    int getIterativeIndex() const { return iterative_index_; }
    void incrementIterativeIndex() { iterative_index_++; }
    // host <-> device of the whole block (BlockVector storage)
    void setPsi(const T* host, void* stream = nullptr)
    {
        psi_.copy_to_dev(host, psi_.size(), stream);
        incrementIterativeIndex();
    }
    void getPsiHost(T* host, void* stream = nullptr) const
    {
        psi_.copy_to_host(host, psi_.size(), stream);
    }
    // src/ExtendedGridOrbitals.cc:202-212 and BlockVector.h:112-126
    void axpy(const double alpha, const ExtendedGridOrbitals<T>& x, void* stream = nullptr)
    {
        MGB_CHECK(mgb_axpy(dtype_of<T>::value, psi_.size(), alpha, x.getPsi(), getPsi(), stream));
        incrementIterativeIndex();
    }
    void scal(const double alpha, void* stream = nullptr)
    {
        MGB_CHECK(mgb_scal(dtype_of<T>::value, psi_.size(), alpha, getPsi(), stream));
        incrementIterativeIndex();
    }
    // Orbitals::assign and operator-= (src/BlockVector.cc:289-299, 303-311)
    void assign(const ExtendedGridOrbitals<T>& x, void* stream = nullptr)
    {
        MGB_CHECK(mgb_copy_dev(getPsi(), x.getPsi(), psi_.size() * sizeof(T), stream));
        incrementIterativeIndex();
    }
    ExtendedGridOrbitals<T>& operator-=(const ExtendedGridOrbitals<T>& x)
    {
        MGB_CHECK(mgb_axpy(dtype_of<T>::value, psi_.size(), -1., x.getPsi(), getPsi(), nullptr));
        incrementIterativeIndex();
        return *this;
    }
    // computeDiagonalElementsDotProduct (src/ExtendedGridOrbitals.cc:1085-1106):
    // ss_dev[i] = vel <phi_i, psi_i>, one launch for all orbitals
    void computeDiagonalElementsDotProduct(const ExtendedGridOrbitals<T>& x, double* ss_dev,
        mgb_comm* comm = nullptr, void* stream = nullptr) const
    {
        MGB_CHECK(mgb_dot_cols(dtype_of<T>::value, numpt_, numst_, grid_.vel(), getPsi(), lda_,
            x.getPsi(), x.getLda(), ss_dev, stream));
        if (comm) MGB_CHECK(mgb_allreduce_sum_f64(comm, ss_dev, (size_t)numst_, stream));
    }
    // dotProductDiagonal (src/ExtendedGridOrbitals.cc:1205-1213) with the
    // diagonal of S^-1 (host array of numst weights, null = ones)
    double dotProduct(const ExtendedGridOrbitals<T>& x, const double* inv_s_diag = nullptr,
        mgb_comm* comm = nullptr)
    {
        if (dots_.size() < (size_t)numst_) dots_.allocate((size_t)numst_);
        computeDiagonalElementsDotProduct(x, dots_.data(), comm, nullptr);
        std::vector<double> ss((size_t)numst_);
        dots_.copy_to_host(ss.data(), ss.size());
        MGB_CHECK(mgb_stream_sync(nullptr));
        double s = 0.;
        for (int i = 0; i < numst_; i++)
            s += (inv_s_diag ? inv_s_diag[i] : 1.) * ss[i];
        return s;
    }
    // computeGram / getLocalOverlap (src/ExtendedGridOrbitals.cc:985-1010,
    // 1138-1162): ss_dev(numst x numst, column-major double) = vel Phi^T Phi
    void computeGram(double* ss_dev, mgb_comm* comm = nullptr, void* stream = nullptr) const
    {
        MGB_CHECK(mgb_syrk_t(dtype_of<T>::value, numst_, numpt_, grid_.vel(), getPsi(), lda_,
            ss_dev, numst_, stream));
        if (comm)
            MGB_CHECK(mgb_allreduce_sum_f64(comm, ss_dev, (size_t)numst_ * numst_, stream));
    }
    // computeLocalProduct / addDotWithNcol2Matrix
    // (src/ExtendedGridOrbitals.cc:1049-1083, 1704-1752): vel Phi^T A
    void computeLocalProduct(const ExtendedGridOrbitals<T>& A, double* ss_dev,
        mgb_comm* comm = nullptr, void* stream = nullptr) const
    {
        MGB_CHECK(mgb_gemm_tn(dtype_of<T>::value, numst_, A.numst(), numpt_, grid_.vel(),
            getPsi(), lda_, A.getPsi(), A.getLda(), 0., ss_dev, numst_, stream));
        if (comm)
            MGB_CHECK(mgb_allreduce_sum_f64(comm, ss_dev, (size_t)numst_ * A.numst(), stream));
    }
    // addDotWithNcol2Matrix (src/ExtendedGridOrbitals.cc:1704-1752): mat_dev
    // (numst x numst, column-major double) += vel Phi^T A summed over the ranks;
    // work_dev: numst x numst doubles
    void addDotWithNcol2Matrix(const ExtendedGridOrbitals<T>& A, double* mat_dev, double* work_dev,
        mgb_comm* comm = nullptr, void* stream = nullptr) const
    {
        computeLocalProduct(A, work_dev, comm, stream);
        MGB_CHECK(mgb_axpy(MGB_F64, (size_t)numst_ * A.numst(), 1., work_dev, mat_dev, stream));
    }
    // computeMatB (src/ExtendedGridOrbitals.cc:901-967): matB_dev(i, j) = vel <orbitals_i | B |
    // this_j>, B the operator's right-hand-side stencil (Lap::rhs; the identity unless
    // Mehrstellen), in blocks of `bcolor` columns through a work block like the reference
    // (bcolor = 32 there).  LapT: Lap<T> (defined below).
    template <class LapT>
    void computeMatB(const ExtendedGridOrbitals<T>& orbitals, const LapT& LapOper,
        double* matB_dev, mgb_comm* comm = nullptr, void* stream = nullptr,
        const int bcolor = 32) const
    {
        if (numst_ == 0) return;
        DeviceMemory<T> work;
        work.allocate(lda_ * (size_t)bcolor);
        for (int icolor = 0; icolor < numst_; icolor += bcolor)
        {
            const int nf = (icolor + bcolor > numst_) ? numst_ - icolor : bcolor;
            LapOper.rhs(getPsi(icolor), lda_, work.data(), lda_, nf, nullptr, stream);
            MGB_CHECK(mgb_gemm_tn(dtype_of<T>::value, orbitals.numst(), nf, numpt_, grid_.vel(),
                orbitals.getPsi(), orbitals.getLda(), work.data(), lda_, 0.,
                matB_dev + (size_t)icolor * orbitals.numst(), orbitals.numst(), stream));
        }
        if (comm)
            MGB_CHECK(mgb_allreduce_sum_f64(
                comm, matB_dev, (size_t)numst_ * orbitals.numst(), stream));
        MGB_CHECK(mgb_stream_sync(stream)); // work is freed on return
    }
    // multiplyByMatrix (src/ExtendedGridOrbitals.cc:448-498): product = Phi M,
    // M column-major numst x n on the device
    void multiplyByMatrix(const double* matrix_dev, const int n,
        ExtendedGridOrbitals<T>& product, void* stream = nullptr) const
    {
        MGB_CHECK(mgb_gemm_nn(dtype_of<T>::value, numpt_, n, numst_, 1., getPsi(), lda_,
            matrix_dev, numst_, 0., product.getPsi(), product.getLda(), stream));
        product.incrementIterativeIndex();
    }
    // orthonormalizeLoewdin (src/ExtendedGridOrbitals.cc:1304-1358): Phi <- Phi P,
    // P = S^-1/2.  The Gram matrix and Phi P are the library's kernels; the
    // n x n transform is the reference's replicated-matrix layer
    // (ProjectedMatrices::computeLoewdinTransform), done on the host by
    // loewdinTransform unless the caller supplies P (column-major, host).  On
    // return matrixTransform, when given, holds P as the reference's argument of
    // that name does.
    void orthonormalizeLoewdin(double* matrixTransform = nullptr, mgb_comm* comm = nullptr,
        const bool transform_given = false)
    {
        const size_t nn = (size_t)numst_ * numst_;
        std::vector<double> S(nn), P(nn);
        DeviceMemory<double> mat(nn);
        if (transform_given)
        {
            assert(matrixTransform);
            P.assign(matrixTransform, matrixTransform + nn);
        }
        else
        {
            computeGram(mat.data(), comm);
            mat.copy_to_host(S.data(), nn);
            if (!loewdinTransform(numst_, S.data(), P.data()))
            {
                std::fprintf(stderr,
                    "mgmol_b200: orthonormalizeLoewdin: Gram matrix not positive definite\n");
                ::mgmol_b200::fail("precondition");
            }
            if (matrixTransform) std::copy(P.begin(), P.end(), matrixTransform);
        }
        mat.copy_to_dev(P.data(), nn);
        DeviceMemory<T> product(psi_.size());
        MGB_CHECK(mgb_gemm_nn(dtype_of<T>::value, numpt_, numst_, numst_, 1., getPsi(), lda_,
            mat.data(), numst_, 0., product.data(), lda_, nullptr));
        MGB_CHECK(mgb_copy_dev(getPsi(), product.data(), psi_.size() * sizeof(T), nullptr));
        MGB_CHECK(mgb_stream_sync(nullptr));
        incrementIterativeIndex();
    }

protected:
    Grid grid_;
    int numst_;
    size_t numpt_, lda_;
    DeviceMemory<T> psi_;
    DeviceMemory<double> dots_;
    int iterative_index_;
};

// The localization masks one rank holds for its colors: the data of MasksSet /
// Map2Masks / GridMask (src/Map2Masks.cc:25-61, src/GridMask.h:41-52) as the
// path consumes it.  op: MGB_MASK_MULT (GridMaskMult) or MGB_MASK_MAX
// (GridMaskMax, the default orbital masks, src/MasksSet.cc:15,158-182).
class Masks
{
public:
    Masks(const Grid& grid, const short mg_levels, const short subdivx, const int ncolors,
        const int op)
        : handle_(nullptr)
    {
        MGB_CHECK(mgb_masks_create(&handle_, grid.c(), mg_levels, subdivx, ncolors, op));
    }
    ~Masks()
    {
        if (handle_) mgb_masks_destroy(handle_);
    }
    Masks(const Masks&)            = delete;
    Masks& operator=(const Masks&) = delete;
    // state: GridMask::mask_not_zero_ (-1 / 0 zero, 1 one, 2 values) of the
    // mask of overlapping_gids[iloc][color]; 0 for an empty slot (gid -1)
    template <typename M>
    void set(const short level, const short iloc, const int color, const short state,
        const M* values_host = nullptr)
    {
        MGB_CHECK(mgb_masks_set(
            handle_, level, iloc, color, state, dtype_of<M>::value, values_host));
    }
    void commit() { MGB_CHECK(mgb_masks_commit(handle_)); }
    // GridFuncVector::app_mask(level) (src/pb/GridFuncVector.cc:2428-2438)
    template <typename T>
    void apply(GridFuncVector<T>& gfv, const short level, void* stream = nullptr) const
    {
        MGB_CHECK(mgb_gfv_app_mask(dtype_of<T>::value, handle_, level, gfv.grid().ghost_pt(),
            gfv.data(), gfv.size(), stream));
    }
    const mgb_masks* handle() const { return handle_; }

private:
    mgb_masks* handle_;
};

// LocGridOrbitals reduced to the hot path (src/LocGridOrbitals.h:60-, .cc):
// numpt x chromatic_number colors; in x-slab iloc color c holds the global
// orbital overlapping_gids[iloc][c] or nothing (-1).  Stencils, V-cycle and
// BLAS-1 are those of the color block; contractions run per slab
// (LocalMatrices with subdivx sub-matrices), masks localize.
template <typename T>
class LocGridOrbitals : public ExtendedGridOrbitals<T>
{
public:
    LocGridOrbitals(const Grid& grid, const int numst,
        const std::vector<std::vector<int>>& overlapping_gids, const Masks* masks = nullptr)
        : ExtendedGridOrbitals<T>(grid, (int)overlapping_gids.at(0).size()),
          numst_global_(numst), overlapping_gids_(overlapping_gids),
          subdivx_((short)overlapping_gids.size()), masks_(masks)
    {
        assert(grid.dim(0) % subdivx_ == 0);
        loc_numpt_ = this->numpt_ / subdivx_;
    }
    short subdivx() const { return subdivx_; }
    size_t getLocNumpt() const { return loc_numpt_; }
    const std::vector<std::vector<int>>& getOverlappingGids() const
    {
        return overlapping_gids_;
    }
    // src/LocGridOrbitals.cc:427-452
    void applyMask(void* stream = nullptr)
    {
        if (masks_)
            MGB_CHECK(mgb_app_mask(dtype_of<T>::value, masks_->handle(), 0, this->getPsi(),
                this->lda_, this->chromatic_number(), stream));
        this->incrementIterativeIndex();
    }
    // getLocalOverlap (src/LocGridOrbitals.cc:1504-1530): ss_dev holds subdivx
    // column-major n x n matrices (LocalMatrices storage), all slabs in one
    // launch
    void getLocalOverlap(double* ss_dev, void* stream = nullptr) const
    {
        const int n = this->chromatic_number();
        MGB_CHECK(mgb_syrk_t_slabs(dtype_of<T>::value, n, loc_numpt_, subdivx_,
            this->grid_.vel(), this->getPsi(), this->lda_, ss_dev, n, stream));
    }
    // computeLocalProduct (src/LocGridOrbitals.cc:1554-1604)
    void computeLocalProduct(const ExtendedGridOrbitals<T>& A, double* ss_dev,
        const bool transpose = false, void* stream = nullptr) const
    {
        const T* a   = transpose ? A.getPsi() : this->getPsi();
        const T* b   = transpose ? this->getPsi() : A.getPsi();
        const size_t la = transpose ? A.getLda() : this->lda_;
        const size_t lb = transpose ? this->lda_ : A.getLda();
        const int m  = transpose ? A.chromatic_number() : this->chromatic_number();
        const int n  = transpose ? this->chromatic_number() : A.chromatic_number();
        MGB_CHECK(mgb_gemm_tn_slabs(dtype_of<T>::value, m, n, loc_numpt_, subdivx_,
            this->grid_.vel(), a, la, b, lb, 0., ss_dev, m, stream));
    }
    // matrixToLocalMatrix (src/LocGridOrbitals.cc:1400-1424), host side:
    // lmatrix (column-major chromatic x ncolor) from the global numst x numst
    // column-major matrix
    void matrixToLocalMatrix(const short iloc, const double* matrix, double* lmatrix) const
    {
        const int nc = this->chromatic_number();
        for (int j = 0; j < nc; j++)
        {
            const int gidj = overlapping_gids_[iloc][j];
            for (int i = 0; i < nc; i++)
            {
                const int gidi      = overlapping_gids_[iloc][i];
                lmatrix[j * nc + i] = (gidi != -1 && gidj != -1)
                                          ? matrix[(size_t)gidj * numst_global_ + gidi]
                                          : 0.;
            }
        }
    }
    // multiplyByMatrix (src/LocGridOrbitals.cc:793-825): per slab product =
    // Phi_iloc * M_iloc, matrices_dev = subdivx column-major n x n matrices
    void multiplyByMatrix(const double* matrices_dev, ExtendedGridOrbitals<T>& product,
        void* stream = nullptr) const
    {
        const int n = this->chromatic_number();
        for (short iloc = 0; iloc < subdivx_; iloc++)
            MGB_CHECK(mgb_gemm_nn(dtype_of<T>::value, loc_numpt_, n, n, 1.,
                this->getPsi() + iloc * loc_numpt_, this->lda_,
                matrices_dev + (size_t)iloc * n * n, n, 0., product.getPsi() + iloc * loc_numpt_,
                product.getLda(), stream));
        product.incrementIterativeIndex();
    }
    const Masks* masks() const { return masks_; }

private:
    int numst_global_;
    std::vector<std::vector<int>> overlapping_gids_;
    short subdivx_;
    size_t loc_numpt_;
    const Masks* masks_;
};

// The slice of pb::PEenv / MGmol_MPI the decomposed path needs (src/pb/PEenv.h:34-203,
// src/tools/MGmol_MPI.h): one process per GPU, the 3-D block decomposition carried by
// Grid (nproc, coord).  The 128-byte id is created on rank 0 (uniqueId) and handed to
// the other ranks by the caller -- MPI_Bcast inside MGmol.
class Communicator
{
public:
    static void uniqueId(unsigned char id128[128]) { MGB_CHECK(mgb_comm_unique_id(id128)); }
    Communicator(const unsigned char id128[128], const int rank, const int nranks)
        : handle_(nullptr), rank_(rank), nranks_(nranks)
    {
        MGB_CHECK(mgb_comm_create(&handle_, id128, rank, nranks));
    }
    ~Communicator()
    {
        if (handle_) mgb_comm_destroy(handle_);
    }
    Communicator(const Communicator&) = delete;
    Communicator& operator=(const Communicator&) = delete;
    mgb_comm* handle() const { return handle_; }
    int rank() const { return rank_; }
    int nranks() const { return nranks_; }
    // MGmol_MPI::allreduce(double*, n, MPI_SUM) on a device array
    void allreduce(double* dev, const size_t n, void* stream = nullptr)
    {
        MGB_CHECK(mgb_allreduce_sum_f64(handle_, dev, n, stream));
    }
    void barrier(void* stream = nullptr) { MGB_CHECK(mgb_comm_barrier(handle_, stream)); }
    // device sync + did a neighbour barrier of the in-place halo paths time out?
    void check() { MGB_CHECK(mgb_comm_check(handle_)); }
    // collective: publish this rank's resident array (its orbital block) so that the
    // neighbours' kernels read its boundary layers in place over NVLink
    void registerArray(const void* dev, void* stream = nullptr)
    {
        MGB_CHECK(mgb_peer_register(handle_, dev, stream));
    }
    void unregisterArray(const void* dev) { MGB_CHECK(mgb_peer_unregister(handle_, dev)); }
    // GridFuncVector::trade_boundaries on a px x py x pz decomposition: Y, Z, X faces
    // (src/pb/GridFuncVector.cc:1544-1622)
    template <typename T>
    void trade_boundaries(GridFuncVector<T>& gfv, void* stream = nullptr)
    {
        if (gfv.updated_boundaries()) return;
        MGB_CHECK(mgb_halo_exchange_ghosted(
            handle_, dtype_of<T>::value, gfv.grid().c(), gfv.data(), gfv.size(), stream));
        gfv.set_updated_boundaries(true);
    }
    // the packed x halo of a no-ghost block: xhalo[nfunc][2g][ny][nz]
    // (initiate/finishEastWestComm, src/pb/GridFuncVector.cc:1195-1256)
    template <typename T>
    void haloExchangeX(const Grid& grid, const int g, const T* noghost, const size_t ld, T* xhalo,
        const int nfunc, void* stream = nullptr)
    {
        MGB_CHECK(mgb_halo_exchange_x(
            handle_, dtype_of<T>::value, grid.c(), g, noghost, ld, xhalo, nfunc, stream));
    }

private:
    mgb_comm* handle_;
    int rank_, nranks_;
};

// pb::Lap<T> as LapFactory creates it
template <typename T>
class Lap
{
public:
    Lap(const Grid& grid, const int type) : grid_(grid), type_(type)
    {
        double h[3] = { grid.hgrid(0), grid.hgrid(1), grid.hgrid(2) }, out[3];
        MGB_CHECK(mgb_lap_constants(type, h, out));
        diagEl_ = out[0], invDiagEl_ = out[1], jacobiFactor_ = out[2];
    }
    double diagEl() const { return diagEl_; }
    double invDiagEl() const { return invDiagEl_; }
    double jacobiFactor() const { return jacobiFactor_; }
    short minNumberGhosts() const { return ghostsFor(type_); }
    int type() const { return type_; }
    // Lap<T>::apply on a block (src/pb/Laph4.h:99-102)
    void apply(GridFuncVector<T>& A, GridFuncVector<T>& B, void* stream = nullptr)
    {
        A.applyLap(type_ == MGB_LAP_4MP ? 0 : type_, B, stream);
    }
    // Lap<T>::applyWithPot (src/pb/Lap.h:35; only Laph4M/4MP/4 implement it,
    // the others abort) for the whole block = Hamiltonian::applyLocal body
    void applyWithPot(const T* phi, const size_t ld, const double* vtot, T* hphi,
        const size_t ldh, const int nfunc, const void* xhalo_phi = nullptr,
        const double* xhalo_v = nullptr, void* stream = nullptr)
    {
        MGB_CHECK(mgb_hpsi(type_, dtype_of<T>::value, grid_.c(), phi, ld, vtot, hphi, ldh,
            nfunc, xhalo_phi, xhalo_v, stream));
    }
    // applyWithPot on one rank of a decomposed domain: the neighbours' boundary planes,
    // rows, columns and Mehrstellen edge lines of phi are read in place from their
    // registered blocks (Communicator::registerArray) by the same fused kernel.
    // vghost: the potential as a ghosted double field with traded boundaries (gfpot of
    // src/Hamiltonian.cc:108-111; Hamiltonian::updatePotentialHalo); may be null for
    // MGB_LAP_4.  Collective over the communicator.
    void applyWithPotPeer(Communicator& comm, const T* phi, const size_t ld, const double* vtot,
        const double* vghost, T* hphi, const size_t ldh, const int nfunc, void* stream = nullptr)
    {
        MGB_CHECK(mgb_hpsi_peer3d(comm.handle(), type_, dtype_of<T>::value, grid_.c(), phi, ld,
            vtot, vghost, hphi, ldh, nfunc, stream));
    }
    // the x-slab form: V's halo as the 2g packed x planes (Communicator::haloExchangeX)
    void applyWithPotPeerX(Communicator& comm, const T* phi, const size_t ld, const double* vtot,
        const double* xhalo_v, T* hphi, const size_t ldh, const int nfunc, void* stream = nullptr)
    {
        MGB_CHECK(mgb_hpsi_peer(comm.handle(), type_, dtype_of<T>::value, grid_.c(), phi, ld, vtot,
            hphi, ldh, nfunc, xhalo_v, stream));
    }
    // Lap<T>::rhs (src/pb/Lap.h:32) for nfunc orbitals of a no-ghost block: B phi
    // (Mehrstellen, src/pb/FDkernels.cc:522-584) or a copy (B = 1)
    void rhs(const T* phi, const size_t ld, T* bphi, const size_t ldb, const int nfunc,
        const void* xhalo_phi = nullptr, void* stream = nullptr) const
    {
        MGB_CHECK(mgb_apply_b(type_, dtype_of<T>::value, grid_.c(), phi, ld, bphi, ldb, nfunc,
            xhalo_phi, stream));
    }

private:
    Grid grid_;
    int type_;
    double diagEl_, invDiagEl_, jacobiFactor_;
};

// src/LapFactory.h:26-56
template <typename T>
struct LapFactory
{
    static Lap<T>* createLap(const Grid& grid, const int lap_type)
    {
        switch (lap_type)
        {
            case 0:
            case 1:
            case 2:
            case 3:
            case 4:
            case 10:
                return new Lap<T>(grid, lap_type);
            default:
                std::fprintf(stderr, "LapFactory::createLap() --- option invalid:%d\n", lap_type);
                ::mgmol_b200::fail("precondition");
        }
        return nullptr;
    }
};

// Hamiltonian<OrbitalsType> (src/Hamiltonian.h:20-52, .cc:43-260)
template <typename T>
class Hamiltonian
{
public:
    Hamiltonian() : itindex_(-1) {}
    void setup(const Grid& myGrid, const int lap_type)
    {
        lapOper_.reset(LapFactory<T>::createLap(myGrid, lap_type));
        pot_.reset(new Potentials(myGrid.size()));
    }
    Lap<T>* lapOper() { return lapOper_.get(); }
    Potentials& potential() { return *pot_; }
    // src/Hamiltonian.cc:43-83: recompute only when 100*phi.index + pot.index
    // changed (or force)
    const ExtendedGridOrbitals<T>& applyLocal(
        ExtendedGridOrbitals<T>& phi, const bool force = false, void* stream = nullptr)
    {
        if (!hlphi_ || hlphi_->numst() != phi.numst())
        {
            hlphi_.reset(new ExtendedGridOrbitals<T>(phi.grid(), phi.numst()));
            itindex_ = -1;
        }
        const int new_index = 100 * phi.getIterativeIndex() + pot_->getIterativeIndex();
        if (force || new_index != itindex_)
        {
            applyLocal(phi.chromatic_number(), phi, *hlphi_, stream);
            itindex_ = new_index;
        }
        return *hlphi_;
    }
    // src/Hamiltonian.cc:85-159: one fused pass
    void applyLocal(const int ncolors, ExtendedGridOrbitals<T>& phi,
        ExtendedGridOrbitals<T>& hphi, void* stream = nullptr)
    {
        lapOper_->applyWithPot(phi.getPsi(), phi.getLda(), pot_->vtot(), hphi.getPsi(),
            hphi.getLda(), ncolors, nullptr, nullptr, stream);
        hphi.incrementIterativeIndex();
    }
    // Decomposed domains.  Once per potential update: the ghosted copy of vtot with
    // boundaries traded over the ranks (gfpot, src/Hamiltonian.cc:108-111).
    void updatePotentialHalo(const Grid& myGrid, Communicator& comm, void* stream = nullptr)
    {
        const Grid gg = myGrid.with_ghosts(lapOper_->minNumberGhosts());
        vghost_.reset(new GridFuncVector<double>(gg, 1));
        vghost_->assign(pot_->vtot(), gg.size(), stream);
        comm.trade_boundaries(*vghost_, stream);
    }
    // applyLocal with phi's halo read in place from the neighbour ranks: phi's block must
    // be registered (comm.registerArray(phi.getPsi())); same cache key as above
    const ExtendedGridOrbitals<T>& applyLocal(ExtendedGridOrbitals<T>& phi, Communicator& comm,
        const bool force = false, void* stream = nullptr)
    {
        if (!vghost_)
        {
            std::fprintf(stderr, "Hamiltonian::applyLocal: updatePotentialHalo was not called\n");
            ::mgmol_b200::fail("precondition");
        }
        if (!hlphi_ || hlphi_->numst() != phi.numst())
        {
            hlphi_.reset(new ExtendedGridOrbitals<T>(phi.grid(), phi.numst()));
            itindex_ = -1;
        }
        const int new_index = 100 * phi.getIterativeIndex() + pot_->getIterativeIndex();
        if (force || new_index != itindex_)
        {
            lapOper_->applyWithPotPeer(comm, phi.getPsi(), phi.getLda(), pot_->vtot(),
                vghost_->data(), hlphi_->getPsi(), hlphi_->getLda(), phi.chromatic_number(),
                stream);
            hlphi_->incrementIterativeIndex();
            itindex_ = new_index;
        }
        return *hlphi_;
    }
    // src/Hamiltonian.cc:163-212: hij_dev += vel Phi1^T (H_loc Phi2), summed over ranks
    void addHlocal2matrix(ExtendedGridOrbitals<T>& phi1, ExtendedGridOrbitals<T>& phi2,
        double* hij_dev, double* work_dev, const bool force = false, mgb_comm* comm = nullptr,
        void* stream = nullptr)
    {
        applyLocal(phi2, force, stream);
        phi1.addDotWithNcol2Matrix(*hlphi_, hij_dev, work_dev, comm, stream);
    }
    // src/Hamiltonian.cc:214-239: hij_dev = vel Phi1^T (H_loc Phi2)
    void addHlocalij(ExtendedGridOrbitals<T>& phi1, ExtendedGridOrbitals<T>& phi2,
        double* hij_dev, mgb_comm* comm = nullptr, void* stream = nullptr)
    {
        applyLocal(phi2, false, stream);
        phi1.computeLocalProduct(*hlphi_, hij_dev, comm, stream);
    }

private:
    std::unique_ptr<Lap<T>> lapOper_;
    std::unique_ptr<Potentials> pot_;
    std::unique_ptr<ExtendedGridOrbitals<T>> hlphi_;
    std::unique_ptr<GridFuncVector<double>> vghost_;
    int itindex_;
};

// OrbitalsPreconditioning<OrbitalsType> (src/OrbitalsPreconditioning.h:27-70)
template <typename T>
class OrbitalsPreconditioning
{
public:
    OrbitalsPreconditioning() : handle_(nullptr), gamma_(-1.), is_set_(false) {}
    ~OrbitalsPreconditioning()
    {
        if (handle_) mgb_precond_destroy(handle_);
    }
    // src/OrbitalsPreconditioning.cc:44-84
    void setup(ExtendedGridOrbitals<T>& orbitals, const short mg_levels, const short lap_type)
    {
        assert(!is_set_);
        const Grid g = orbitals.grid().with_ghosts(ghostsFor(lap_type));
        MGB_CHECK(mgb_precond_create(
            &handle_, lap_type, mg_levels, g.c(), orbitals.chromatic_number()));
        mg_levels_ = mg_levels;
        is_set_    = true;
    }
    // the same with `currentMasks` (LocGridOrbitals): :59-67 Map2Masks +
    // GridFuncVector::setMasks
    void setup(ExtendedGridOrbitals<T>& orbitals, const short mg_levels, const short lap_type,
        const Masks* currentMasks)
    {
        setup(orbitals, mg_levels, lap_type);
        MGB_CHECK(mgb_precond_set_masks(handle_, currentMasks ? currentMasks->handle() : nullptr));
    }
    // src/OrbitalsPreconditioning.cc:120-145
    void setGamma(const Lap<T>& lapOper, const double vmax, const short mg_levels,
        const double small_eig)
    {
        gamma_ = mgb_gamma(lapOper.invDiagEl(), mg_levels, vmax, small_eig);
    }
    void setGamma(const double gamma) { gamma_ = gamma; }
    double gamma() const { return gamma_; }
    // src/OrbitalsPreconditioning.cc:87-117: orbitals <- M^-1 orbitals
    void precond_mg(ExtendedGridOrbitals<T>& orbitals, void* stream = nullptr)
    {
        assert(is_set_);
        assert(gamma_ > 0.);
        MGB_CHECK(mgb_precond_mg(handle_, dtype_of<T>::value, orbitals.getPsi(),
            orbitals.getLda(), orbitals.chromatic_number(), gamma_, stream));
        orbitals.incrementIterativeIndex();
    }
    // decomposed domains: the communicator the V-cycle reads / exchanges its ghosts
    // over (every trade_boundaries of Preconditioning<float>::mg); calls are then
    // collective.  nullptr: single rank.
    void setComm(Communicator* comm)
    {
        MGB_CHECK(mgb_precond_set_comm(handle_, comm ? comm->handle() : nullptr));
    }
    mgb_precond* handle() { return handle_; }

private:
    mgb_precond* handle_;
    double gamma_;
    short mg_levels_;
    bool is_set_;
};

// The non-local Kleinman-Bylander projectors overlapping the local box, as the sparse
// vectors KBprojectorSparse holds (src/KBprojectorSparse.h:39-52), with the two grid-sized
// steps that follow applyLocal in MGmol::getHpsiAndTheta (src/computeHij.cc:404-455):
// KBPsiMatrixSparse::computeKBpsi (src/KBPsiMatrixSparse.cc:136-212) and
// computeHnlPhiAndAdd2HPhi (src/computeHij.cc:294-375) over get_vnlpsi
// (src/get_vnlpsi.cc:24-87).  KBPROJDTYPE = ORBDTYPE (src/global.h:38).
template <typename T>
class KBProjectors
{
public:
    explicit KBProjectors(const Grid& grid) : grid_(grid), handle_(nullptr)
    {
        MGB_CHECK(mgb_kb_create(&handle_, dtype_of<T>::value, grid.size()));
    }
    ~KBProjectors()
    {
        if (handle_) mgb_kb_destroy(handle_);
    }
    KBProjectors(const KBProjectors&) = delete;
    KBProjectors& operator=(const KBProjectors&) = delete;
    // One ion (Ions::overlappingNL_ions order): nlindex = node positions in the no-ghost
    // storage, proj = nproj arrays of nlindex.size() values (host), coeff[p] = kbcoeff_p *
    // sign_p.  Returns the row of the ion's first projector in kbpsi.
    int addIon(const std::vector<int>& nlindex, const T* proj, const std::vector<double>& coeff)
    {
        int row = 0;
        MGB_CHECK(mgb_kb_add_ion(handle_, (int)nlindex.size(), nlindex.data(), (int)coeff.size(),
            proj, coeff.data(), &row));
        return row;
    }
    void commit() { MGB_CHECK(mgb_kb_commit(handle_)); }
    int nrows() const { return mgb_kb_nrows(handle_); }
    // kbpsi_dev[row * nfunc + f] = vel <beta_row | psi_f> (double, device, nrows x nfunc),
    // summed over the ranks (globalSumKBpsi).  lapOper: the reference's `flag` -- project
    // B phi instead of phi (Mehrstellen, kbBpsi); work_dev then holds B phi (numpt x nfunc).
    void computeKBpsi(const ExtendedGridOrbitals<T>& orbitals, double* kbpsi_dev,
        Communicator* comm = nullptr, const Lap<T>* lapOper = nullptr, T* work_dev = nullptr,
        void* stream = nullptr)
    {
        const T* psi    = orbitals.getPsi();
        const int nfunc = orbitals.chromatic_number();
        if (lapOper)
        {
            if (!work_dev)
            {
                std::fprintf(stderr, "KBProjectors::computeKBpsi: B phi needs a work block\n");
                ::mgmol_b200::fail("precondition");
            }
            lapOper->rhs(psi, orbitals.getLda(), work_dev, orbitals.getLda(), nfunc, nullptr, stream);
            psi = work_dev;
        }
        MGB_CHECK(mgb_kb_psi(handle_, dtype_of<T>::value, grid_.vel(), psi, orbitals.getLda(), nfunc,
            kbpsi_dev, stream));
        if (comm) comm->allreduce(kbpsi_dev, (size_t)nrows() * nfunc, stream);
    }
    // get_vnlpsi for every function: out_f = sum over the ions of (T)(sum_p alpha_p beta_p)
    void getVnlPsi(const double* kbpsi_dev, T* out_dev, const size_t ldo, const int nfunc,
        void* stream = nullptr)
    {
        MGB_CHECK(mgb_kb_vnlpsi(
            handle_, dtype_of<T>::value, kbpsi_dev, out_dev, ldo, nfunc, 0, stream));
    }
    // H phi += V_nl phi in one pass over the touched points; with the Mehrstellen operator
    // H phi += B (V_nl phi) (src/computeHij.cc:346-372), through two work blocks
    void computeHnlPhiAndAdd2HPhi(const double* kbpsi_dev, ExtendedGridOrbitals<T>& hphi,
        const Lap<T>* lapOper = nullptr, T* work_dev = nullptr, T* bwork_dev = nullptr,
        void* stream = nullptr)
    {
        const int nfunc = hphi.chromatic_number();
        if (lapOper && (lapOper->type() == MGB_LAP_4M || lapOper->type() == MGB_LAP_4MP))
        {
            if (!work_dev || !bwork_dev)
            {
                std::fprintf(stderr, "KBProjectors::computeHnlPhiAndAdd2HPhi: two work blocks\n");
                ::mgmol_b200::fail("precondition");
            }
            getVnlPsi(kbpsi_dev, work_dev, hphi.getLda(), nfunc, stream);
            lapOper->rhs(work_dev, hphi.getLda(), bwork_dev, hphi.getLda(), nfunc, nullptr, stream);
            MGB_CHECK(mgb_axpy(dtype_of<T>::value, hphi.getLda() * (size_t)nfunc, 1., bwork_dev,
                hphi.getPsi(), stream));
        }
        else
            MGB_CHECK(mgb_kb_vnlpsi(handle_, dtype_of<T>::value, kbpsi_dev, hphi.getPsi(),
                hphi.getLda(), nfunc, 1, stream));
        hphi.incrementIterativeIndex();
    }
    mgb_kb* handle() { return handle_; }

private:
    Grid grid_;
    mgb_kb* handle_;
};

// MGmol::computeResidualUsingHPhi (src/MGmol.cc:1227-1287):
// res = (B psi) theta - hphi in [Ry]; localT_dev = theta, column-major
// numst x numst double on the device
template <typename T>
void computeResidualUsingHPhi(const Lap<T>& lapOper, const ExtendedGridOrbitals<T>& psi,
    const ExtendedGridOrbitals<T>& hphi, const double* localT_dev, ExtendedGridOrbitals<T>& res,
    void* stream = nullptr)
{
    MGB_CHECK(mgb_residual(lapOper.type(), dtype_of<T>::value, psi.grid().c(), psi.getPsi(),
        psi.getLda(), hphi.getPsi(), hphi.getLda(), localT_dev, psi.numst(), res.getPsi(),
        res.getLda(), psi.chromatic_number(), nullptr, stream));
    res.incrementIterativeIndex();
}

// Rho::computeRhoSubdomainUsingBlas3 (src/Rho.cc:359-448) on the whole local box:
// rho_dev (RHODTYPE double, numpt values) += sum_ij X_ij phi1_i phi2_j, localX_dev
// column-major numst x numst double on the device
template <typename T>
void computeRhoUsingBlas3(const ExtendedGridOrbitals<T>& orbitals1, const double* localX_dev,
    double* rho_dev, const ExtendedGridOrbitals<T>* orbitals2 = nullptr, void* stream = nullptr)
{
    const ExtendedGridOrbitals<T>& o2 = orbitals2 ? *orbitals2 : orbitals1;
    MGB_CHECK(mgb_rho_blas3(dtype_of<T>::value, orbitals1.getNumpt(), orbitals1.chromatic_number(),
        orbitals1.getPsi(), orbitals1.getLda(), localX_dev, orbitals1.numst(), o2.getPsi(),
        o2.getLda(), rho_dev, stream));
}

// AndersonMix<T> (src/AndersonMix.h:21-52, src/AndersonMix.cc:27-319): host
// control flow and an m x m solve; everything grid-sized is T's assign, -=,
// dotProduct, axpy, scal.  The m x m matrix is scaled to a unit diagonal and
// its determinant (product of Jacobi eigenvalues) tested exactly as the
// reference does with DSYEV, then solved by Cholesky (DPOTRF/DPOTRS).
template <class T>
class AndersonMix
{
public:
    // make(x) must return a new T shaped like x (the reference's T("xi", x))
    template <class Make>
    AndersonMix(const int m, const double beta, T& x, Make make)
        : m_(m), mm_(-1), beta_(beta), x_(x), mat_((size_t)m * m, 0.), rhs_(m, 0.), theta_(m, 0.)
    {
        for (int i = 0; i < m; i++)
        {
            xi_.emplace_back(make(x));
            fi_.emplace_back(make(x));
        }
        if (m > 1) tmp_.reset(make(x));
    }
    void restart() { mm_ = -1; }
    int historyLength() const { return mm_; }
    // src/AndersonMix.cc:72-319
    void update(T& f, T& work)
    {
        if (mm_ < m_) mm_++;
        if (mm_ > 0)
        {
            for (int i = 0; i < mm_; i++)
            {
                work.assign(f);
                work -= *fi_[i];
                mat_[i * m_ + i] = work.dotProduct(work);
                rhs_[i]          = work.dotProduct(f);
                for (int j = 0; j < i; j++)
                {
                    tmp_->assign(f);
                    *tmp_ -= *fi_[j];
                    mat_[j * m_ + i] = work.dotProduct(*tmp_);
                }
            }
            solve();
        }
        if (m_ > 0)
        {
            mixHistory(x_, xi_, work);
            mixHistory(f, fi_, work);
        }
        x_.axpy(mm_ > 0 ? beta_ : 1., f);
    }

private:
    static constexpr double min_det_mat = 0.01, max_theta = 0.5, min_theta = -3.;
    void mixHistory(T& cur, std::vector<std::unique_ptr<T>>& hist, T& work)
    {
        work.assign(cur);
        double factor = 1.;
        for (int j = 0; j < mm_; j++)
            factor -= theta_[j];
        if (mm_ > 0) cur.scal(factor);
        for (int j = 0; j < mm_; j++)
            cur.axpy(theta_[j], *hist[j]);
        std::unique_ptr<T> last = std::move(hist[m_ - 1]);
        for (int j = m_ - 1; j > 0; j--)
            hist[j] = std::move(hist[j - 1]);
        hist[0] = std::move(last);
        hist[0]->assign(work);
    }
    // lower triangle of mat_ (column-major, leading dimension m_) -> dense n x n
    std::vector<double> dense(const int n) const
    {
        std::vector<double> a((size_t)n * n);
        for (int i = 0; i < n; i++)
            for (int j = 0; j <= i; j++)
                a[i * n + j] = a[j * n + i] = mat_[j * m_ + i];
        return a;
    }
    static double detScaled(std::vector<double> a, const int n)
    {
        std::vector<double> d(n);
        for (int i = 0; i < n; i++)
            d[i] = 1. / std::sqrt(a[i * n + i]);
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++)
                a[i * n + j] *= d[i] * d[j];
        // determinant by Gaussian elimination (= product of the eigenvalues)
        double det = 1.;
        for (int k = 0; k < n; k++)
        {
            det *= a[k * n + k];
            if (a[k * n + k] == 0.) return 0.;
            for (int i = k + 1; i < n; i++)
            {
                const double l = a[i * n + k] / a[k * n + k];
                for (int j = k; j < n; j++)
                    a[i * n + j] -= l * a[k * n + j];
            }
        }
        return det;
    }
    void solve()
    {
        bool flag = true;
        while (flag)
        {
            flag = false;
            while (mm_ > 1)
            {
                if (detScaled(dense(mm_), mm_) < min_det_mat)
                    mm_--;
                else
                    break;
            }
            const int n = mm_;
            std::vector<double> a = dense(n), y(n);
            // Cholesky A = L L^T, then two triangular solves
            for (int j = 0; j < n; j++)
            {
                for (int k = 0; k < j; k++)
                    a[j * n + j] -= a[j * n + k] * a[j * n + k];
                if (!(a[j * n + j] > 0.))
                {
                    std::fprintf(stderr, "AndersonMix, dpotrf: matrix not positive definite\n");
                    std::exit(0);
                }
                a[j * n + j] = std::sqrt(a[j * n + j]);
                for (int i = j + 1; i < n; i++)
                {
                    for (int k = 0; k < j; k++)
                        a[i * n + j] -= a[i * n + k] * a[j * n + k];
                    a[i * n + j] /= a[j * n + j];
                }
            }
            for (int i = 0; i < n; i++)
            {
                double s = rhs_[i];
                for (int k = 0; k < i; k++)
                    s -= a[i * n + k] * y[k];
                y[i] = s / a[i * n + i];
            }
            for (int i = n - 1; i >= 0; i--)
            {
                double s = y[i];
                for (int k = i + 1; k < n; k++)
                    s -= a[k * n + i] * theta_[k];
                theta_[i] = s / a[i * n + i];
            }
            for (int j = 0; j < mm_; j++)
            {
                if (theta_[j] > max_theta)
                {
                    if (mm_ > 1)
                    {
                        mm_--;
                        flag = true;
                        break;
                    }
                    theta_[j] = theta_[j] > 1. ? -0.5 : 0.;
                }
                else if (theta_[j] < min_theta)
                {
                    if (mm_ > 1)
                    {
                        mm_--;
                        flag = true;
                        break;
                    }
                    theta_[j] = min_theta;
                }
            }
        }
    }

    const int m_;
    int mm_;
    double beta_;
    T& x_;
    std::vector<std::unique_ptr<T>> xi_, fi_;
    std::unique_ptr<T> tmp_;
    std::vector<double> mat_, rhs_, theta_;
};

} // namespace mgmol_b200

#endif // MGMOL_B200_HPP
