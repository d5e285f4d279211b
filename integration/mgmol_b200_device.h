// Reference-side shim: what a maintainer adds to MGmol so that its
// MemorySpace::Device seam runs on libmgmol_b200.so instead of MAGMA /
// OpenMP-offload.  It is written against the reference's OWN headers
// (src/pb/Grid.h, src/pb/FDkernels.h, src/tools/memory_space.h) and built here by
// integration/Makefile against them (with the serial MPI stub of oracle/); it
// contains no MGmol code, only the two overload sets the reference leaves
// unimplemented or MAGMA-bound:
//
//   * MemorySpace::Memory<T, MemorySpace::Device>            src/tools/memory_space.h:253-333
//     (allocate / free / copy / copy_view_to_host / copy_view_to_dev / set) over
//     mgb_malloc, mgb_free, mgb_copy_*, mgb_memset instead of magma_malloc ...
//   * pb::FDkernelDel2_{2nd,4th,4th_Mehr,6th,8th}(const Grid&, T* v, T* b, size_t nfunc,
//     MemorySpace::Device) and pb::FDkernelRHS_4th_Mehr1(..., short rhs_ghosts, size_t
//     nfunc, MemorySpace::Device)                             src/pb/FDkernels.h:38-74
//     (the reference's Device versions abort() except the OpenMP-offload 4th order one)
//     over mgb_fd_apply; v and b are device blocks of nfunc ghosted functions, exactly
//     the arguments the Host overloads take.
//
// With -DHAVE_MAGMA the reference's own header already declares these Device overloads
// (and defines the aborting ones inline); the shim then replaces those inline bodies.  The
// test in this directory builds WITHOUT HAVE_MAGMA (no MAGMA in this image), where the tag
// types exist but the Device overloads and the Memory<T, Device> specialisation do not.
#ifndef MGMOL_B200_DEVICE_SHIM_H
#define MGMOL_B200_DEVICE_SHIM_H

#include "FDkernels.h" // the reference's: pb::Grid, MemorySpace tags, Host overloads
#include "mgmol_b200.h"

#include <cstdio>
#include <cstdlib>

namespace mgmol_b200_shim
{
inline void check(const int rc, const char* what)
{
    if (rc != MGB_OK)
    {
        // the reference's error behaviour on this path: print and abort
        // (src/pb/FDkernels.h:42-45)
        std::fprintf(stderr, "%s: %s\n", what, mgb_last_error());
        std::abort();
    }
}
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
// pb::Grid -> mgb_grid (dims, ghosts, spacing, the process grid of its PEenv)
inline mgb_grid grid_of(const pb::Grid& g, const int bc = 1)
{
    mgb_grid m;
    for (int d = 0; d < 3; d++)
    {
        m.dim[d]   = (int)g.dim(d);
        m.gdim[d]  = (int)g.gdim(d);
        m.h[d]     = g.hgrid(d);
        m.bc[d]    = bc;
        m.nproc[d] = g.mype_env().n_mpi_task(d);
        m.coord[d] = g.mype_env().my_mpi(d);
    }
    m.ghosts = g.ghost_pt();
    return m;
}
template <typename T>
inline void fd(const int kind, const pb::Grid& grid, T* v, T* b, const size_t nfunc,
    const short rhs_ghosts, const char* name)
{
    const mgb_grid m = grid_of(grid);
    check(mgb_fd_apply(kind, dtype_of<T>::value, &m, v, b, (int)nfunc, rhs_ghosts, nullptr), name);
    check(mgb_stream_sync(nullptr), name); // the reference's kernels return when done
}
}

#ifndef HAVE_MAGMA
namespace MemorySpace
{
// src/tools/memory_space.h:253-333 with the device library instead of MAGMA
template <typename T>
struct Memory<T, MemorySpace::Device>
{
    static T* allocate(unsigned int size)
    {
        void* p = nullptr;
        mgmol_b200_shim::check(mgb_malloc(&p, (size_t)size * sizeof(T)), "Memory<T,Device>::allocate");
        return static_cast<T*>(p);
    }
    static T* allocate_host_view(unsigned int size) { return new T[size]; }
    static void free(T* ptr_dev) { mgmol_b200_shim::check(mgb_free(ptr_dev), "Memory<T,Device>::free"); }
    static void free_host_view(T* ptr) { delete[] ptr; }
    static void copy(T const* in, unsigned int size, T* out)
    {
        mgmol_b200_shim::check(
            mgb_copy_dev(out, in, (size_t)size * sizeof(T), nullptr), "Memory<T,Device>::copy");
    }
    static void copy_view_to_host(T* vec_dev, unsigned int size, T*& vec)
    {
        mgmol_b200_shim::check(mgb_copy_to_host(vec, vec_dev, (size_t)size * sizeof(T), nullptr),
            "Memory<T,Device>::copy_view_to_host");
        mgmol_b200_shim::check(mgb_stream_sync(nullptr), "Memory<T,Device>::copy_view_to_host");
    }
    static void copy_view_to_dev(T* vec, unsigned int size, T* vec_dev)
    {
        mgmol_b200_shim::check(mgb_copy_to_dev(vec_dev, vec, (size_t)size * sizeof(T), nullptr),
            "Memory<T,Device>::copy_view_to_dev");
        mgmol_b200_shim::check(mgb_stream_sync(nullptr), "Memory<T,Device>::copy_view_to_dev");
    }
    static void set(T* ptr, unsigned int size, int val)
    {
        mgmol_b200_shim::check(
            mgb_memset(ptr, val, (size_t)size * sizeof(T), nullptr), "Memory<T,Device>::set");
    }
};
}
#endif

namespace pb
{
// src/pb/FDkernels.h:38-74: the Device overloads
template <typename ScalarType>
void FDkernelDel2_2nd(
    const Grid& grid, ScalarType* v, ScalarType* b, const size_t nfunc, MemorySpace::Device)
{
    mgmol_b200_shim::fd(MGB_FD_DEL2_2ND, grid, v, b, nfunc, 0, "FDkernelDel2_2nd");
}
template <typename ScalarType>
void FDkernelDel2_4th(
    const Grid& grid, ScalarType* v, ScalarType* b, const size_t nfunc, MemorySpace::Device)
{
    mgmol_b200_shim::fd(MGB_FD_DEL2_4TH, grid, v, b, nfunc, 0, "FDkernelDel2_4th");
}
template <typename ScalarType>
void FDkernelDel2_4th_Mehr(
    const Grid& grid, ScalarType* v, ScalarType* b, const size_t nfunc, MemorySpace::Device)
{
    mgmol_b200_shim::fd(MGB_FD_DEL2_4TH_MEHR, grid, v, b, nfunc, 0, "FDkernelDel2_4th_Mehr");
}
template <typename ScalarType>
void FDkernelDel2_6th(
    const Grid& grid, ScalarType* v, ScalarType* b, const size_t nfunc, MemorySpace::Device)
{
    mgmol_b200_shim::fd(MGB_FD_DEL2_6TH, grid, v, b, nfunc, 0, "FDkernelDel2_6th");
}
template <typename ScalarType>
void FDkernelDel2_8th(
    const Grid& grid, ScalarType* v, ScalarType* b, const size_t nfunc, MemorySpace::Device)
{
    mgmol_b200_shim::fd(MGB_FD_DEL2_8TH, grid, v, b, nfunc, 0, "FDkernelDel2_8th");
}
template <typename ScalarType>
void FDkernelRHS_4th_Mehr1(const Grid& grid, ScalarType* v, ScalarType* b, const short rhs_ghosts,
    const size_t nfunc, MemorySpace::Device)
{
    mgmol_b200_shim::fd(MGB_FD_RHS_4TH_MEHR1, grid, v, b, nfunc, rhs_ghosts, "FDkernelRHS_4th_Mehr1");
}
}
#endif
