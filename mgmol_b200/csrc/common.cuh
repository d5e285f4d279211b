// Shared host/device helpers for the mgmol_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "mgmol_b200.h"

namespace mgb
{

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launch_count;

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define MGB_CUDA(call)                                                         \
    do                                                                         \
    {                                                                          \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess)                                                \
        {                                                                      \
            ::mgb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,     \
                cudaGetErrorString(e__));                                      \
            return MGB_ECUDA;                                                  \
        }                                                                      \
    } while (0)

// after every kernel launch: count it and surface launch-configuration errors
#define MGB_LAUNCHED(name)                                                     \
    do                                                                         \
    {                                                                          \
        ::mgb::g_launch_count.fetch_add(1, std::memory_order_relaxed);         \
        cudaError_t e__ = cudaGetLastError();                                  \
        if (e__ != cudaSuccess)                                                \
        {                                                                      \
            ::mgb::set_error("%s:%d: launch of %s failed -> %s", __FILE__,     \
                __LINE__, name, cudaGetErrorString(e__));                      \
            return MGB_ECUDA;                                                  \
        }                                                                      \
    } while (0)

#define MGB_REQUIRE(cond, ...)                                                 \
    do                                                                         \
    {                                                                          \
        if (!(cond))                                                           \
        {                                                                      \
            ::mgb::set_error(__VA_ARGS__);                                     \
            return MGB_EINVAL;                                                 \
        }                                                                      \
    } while (0)

int require_device();

// Layout of one local box, ghosted with width g (pb::Grid::inc / sizeg).
struct Box
{
    int nx, ny, nz, g;
    long long incy, incx, sizeg, npt;
    __host__ __device__ Box() {}
    __host__ __device__ Box(int nx_, int ny_, int nz_, int g_)
        : nx(nx_), ny(ny_), nz(nz_), g(g_)
    {
        incy  = nz + 2 * g;
        incx  = (long long)(ny + 2 * g) * incy;
        sizeg = (long long)(nx + 2 * g) * incx;
        npt   = (long long)nx * ny * nz;
    }
    // offset of interior point (ix,iy,iz) inside a ghosted function
    __host__ __device__ long long at(int ix, int iy, int iz) const
    {
        return (long long)(ix + g) * incx + (long long)(iy + g) * incy + (iz + g);
    }
};

inline Box box_of(const mgb_grid* gr, int g)
{
    return Box(gr->dim[0], gr->dim[1], gr->dim[2], g);
}

int check_grid(const mgb_grid* gr);

// 2-D thread block (z fastest, then y) covering interior rows; grid.y = x
// planes, grid.z = functions.
struct RowLaunch
{
    dim3 block, grid;
    int tiles_z;
};
inline RowLaunch row_launch(int nx, int ny, int nz, int nfunc)
{
    RowLaunch L;
    int bx = 32;
    while (bx < nz && bx < 128) bx *= 2;
    int by    = 256 / bx;
    L.block   = dim3(bx, by, 1);
    L.tiles_z = (nz + bx - 1) / bx;
    L.grid    = dim3((unsigned)(L.tiles_z * ((ny + by - 1) / by)), (unsigned)nx,
        (unsigned)nfunc);
    return L;
}

} // namespace mgb
