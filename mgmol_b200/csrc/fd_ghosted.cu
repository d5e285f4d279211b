// Batched kernels on ghosted blocks: the MemorySpace::Device side of
// pb::FDkernel*, pb::MGkernel* and the pb::GridFuncVector batch operations.
//
// This translation unit is compiled with -fmad=false and every expression
// follows the operand order and the float<->double casts of the reference
// loop it replaces, so results are BIT-IDENTICAL to the reference CPU build
// (generic x86-64, no FMA contraction).  These kernels are HBM-bound; the
// extra DMUL/DADD issue slots are free.  The roofline-tuned fused kernels
// live in hpsi_fused.cu / mg_fused.cu.
#include "common.cuh"
#include "hpsi.h"

namespace mgb
{

// ---------------------------------------------------------------------------
// Stencil bodies.  p points at the centre value inside a ghosted function.
// ---------------------------------------------------------------------------
struct Coef4M
{
    double c0, cx, cy, cz, cxy, cyz, cxz;
};
struct CoefCross
{
    double c0;
    double c[3][4]; // c[dir][k-1] for offsets 1..4
};

// src/pb/FDkernels.cc:457-463
static Coef4M coef_4m(const double h[3])
{
    const double inv12 = 1. / 12.;
    const double a[3]
        = { 1. / (h[0] * h[0]), 1. / (h[1] * h[1]), 1. / (h[2] * h[2]) };
    Coef4M c;
    c.c0  = 16. * inv12 * (a[0] + a[1] + a[2]);
    c.cx  = -10. * inv12 * a[0] + 0.125 * c.c0;
    c.cy  = -10. * inv12 * a[1] + 0.125 * c.c0;
    c.cz  = -10. * inv12 * a[2] + 0.125 * c.c0;
    c.cxy = -inv12 * (a[0] + a[1]);
    c.cyz = -inv12 * (a[2] + a[1]);
    c.cxz = -inv12 * (a[0] + a[2]);
    return c;
}

// order 2: FDkernels.cc:28-42; 4: :86-104; 6: :263-281; 8: :350-374
static CoefCross coef_cross(int order, const double h[3])
{
    CoefCross c;
    for (int d = 0; d < 3; d++)
        for (int k = 0; k < 4; k++)
            c.c[d][k] = 0.;
    const double a[3]
        = { 1. / (h[0] * h[0]), 1. / (h[1] * h[1]), 1. / (h[2] * h[2]) };
    if (order == 2)
    {
        for (int d = 0; d < 3; d++)
            c.c[d][0] = -a[d];
        c.c0 = -2. * (c.c[0][0] + c.c[1][0] + c.c[2][0]);
    }
    else if (order == 4)
    {
        const double inv12 = 1. / 12.;
        for (int d = 0; d < 3; d++)
        {
            const double cc = inv12 * a[d];
            c.c[d][0]       = -16. * cc;
            c.c[d][1]       = 1. * cc;
        }
        c.c0 = -2.
               * (c.c[0][0] + c.c[0][1] + c.c[1][0] + c.c[1][1] + c.c[2][0]
                     + c.c[2][1]);
    }
    else if (order == 6)
    {
        for (int d = 0; d < 3; d++)
        {
            const double cc = (1. / 180.) * a[d];
            c.c[d][0]       = -270. * cc;
            c.c[d][1]       = 27. * cc;
            c.c[d][2]       = -2. * cc;
        }
        c.c0 = -2.
               * (c.c[0][0] + c.c[0][1] + c.c[0][2] + c.c[1][0] + c.c[1][1]
                     + c.c[1][2] + c.c[2][0] + c.c[2][1] + c.c[2][2]);
    }
    else
    {
        for (int d = 0; d < 3; d++)
        {
            const double cc = (1. / 5040.) * a[d];
            c.c[d][0]       = -8064. * cc;
            c.c[d][1]       = 1008. * cc;
            c.c[d][2]       = -128. * cc;
            c.c[d][3]       = 9. * cc;
        }
        c.c0 = -2.
               * (c.c[0][0] + c.c[0][1] + c.c[0][2] + c.c[0][3] + c.c[1][0]
                     + c.c[1][1] + c.c[1][2] + c.c[1][3] + c.c[2][0] + c.c[2][1]
                     + c.c[2][2] + c.c[2][3]);
    }
    return c;
}

// thread -> interior point of function f; returns false when out of range
__device__ __forceinline__ bool interior_point(
    const Box& b, int tiles_z, int& ix, int& iy, int& iz, int& f)
{
    const int tz = blockIdx.x % tiles_z;
    const int ty = blockIdx.x / tiles_z;
    iz           = tz * blockDim.x + threadIdx.x;
    iy           = ty * blockDim.y + threadIdx.y;
    ix           = blockIdx.y;
    f            = blockIdx.z;
    return iz < b.nz && iy < b.ny;
}

// FDkernelDel2_4th_Mehr, src/pb/FDkernels.cc:447-520
template <typename T>
__global__ void k_del2_4th_mehr(
    Box b, int tiles_z, Coef4M c, const T* __restrict__ v, T* __restrict__ u)
{
    int ix, iy, iz, f;
    if (!interior_point(b, tiles_z, ix, iy, iz, f)) return;
    const long long o = (long long)f * b.sizeg + b.at(ix, iy, iz);
    const T* v0       = v + o;
    const long long X = b.incx, Y = b.incy;
    const T* vmx = v0 - X;
    const T* vpx = v0 + X;
    const T* vmy = v0 - Y;
    const T* vpy = v0 + Y;
    u[o] = (T)(c.c0 * (double)v0[0] + c.cz * (double)(v0[-1] + v0[1])
               + c.cy * (double)(vmy[0] + vpy[0])
               + c.cx * (double)(vmx[0] + vpx[0])
               + c.cxz * (double)(vmx[-1] + vmx[1] + vpx[-1] + vpx[1])
               + c.cyz * (double)(vmy[-1] + vmy[1] + vpy[-1] + vpy[1])
               + c.cxy * (double)(vmx[-Y] + vpx[-Y] + vmx[Y] + vpx[Y]));
}

// FDkernelRHS_4th_Mehr1, src/pb/FDkernels.cc:522-584
template <typename T>
__global__ void k_rhs_4th_mehr1(
    Box b, Box br, int tiles_z, const T* __restrict__ v, T* __restrict__ rhs)
{
    int ix, iy, iz, f;
    if (!interior_point(b, tiles_z, ix, iy, iz, f)) return;
    const T* v0        = v + (long long)f * b.sizeg + b.at(ix, iy, iz);
    const long long X  = b.incx, Y = b.incy;
    const double c0 = 0.5, c1 = 1. / 12.;
    rhs[(long long)f * br.sizeg + br.at(ix, iy, iz)]
        = (T)(c0 * (double)v0[0]
              + c1
                    * (double)(v0[-X] + v0[X] + v0[-Y] + v0[Y] + v0[-1]
                               + v0[1]));
}

// FDkernelDel2_2nd, src/pb/FDkernels.cc:18-75 (neighbour pairs are added in
// T, then promoted through the double coefficient)
template <typename T>
__global__ void k_del2_2nd(
    Box b, int tiles_z, CoefCross c, const T* __restrict__ v, T* __restrict__ u)
{
    int ix, iy, iz, f;
    if (!interior_point(b, tiles_z, ix, iy, iz, f)) return;
    const long long o = (long long)f * b.sizeg + b.at(ix, iy, iz);
    const T* p        = v + o;
    const long long X = b.incx, Y = b.incy;
    u[o] = (T)(c.c0 * p[0] + c.c[0][0] * (p[-X] + p[X])
               + c.c[1][0] * (p[-Y] + p[Y]) + c.c[2][0] * (p[-1] + p[1]));
}

// FDkernelDel2_4th / 6th / 8th, src/pb/FDkernels.cc:77-162, 255-344, 346-445:
// every tap cast to double, pairs added in double, one final cast.
template <typename T, int R>
__global__ void k_del2_cross(
    Box b, int tiles_z, CoefCross c, const T* __restrict__ v, T* __restrict__ u)
{
    int ix, iy, iz, f;
    if (!interior_point(b, tiles_z, ix, iy, iz, f)) return;
    const long long o = (long long)f * b.sizeg + b.at(ix, iy, iz);
    const T* p        = v + o;
    const long long X = b.incx, Y = b.incy;
    double s          = c.c0 * (double)p[0];
#pragma unroll
    for (int k = 1; k <= R; k++)
    {
        s = s + c.c[0][k - 1] * ((double)p[-k * X] + (double)p[k * X]);
        s = s + c.c[1][k - 1] * ((double)p[-k * Y] + (double)p[k * Y]);
        s = s + c.c[2][k - 1] * ((double)p[-k] + (double)p[k]);
    }
    u[o] = (T)s;
}


// FDoper<T>::del2_4th_withPot, src/pb/FDoper.cc:321-399: ghosted input (g>=2),
// no-ghost potential, no-ghost output (leading dimension ldo).
template <typename T>
__global__ void k_del2_4th_withpot(Box b, int tiles_z, CoefCross c,
    const T* __restrict__ v, const double* __restrict__ pot,
    T* __restrict__ out, long long ldo)
{
    int ix, iy, iz, f;
    if (!interior_point(b, tiles_z, ix, iy, iz, f)) return;
    const T* p        = v + (long long)f * b.sizeg + b.at(ix, iy, iz);
    const long long X = b.incx, Y = b.incy;
    const long long o = ((long long)ix * b.ny + iy) * b.nz + iz;
    out[(long long)f * ldo + o]
        = (T)((c.c0 + pot[o]) * (double)p[0]
              + c.c[0][0] * ((double)p[-X] + (double)p[X])
              + c.c[1][0] * ((double)p[-Y] + (double)p[Y])
              + c.c[2][0] * ((double)p[-1] + (double)p[1])
              + c.c[0][1] * ((double)p[-2 * X] + (double)p[2 * X])
              + c.c[1][1] * ((double)p[-2 * Y] + (double)p[2 * Y])
              + c.c[2][1] * ((double)p[-2] + (double)p[2]));
}

// ---------------------------------------------------------------------------
// Elementwise GridFuncVector operations
// ---------------------------------------------------------------------------

// GridFuncVector::pointwiseProduct, src/pb/GridFuncVector.cc:90-136
template <typename T>
__global__ void k_pointwise_product(long long sizeg, int nfunc,
    const T* __restrict__ A, const double* __restrict__ V, T* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sizeg) return;
    const double vv = V[i];
    for (int f = blockIdx.y; f < nfunc; f += gridDim.y)
        out[(long long)f * sizeg + i] = (T)(A[(long long)f * sizeg + i] * vv);
}

// MPaxpy, src/linear_algebra/mputils.cc:222-244
template <typename T>
__global__ void k_axpy(size_t n, double alpha, const T* __restrict__ x, T* y)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        y[i] += (T)(alpha * (double)x[i]);
}

// MPscal, src/linear_algebra/mputils.cc:53-83
template <typename T>
__global__ void k_scal(size_t n, double alpha, T* x)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        x[i] = (alpha == 0.) ? (T)0 : (T)(alpha * (double)x[i]);
}

// Jacobi tail: w -= f ; v += -omega*w over the whole ghosted extent
// (GridFuncVector::jacobi, src/pb/GridFuncVector.cc:2420-2422)
__global__ void k_jacobi_tail(size_t n, double momega,
    const float* __restrict__ f, float* __restrict__ w, float* __restrict__ v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
    {
        float wi = w[i];
        wi += (float)(-1. * (double)f[i]);
        w[i] = wi;
        v[i] += (float)(momega * (double)wi);
    }
}

// MPdot (double accumulation of double products), deterministic two-pass tree
template <typename T>
__global__ void k_dot_partial(size_t n, const T* __restrict__ x,
    const T* __restrict__ y, double* __restrict__ partial)
{
    __shared__ double sm[32];
    double s = 0.;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        s += (double)x[i] * (double)y[i];
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32)
    {
        s = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.;
        for (int o = 16; o > 0; o >>= 1)
            s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }
}
__global__ void k_dot_final(int nblocks, const double* __restrict__ partial,
    double* __restrict__ out)
{
    __shared__ double sm[32];
    double s = 0.;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x)
        s += partial[i];
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32)
    {
        s = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.;
        for (int o = 16; o > 0; o >>= 1)
            s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) out[0] = s;
    }
}

// ---------------------------------------------------------------------------
// Ghost add / strip with precision conversion
// ---------------------------------------------------------------------------
// BlockVector::setDataWithGhosts + GridFunc::assign: the whole ghosted block
// is written (ghost points = 0), so no separate memset pass is needed.
template <typename Tin, typename Tout>
__global__ void k_set_with_ghosts(Box b, const Tin* __restrict__ in,
    long long ld, Tout* __restrict__ out)
{
    const int Z = blockIdx.x * blockDim.x + threadIdx.x;
    const int Y = blockIdx.y * blockDim.y + threadIdx.y;
    if (Z >= b.nz + 2 * b.g || Y >= b.ny + 2 * b.g) return;
    const int nxg = b.nx + 2 * b.g;
    const int f   = blockIdx.z / nxg;
    const int X   = blockIdx.z % nxg;
    const int ix = X - b.g, iy = Y - b.g, iz = Z - b.g;
    Tout val = (Tout)0;
    if (ix >= 0 && ix < b.nx && iy >= 0 && iy < b.ny && iz >= 0 && iz < b.nz)
        val = (Tout)in[(long long)f * ld + ((long long)ix * b.ny + iy) * b.nz + iz];
    out[(long long)f * b.sizeg + (long long)X * b.incx + (long long)Y * b.incy + Z]
        = val;
}

template <typename Tin, typename Tout>
__global__ void k_get_values(Box b, int tiles_z, const Tin* __restrict__ in,
    Tout* __restrict__ out, long long ld)
{
    int ix, iy, iz, f;
    if (!interior_point(b, tiles_z, ix, iy, iz, f)) return;
    out[(long long)f * ld + ((long long)ix * b.ny + iy) * b.nz + iz]
        = (Tout)in[(long long)f * b.sizeg + b.at(ix, iy, iz)];
}

// ---------------------------------------------------------------------------
// Single-rank boundary trade
// ---------------------------------------------------------------------------
struct BcInfo
{
    int dir[3];   // 1 = Dirichlet direction
    int first[3]; // rank is first / last along the direction
    int last[3];
};

// GridFunc::setBoundaryValues(0, direction), src/pb/GridFunc.cc:2192-2336 --
// literal zero set, including the i0/i1/j0/j1 loop bounds.
template <typename T>
__global__ void k_dirichlet_zero(Box b, BcInfo bc, T* __restrict__ u)
{
    const int Z = blockIdx.x * blockDim.x + threadIdx.x;
    const int Y = blockIdx.y * blockDim.y + threadIdx.y;
    if (Z >= b.nz + 2 * b.g || Y >= b.ny + 2 * b.g) return;
    const int nxg = b.nx + 2 * b.g;
    const int f   = blockIdx.z / nxg;
    const int X   = blockIdx.z % nxg;
    const int g   = b.g;
    const int i0  = bc.first[0] ? g + 1 : 0;
    const int i1  = bc.last[0] ? g + b.nx : 2 * g + b.nx;
    const int j0  = bc.first[1] ? g + 1 : 0;
    const int j1  = bc.last[1] ? g + b.ny : 2 * g + b.ny;
    const bool A  = (X >= i0 && X < i1);
    bool zero     = false;
    if (A && bc.dir[2] && Y >= j0 && Y < j1)
        zero = (bc.first[2] && Z <= g) || (bc.last[2] && Z >= b.nz + g);
    if (A && bc.dir[1])
        zero = zero || (bc.first[1] && Y <= g) || (bc.last[1] && Y >= b.ny + g);
    if (bc.dir[0])
        zero = zero || (bc.first[0] && X <= g) || (bc.last[0] && X >= b.nx + g);
    if (zero)
        u[(long long)f * b.sizeg + (long long)X * b.incx + (long long)Y * b.incy
            + Z]
            = (T)0;
}

// local periodic wraps, src/pb/GridFuncVector.cc:586-603 (Y), :1052-1071 (Z),
// East/West analogue (X).  dir: 1 = Y (interior x, interior z), 2 = Z
// (interior x, all y), 0 = X (all y, all z).
template <typename T>
__global__ void k_wrap(Box b, int dir, T* __restrict__ u)
{
    const int g = b.g;
    // enumerate ghost points of this direction: a = layer in [0, 2g)
    const int a  = blockIdx.y % (2 * g);
    const int f  = blockIdx.y / (2 * g);
    T* pu        = u + (long long)f * b.sizeg;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (dir == 1)
    {
        // t over nx * nz
        if (t >= (long long)b.nx * b.nz) return;
        const int ix = (int)(t / b.nz), iz = (int)(t % b.nz);
        const int Yg = (a < g) ? a : b.ny + a;          // ghost row
        const int Ys = (a < g) ? a + b.ny : a;          // source row
        const long long base = (long long)(ix + g) * b.incx + iz + g;
        pu[base + (long long)Yg * b.incy] = pu[base + (long long)Ys * b.incy];
    }
    else if (dir == 2)
    {
        const int nyg = b.ny + 2 * g;
        if (t >= (long long)b.nx * nyg) return;
        const int ix = (int)(t / nyg), Y = (int)(t % nyg);
        const int Zg = (a < g) ? a : b.nz + a;
        const int Zs = (a < g) ? a + b.nz : a;
        const long long base = (long long)(ix + g) * b.incx + (long long)Y * b.incy;
        pu[base + Zg] = pu[base + Zs];
    }
    else
    {
        if (t >= b.incx) return;
        const int Xg = (a < g) ? a : b.nx + a;
        const int Xs = (a < g) ? a + b.nx : a;
        pu[(long long)Xg * b.incx + t] = pu[(long long)Xs * b.incx + t];
    }
}

// all-periodic closed form of the Y->Z->X sequence: every ghost point takes
// the interior value at the wrapped index.
template <typename T>
__global__ void k_fill_ghosts_periodic(Box b, T* __restrict__ u)
{
    const int Z = blockIdx.x * blockDim.x + threadIdx.x;
    const int Y = blockIdx.y * blockDim.y + threadIdx.y;
    if (Z >= b.nz + 2 * b.g || Y >= b.ny + 2 * b.g) return;
    const int nxg = b.nx + 2 * b.g;
    const int f   = blockIdx.z / nxg;
    const int X   = blockIdx.z % nxg;
    const int g   = b.g;
    int ix = X - g, iy = Y - g, iz = Z - g;
    const bool inside
        = ix >= 0 && ix < b.nx && iy >= 0 && iy < b.ny && iz >= 0 && iz < b.nz;
    if (inside) return;
    ix = (ix + b.nx) % b.nx;
    iy = (iy + b.ny) % b.ny;
    iz = (iz + b.nz) % b.nz;
    T* pu = u + (long long)f * b.sizeg;
    pu[(long long)X * b.incx + (long long)Y * b.incy + Z] = pu[b.at(ix, iy, iz)];
}

// ---------------------------------------------------------------------------
// Multigrid transfer kernels
// ---------------------------------------------------------------------------
// MGkernelRestrict3D, src/pb/MGkernels.cc:203-283; thread per coarse point
template <typename T>
__global__ void k_restrict3D(Box bf, Box bc, int tiles_z,
    const T* __restrict__ fine, T* __restrict__ coarse)
{
    int ix, iy, iz, f;
    if (!interior_point(bc, tiles_z, ix, iy, iz, f)) return;
    const long long X = bf.incx, Y = bf.incy;
    const T* u0 = fine + (long long)f * bf.sizeg + bf.at(2 * ix, 2 * iy, 2 * iz);
    const T* umx = u0 - X;
    const T* upx = u0 + X;
    const T* umy = u0 - Y;
    const T* upy = u0 + Y;
    const T* umxpy = u0 - X + Y;
    const T* upxpy = u0 + X + Y;
    const T* umxmy = u0 - X - Y;
    const T* upxmy = u0 + X - Y;
    double face = (double)upx[0] + (double)umx[0] + (double)upy[0]
                  + (double)umy[0] + (double)u0[-1] + (double)u0[1];
    double corner = (double)upxpy[-1] + (double)upxpy[1] + (double)upxmy[-1]
                    + (double)upxmy[1] + (double)umxpy[-1] + (double)umxpy[1]
                    + (double)umxmy[-1] + (double)umxmy[1];
    double edge = (double)upy[-1] + (double)upy[1] + (double)umy[-1]
                  + (double)umy[1] + (double)upx[-1] + (double)upx[1]
                  + (double)umx[-1] + (double)umx[1] + (double)umxmy[0]
                  + (double)upxmy[0] + (double)umxpy[0] + (double)upxpy[0];
    coarse[(long long)f * bc.sizeg + bc.at(ix, iy, iz)]
        = (T)((1. / 64.) * (8. * u0[0] + 4. * face + 2. * edge + corner));
}

// MGkernelExtend3D, src/pb/MGkernels.cc:28-201.  Every fine point depends only
// on injected coarse values, so the two reference passes collapse to a closed
// form per fine point; the T-typed sums keep the reference's operand order.
// Threads cover fine ghosted indices [g, n+g] (the injection also writes the
// first high ghost, MGkernels.cc:65-79).
template <typename T>
__global__ void k_extend3D(Box bf, Box bc, int tiles_z,
    const T* __restrict__ coarse, T* __restrict__ fine)
{
    const int tz = blockIdx.x % tiles_z;
    const int ty = blockIdx.x / tiles_z;
    const int pz = tz * blockDim.x + threadIdx.x; // 0..nz (inclusive)
    const int py = ty * blockDim.y + threadIdx.y;
    const int px = blockIdx.y;
    const int f  = blockIdx.z;
    if (pz > bf.nz || py > bf.ny || px > bf.nx) return;
    const bool ox = px & 1, oy = py & 1, oz = pz & 1;
    const bool ghost = (px == bf.nx) || (py == bf.ny) || (pz == bf.nz);
    if (ghost && (ox || oy || oz)) return; // not touched by the reference
    const long long X = bc.incx, Y = bc.incy;
    // coarse point at or just below this fine point
    const T* c = coarse + (long long)f * bc.sizeg + bc.at(px >> 1, py >> 1, pz >> 1);
    T val;
    if (!ox && !oy && !oz)
        val = c[0];
    else if (!ox && !oy && oz)
        val = (T)(0.5 * (c[0] + c[1]));
    else if (!ox && oy && !oz)
        val = (T)(0.5 * (c[Y] + c[0]));
    else if (!ox && oy && oz)
        val = (T)(0.25 * (c[1 + Y] + c[1] + c[Y] + c[0]));
    else if (ox && !oy && oz)
        val = (T)(0.25 * (c[X + 1] + c[X] + c[1] + c[0]));
    else if (ox && !oy && !oz)
        val = (T)(0.5 * (c[X] + c[0]));
    else if (ox && oy && oz)
        val = (T)(0.125
                  * (c[X + Y + 1] + c[X + Y] + c[X + 1] + c[X] + c[Y + 1] + c[Y]
                      + c[1] + c[0]));
    else
        val = (T)(0.25 * (c[X + Y] + c[X] + c[Y] + c[0]));
    fine[(long long)f * bf.sizeg + bf.at(px, py, pz)] = val;
}

// ---------------------------------------------------------------------------
// Host launchers
// ---------------------------------------------------------------------------
template <typename T>
static int fd_apply_t(int kind, const mgb_grid* gr, const T* v, T* out,
    int nfunc, int rhs_ghosts, cudaStream_t st)
{
    const int g = gr->ghosts;
    Box b       = box_of(gr, g);
    RowLaunch L = row_launch(b.nx, b.ny, b.nz, nfunc);
    switch (kind)
    {
        case MGB_FD_DEL2_4TH_MEHR:
            MGB_REQUIRE(g >= 1, "Mehrstellen needs >= 1 ghost");
            k_del2_4th_mehr<T><<<L.grid, L.block, 0, st>>>(
                b, L.tiles_z, coef_4m(gr->h), v, out);
            MGB_LAUNCHED("k_del2_4th_mehr");
            break;
        case MGB_FD_DEL2_2ND:
            MGB_REQUIRE(g >= 1, "2nd order needs >= 1 ghost");
            k_del2_2nd<T><<<L.grid, L.block, 0, st>>>(
                b, L.tiles_z, coef_cross(2, gr->h), v, out);
            MGB_LAUNCHED("k_del2_2nd");
            break;
        case MGB_FD_DEL2_4TH:
            MGB_REQUIRE(g >= 2, "4th order needs >= 2 ghosts");
            k_del2_cross<T, 2><<<L.grid, L.block, 0, st>>>(
                b, L.tiles_z, coef_cross(4, gr->h), v, out);
            MGB_LAUNCHED("k_del2_4th");
            break;
        case MGB_FD_DEL2_6TH:
            MGB_REQUIRE(g >= 3, "6th order needs >= 3 ghosts");
            k_del2_cross<T, 3><<<L.grid, L.block, 0, st>>>(
                b, L.tiles_z, coef_cross(6, gr->h), v, out);
            MGB_LAUNCHED("k_del2_6th");
            break;
        case MGB_FD_DEL2_8TH:
            MGB_REQUIRE(g >= 4, "8th order needs >= 4 ghosts");
            k_del2_cross<T, 4><<<L.grid, L.block, 0, st>>>(
                b, L.tiles_z, coef_cross(8, gr->h), v, out);
            MGB_LAUNCHED("k_del2_8th");
            break;
        case MGB_FD_RHS_4TH_MEHR1:
        {
            MGB_REQUIRE(g >= 1, "Mehrstellen RHS needs >= 1 ghost");
            MGB_REQUIRE(rhs_ghosts >= 0, "rhs_ghosts < 0");
            Box br = box_of(gr, rhs_ghosts);
            k_rhs_4th_mehr1<T><<<L.grid, L.block, 0, st>>>(
                b, br, L.tiles_z, v, out);
            MGB_LAUNCHED("k_rhs_4th_mehr1");
            break;
        }
        default:
            set_error("mgb_fd_apply: unknown kernel kind %d", kind);
            return MGB_EINVAL;
    }
    return MGB_OK;
}

static dim3 ghosted_grid(const Box& b, int nfunc, dim3& block)
{
    block = dim3(32, 8, 1);
    return dim3((unsigned)((b.nz + 2 * b.g + 31) / 32),
        (unsigned)((b.ny + 2 * b.g + 7) / 8),
        (unsigned)((b.nx + 2 * b.g) * nfunc));
}

template <typename Tin, typename Tout>
static int set_with_ghosts_t(const mgb_grid* gr, const Tin* in, size_t ld,
    Tout* out, int nfunc, cudaStream_t st)
{
    Box b = box_of(gr, gr->ghosts);
    dim3 block;
    // grid.z limit 65535: split functions across several launches if needed
    const int nxg   = b.nx + 2 * b.g;
    const int fstep = 65535 / nxg > 0 ? 65535 / nxg : 1;
    for (int f0 = 0; f0 < nfunc; f0 += fstep)
    {
        const int nf = (nfunc - f0 < fstep) ? nfunc - f0 : fstep;
        dim3 grid    = ghosted_grid(b, nf, block);
        k_set_with_ghosts<Tin, Tout><<<grid, block, 0, st>>>(
            b, in + (size_t)f0 * ld, (long long)ld, out + (size_t)f0 * b.sizeg);
        MGB_LAUNCHED("k_set_with_ghosts");
    }
    return MGB_OK;
}

template <typename Tin, typename Tout>
static int get_values_t(const mgb_grid* gr, const Tin* in, Tout* out,
    size_t ld, int nfunc, cudaStream_t st)
{
    Box b       = box_of(gr, gr->ghosts);
    RowLaunch L = row_launch(b.nx, b.ny, b.nz, nfunc);
    k_get_values<Tin, Tout><<<L.grid, L.block, 0, st>>>(
        b, L.tiles_z, in, out, (long long)ld);
    MGB_LAUNCHED("k_get_values");
    return MGB_OK;
}

// Dirichlet zeroing phase of trade_boundaries (src/pb/GridFuncVector.cc:
// 1554-1557)
template <typename T>
static int trade_dirichlet_t(const mgb_grid* gr, T* u, int nfunc, cudaStream_t st)
{
    const int g = gr->ghosts;
    Box b       = box_of(gr, g);
    BcInfo bc;
    bool any_dir = false;
    for (int d = 0; d < 3; d++)
    {
        bc.dir[d]   = (gr->bc[d] != 1);
        bc.first[d] = (gr->coord[d] == 0);
        bc.last[d]  = (gr->coord[d] == gr->nproc[d] - 1);
        any_dir     = any_dir || bc.dir[d];
    }
    if (!any_dir) return MGB_OK;
    const int nxg   = b.nx + 2 * g;
    const int fstep = 65535 / nxg > 0 ? 65535 / nxg : 1;
    for (int f0 = 0; f0 < nfunc; f0 += fstep)
    {
        const int nf = (nfunc - f0 < fstep) ? nfunc - f0 : fstep;
        dim3 block;
        dim3 grid = ghosted_grid(b, nf, block);
        k_dirichlet_zero<T><<<grid, block, 0, st>>>(b, bc, u + (size_t)f0 * b.sizeg);
        MGB_LAUNCHED("k_dirichlet_zero");
    }
    return MGB_OK;
}

// local periodic wrap of one direction owned by a single rank
template <typename T>
static int trade_wrap_t(const mgb_grid* gr, T* u, int nfunc, int d, cudaStream_t st)
{
    const int g = gr->ghosts;
    Box b       = box_of(gr, g);
    if (!(gr->nproc[d] == 1 && gr->bc[d] == 1)) return MGB_OK;
    long long n  = (d == 1)   ? (long long)b.nx * b.nz
                   : (d == 2) ? (long long)b.nx * (b.ny + 2 * g)
                              : b.incx;
    const int fs = 65535 / (2 * g);
    for (int f0 = 0; f0 < nfunc; f0 += fs)
    {
        const int nf = (nfunc - f0 < fs) ? nfunc - f0 : fs;
        dim3 grid((unsigned)((n + 255) / 256), (unsigned)(2 * g * nf), 1);
        k_wrap<T><<<grid, 256, 0, st>>>(b, d, u + (size_t)f0 * b.sizeg);
        MGB_LAUNCHED("k_wrap");
    }
    return MGB_OK;
}

template <typename T>
static int trade_local_t(
    const mgb_grid* gr, T* u, int nfunc, cudaStream_t st)
{
    const int g = gr->ghosts;
    Box b       = box_of(gr, g);
    if (g == 0) return MGB_OK;
    MGB_REQUIRE(b.nx >= g && b.ny >= g && b.nz >= g,
        "trade_boundaries: local dims smaller than ghost width");
    const bool single[3]
        = { gr->nproc[0] == 1, gr->nproc[1] == 1, gr->nproc[2] == 1 };
    const bool all_periodic_single = single[0] && single[1] && single[2]
                                     && gr->bc[0] == 1 && gr->bc[1] == 1
                                     && gr->bc[2] == 1;
    if (all_periodic_single)
    {
        const int nxg   = b.nx + 2 * g;
        const int fstep = 65535 / nxg > 0 ? 65535 / nxg : 1;
        for (int f0 = 0; f0 < nfunc; f0 += fstep)
        {
            const int nf = (nfunc - f0 < fstep) ? nfunc - f0 : fstep;
            dim3 block;
            dim3 grid = ghosted_grid(b, nf, block);
            k_fill_ghosts_periodic<T><<<grid, block, 0, st>>>(
                b, u + (size_t)f0 * b.sizeg);
            MGB_LAUNCHED("k_fill_ghosts_periodic");
        }
        return MGB_OK;
    }
    if (int rc = trade_dirichlet_t<T>(gr, u, nfunc, st)) return rc;
    const int order[3] = { 1, 2, 0 }; // Y, Z, X
    for (int o = 0; o < 3; o++)
        if (int rc = trade_wrap_t<T>(gr, u, nfunc, order[o], st)) return rc;
    return MGB_OK;
}

// copy a sub-box [lo, lo+ext) (ghosted coordinates) of every function between
// a ghosted block and a packed buffer [nfunc][ext0][ext1][ext2]
template <typename T, bool PACK>
__global__ void k_subbox(Box b, int lo0, int lo1, int lo2, int e0, int e1, int e2,
    T* __restrict__ u, T* __restrict__ buf)
{
    const long long per = (long long)e0 * e1 * e2;
    const long long t   = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= per) return;
    const int f  = blockIdx.y;
    const int k  = (int)(t % e2);
    const int j  = (int)((t / e2) % e1);
    const int i  = (int)(t / ((long long)e2 * e1));
    T* pu = u + (long long)f * b.sizeg + (long long)(lo0 + i) * b.incx
            + (long long)(lo1 + j) * b.incy + (lo2 + k);
    T* pb = buf + (long long)f * per + t;
    if (PACK)
        *pb = *pu;
    else
        *pu = *pb;
}

int subbox_copy(int dtype, bool pack, const mgb_grid* gr, const int lo[3],
    const int ext[3], void* u, void* buf, int nfunc, cudaStream_t st)
{
    Box b = box_of(gr, gr->ghosts);
    const long long per = (long long)ext[0] * ext[1] * ext[2];
    if (per == 0 || nfunc == 0) return MGB_OK;
    for (int f0 = 0; f0 < nfunc; f0 += 65535)
    {
        const int nf = (nfunc - f0 < 65535) ? nfunc - f0 : 65535;
        dim3 grid((unsigned)((per + 255) / 256), (unsigned)nf, 1);
        const size_t es = dtype == MGB_F64 ? 8 : 4;
        void* uu = (char*)u + (size_t)f0 * b.sizeg * es;
        void* bb = (char*)buf + (size_t)f0 * per * es;
        if (dtype == MGB_F64)
        {
            if (pack)
                k_subbox<double, true><<<grid, 256, 0, st>>>(b, lo[0], lo[1], lo[2],
                    ext[0], ext[1], ext[2], (double*)uu, (double*)bb);
            else
                k_subbox<double, false><<<grid, 256, 0, st>>>(b, lo[0], lo[1], lo[2],
                    ext[0], ext[1], ext[2], (double*)uu, (double*)bb);
        }
        else
        {
            if (pack)
                k_subbox<float, true><<<grid, 256, 0, st>>>(b, lo[0], lo[1], lo[2],
                    ext[0], ext[1], ext[2], (float*)uu, (float*)bb);
            else
                k_subbox<float, false><<<grid, 256, 0, st>>>(b, lo[0], lo[1], lo[2],
                    ext[0], ext[1], ext[2], (float*)uu, (float*)bb);
        }
        MGB_LAUNCHED("k_subbox");
    }
    return MGB_OK;
}

int trade_dirichlet(int dtype, const mgb_grid* gr, void* u, int nfunc, cudaStream_t st)
{
    return dtype == MGB_F64 ? trade_dirichlet_t<double>(gr, (double*)u, nfunc, st)
                            : trade_dirichlet_t<float>(gr, (float*)u, nfunc, st);
}
int trade_wrap(int dtype, const mgb_grid* gr, void* u, int nfunc, int d, cudaStream_t st)
{
    return dtype == MGB_F64 ? trade_wrap_t<double>(gr, (double*)u, nfunc, d, st)
                            : trade_wrap_t<float>(gr, (float*)u, nfunc, d, st);
}

template <typename T>
static int transfer_t(bool restrict_, const mgb_grid* fine, const T* src,
    T* dst, int nfunc, cudaStream_t st)
{
    const int g = fine->ghosts;
    MGB_REQUIRE(g >= 1, "MG transfer needs >= 1 ghost");
    MGB_REQUIRE(fine->dim[0] % 2 == 0 && fine->dim[1] % 2 == 0
                    && fine->dim[2] % 2 == 0,
        "MG transfer: fine dims must be even");
    Box bf = box_of(fine, g);
    Box bc(bf.nx / 2, bf.ny / 2, bf.nz / 2, g);
    if (restrict_)
    {
        RowLaunch L = row_launch(bc.nx, bc.ny, bc.nz, nfunc);
        k_restrict3D<T><<<L.grid, L.block, 0, st>>>(bf, bc, L.tiles_z, src, dst);
        MGB_LAUNCHED("k_restrict3D");
    }
    else
    {
        RowLaunch L = row_launch(bf.nx + 1, bf.ny + 1, bf.nz + 1, nfunc);
        k_extend3D<T><<<L.grid, L.block, 0, st>>>(bf, bc, L.tiles_z, src, dst);
        MGB_LAUNCHED("k_extend3D");
    }
    return MGB_OK;
}

static unsigned flat_blocks(size_t n)
{
    size_t nb = (n + 255) / 256;
    if (nb > 148 * 32) nb = 148 * 32; // grid-stride: 32 CTAs per SM
    if (nb == 0) nb = 1;
    return (unsigned)nb;
}

// internal entry points used by other translation units
int del2_4th_withpot(int dtype, const mgb_grid* gr, const void* v,
    const double* pot, void* out, size_t ldo, int nfunc, cudaStream_t st)
{
    MGB_REQUIRE(gr->ghosts >= 2, "del2_4th_withPot needs >= 2 ghosts");
    Box b       = box_of(gr, gr->ghosts);
    RowLaunch L = row_launch(b.nx, b.ny, b.nz, nfunc);
    if (dtype == MGB_F64)
        k_del2_4th_withpot<double><<<L.grid, L.block, 0, st>>>(b, L.tiles_z,
            coef_cross(4, gr->h), (const double*)v, pot, (double*)out,
            (long long)ldo);
    else
        k_del2_4th_withpot<float><<<L.grid, L.block, 0, st>>>(b, L.tiles_z,
            coef_cross(4, gr->h), (const float*)v, pot, (float*)out,
            (long long)ldo);
    MGB_LAUNCHED("k_del2_4th_withpot");
    return MGB_OK;
}

int jacobi_literal(int lap_type, const mgb_grid* gr, float* v, const float* f,
    float* w, int nfunc, double omega, cudaStream_t st)
{
    int kind;
    switch (lap_type)
    {
        case MGB_LAP_4M:
            kind = MGB_FD_DEL2_4TH_MEHR;
            break;
        case MGB_LAP_2:
            kind = MGB_FD_DEL2_2ND;
            break;
        case MGB_LAP_4:
            kind = MGB_FD_DEL2_4TH;
            break;
        case MGB_LAP_6:
            kind = MGB_FD_DEL2_6TH;
            break;
        case MGB_LAP_8:
            kind = MGB_FD_DEL2_8TH;
            break;
        default:
            // GridFuncVector::applyLap aborts (src/pb/GridFuncVector.cc:2392)
            set_error("jacobi: lap_type %d has no batched applyLap", lap_type);
            return MGB_EINVAL;
    }
    int rc = fd_apply_t<float>(kind, gr, v, w, nfunc, 0, st);
    if (rc) return rc;
    const size_t n = (size_t)box_of(gr, gr->ghosts).sizeg * nfunc;
    k_jacobi_tail<<<flat_blocks(n), 256, 0, st>>>(n, -1. * omega, f, w, v);
    MGB_LAUNCHED("k_jacobi_tail");
    return MGB_OK;
}

} // namespace mgb

using namespace mgb;

extern "C"
{

int mgb_fd_apply(int kind, int dtype, const mgb_grid* grid, const void* v,
    void* out, int nfunc, int rhs_ghosts, void* stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(grid)) return rc;
    MGB_REQUIRE(v && out && nfunc >= 0, "mgb_fd_apply: null pointer");
    if (nfunc == 0) return MGB_OK;
    if (dtype == MGB_F64)
        return fd_apply_t<double>(kind, grid, (const double*)v, (double*)out,
            nfunc, rhs_ghosts, as_stream(stream));
    if (dtype == MGB_F32)
        return fd_apply_t<float>(kind, grid, (const float*)v, (float*)out,
            nfunc, rhs_ghosts, as_stream(stream));
    set_error("mgb_fd_apply: bad dtype %d", dtype);
    return MGB_EINVAL;
}

int mgb_gfv_set_with_ghosts(int in_dtype, int out_dtype, const mgb_grid* grid,
    const void* noghost, size_t ld, void* ghosted, int nfunc, void* stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(grid)) return rc;
    MGB_REQUIRE(noghost && ghosted, "mgb_gfv_set_with_ghosts: null pointer");
    MGB_REQUIRE(ld >= (size_t)box_of(grid, 0).npt, "ld < npt");
    if (nfunc == 0) return MGB_OK;
    cudaStream_t st = as_stream(stream);
    if (in_dtype == MGB_F64 && out_dtype == MGB_F64)
        return set_with_ghosts_t(grid, (const double*)noghost, ld,
            (double*)ghosted, nfunc, st);
    if (in_dtype == MGB_F64 && out_dtype == MGB_F32)
        return set_with_ghosts_t(
            grid, (const double*)noghost, ld, (float*)ghosted, nfunc, st);
    if (in_dtype == MGB_F32 && out_dtype == MGB_F32)
        return set_with_ghosts_t(
            grid, (const float*)noghost, ld, (float*)ghosted, nfunc, st);
    if (in_dtype == MGB_F32 && out_dtype == MGB_F64)
        return set_with_ghosts_t(
            grid, (const float*)noghost, ld, (double*)ghosted, nfunc, st);
    set_error("mgb_gfv_set_with_ghosts: bad dtype");
    return MGB_EINVAL;
}

int mgb_gfv_get_values(int in_dtype, int out_dtype, const mgb_grid* grid,
    const void* ghosted, void* noghost, size_t ld, int nfunc, void* stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(grid)) return rc;
    MGB_REQUIRE(noghost && ghosted, "mgb_gfv_get_values: null pointer");
    MGB_REQUIRE(ld >= (size_t)box_of(grid, 0).npt, "ld < npt");
    if (nfunc == 0) return MGB_OK;
    cudaStream_t st = as_stream(stream);
    if (in_dtype == MGB_F64 && out_dtype == MGB_F64)
        return get_values_t(
            grid, (const double*)ghosted, (double*)noghost, ld, nfunc, st);
    if (in_dtype == MGB_F64 && out_dtype == MGB_F32)
        return get_values_t(
            grid, (const double*)ghosted, (float*)noghost, ld, nfunc, st);
    if (in_dtype == MGB_F32 && out_dtype == MGB_F32)
        return get_values_t(
            grid, (const float*)ghosted, (float*)noghost, ld, nfunc, st);
    if (in_dtype == MGB_F32 && out_dtype == MGB_F64)
        return get_values_t(
            grid, (const float*)ghosted, (double*)noghost, ld, nfunc, st);
    set_error("mgb_gfv_get_values: bad dtype");
    return MGB_EINVAL;
}

int mgb_gfv_trade_boundaries(
    int dtype, const mgb_grid* grid, void* ghosted, int nfunc, void* stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(grid)) return rc;
    MGB_REQUIRE(ghosted, "mgb_gfv_trade_boundaries: null pointer");
    if (nfunc == 0) return MGB_OK;
    if (dtype == MGB_F64)
        return trade_local_t<double>(
            grid, (double*)ghosted, nfunc, as_stream(stream));
    if (dtype == MGB_F32)
        return trade_local_t<float>(
            grid, (float*)ghosted, nfunc, as_stream(stream));
    set_error("mgb_gfv_trade_boundaries: bad dtype");
    return MGB_EINVAL;
}

int mgb_gfv_pointwise_product(int dtype, const mgb_grid* grid, const void* A,
    const double* Vghost, void* out, int nfunc, void* stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(grid)) return rc;
    MGB_REQUIRE(A && Vghost && out, "mgb_gfv_pointwise_product: null pointer");
    if (nfunc == 0) return MGB_OK;
    Box b = box_of(grid, grid->ghosts);
    dim3 g((unsigned)((b.sizeg + 255) / 256), (unsigned)(nfunc < 64 ? nfunc : 64));
    if (dtype == MGB_F64)
        k_pointwise_product<double><<<g, 256, 0, as_stream(stream)>>>(
            b.sizeg, nfunc, (const double*)A, Vghost, (double*)out);
    else if (dtype == MGB_F32)
        k_pointwise_product<float><<<g, 256, 0, as_stream(stream)>>>(
            b.sizeg, nfunc, (const float*)A, Vghost, (float*)out);
    else
    {
        set_error("mgb_gfv_pointwise_product: bad dtype");
        return MGB_EINVAL;
    }
    MGB_LAUNCHED("k_pointwise_product");
    return MGB_OK;
}

int mgb_axpy(
    int dtype, size_t n, double alpha, const void* x, void* y, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(x && y, "mgb_axpy: null pointer");
    if (n == 0) return MGB_OK;
    if (dtype == MGB_F64)
        k_axpy<double><<<flat_blocks(n), 256, 0, as_stream(stream)>>>(
            n, alpha, (const double*)x, (double*)y);
    else if (dtype == MGB_F32)
        k_axpy<float><<<flat_blocks(n), 256, 0, as_stream(stream)>>>(
            n, alpha, (const float*)x, (float*)y);
    else
    {
        set_error("mgb_axpy: bad dtype");
        return MGB_EINVAL;
    }
    MGB_LAUNCHED("k_axpy");
    return MGB_OK;
}

int mgb_scal(int dtype, size_t n, double alpha, void* x, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(x, "mgb_scal: null pointer");
    if (n == 0 || alpha == 1.) return MGB_OK;
    if (dtype == MGB_F64)
        k_scal<double><<<flat_blocks(n), 256, 0, as_stream(stream)>>>(
            n, alpha, (double*)x);
    else if (dtype == MGB_F32)
        k_scal<float><<<flat_blocks(n), 256, 0, as_stream(stream)>>>(
            n, alpha, (float*)x);
    else
    {
        set_error("mgb_scal: bad dtype");
        return MGB_EINVAL;
    }
    MGB_LAUNCHED("k_scal");
    return MGB_OK;
}

int mgb_dot(int dtype, size_t n, const void* x, const void* y,
    double* result_dev, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(x && y && result_dev, "mgb_dot: null pointer");
    cudaStream_t st = as_stream(stream);
    static double* partial = nullptr; // 148*8 doubles, allocated once
    const int nb_max       = 148 * 8;
    if (!partial) MGB_CUDA(cudaMalloc(&partial, sizeof(double) * nb_max));
    int nb = (int)((n + 255) / 256);
    if (nb > nb_max) nb = nb_max;
    if (nb < 1) nb = 1;
    if (dtype == MGB_F64)
        k_dot_partial<double>
            <<<nb, 256, 0, st>>>(n, (const double*)x, (const double*)y, partial);
    else if (dtype == MGB_F32)
        k_dot_partial<float>
            <<<nb, 256, 0, st>>>(n, (const float*)x, (const float*)y, partial);
    else
    {
        set_error("mgb_dot: bad dtype");
        return MGB_EINVAL;
    }
    MGB_LAUNCHED("k_dot_partial");
    k_dot_final<<<1, 256, 0, st>>>(nb, partial, result_dev);
    MGB_LAUNCHED("k_dot_final");
    return MGB_OK;
}

int mgb_gfv_jacobi(int lap_type, const mgb_grid* grid, float* v, const float* f,
    float* w, int nfunc, double omega, void* stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(grid)) return rc;
    MGB_REQUIRE(v && f && w, "mgb_gfv_jacobi: null pointer");
    if (nfunc == 0) return MGB_OK;
    return jacobi_literal(lap_type, grid, v, f, w, nfunc, omega, as_stream(stream));
}

int mgb_gfv_restrict3D(int dtype, const mgb_grid* fine, const void* ufine,
    void* ucoarse, int nfunc, void* stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(fine)) return rc;
    MGB_REQUIRE(ufine && ucoarse, "mgb_gfv_restrict3D: null pointer");
    if (nfunc == 0) return MGB_OK;
    if (dtype == MGB_F64)
        return transfer_t<double>(true, fine, (const double*)ufine,
            (double*)ucoarse, nfunc, as_stream(stream));
    if (dtype == MGB_F32)
        return transfer_t<float>(true, fine, (const float*)ufine,
            (float*)ucoarse, nfunc, as_stream(stream));
    set_error("mgb_gfv_restrict3D: bad dtype");
    return MGB_EINVAL;
}

int mgb_gfv_extend3D(int dtype, const mgb_grid* fine, const void* ucoarse,
    void* ufine, int nfunc, void* stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(fine)) return rc;
    MGB_REQUIRE(ufine && ucoarse, "mgb_gfv_extend3D: null pointer");
    if (nfunc == 0) return MGB_OK;
    if (dtype == MGB_F64)
        return transfer_t<double>(false, fine, (const double*)ucoarse,
            (double*)ufine, nfunc, as_stream(stream));
    if (dtype == MGB_F32)
        return transfer_t<float>(false, fine, (const float*)ucoarse,
            (float*)ufine, nfunc, as_stream(stream));
    set_error("mgb_gfv_extend3D: bad dtype");
    return MGB_EINVAL;
}

} // extern "C"
