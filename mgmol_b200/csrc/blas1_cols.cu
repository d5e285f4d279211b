// Per-orbital BLAS-1 on no-ghost blocks: the dot products
// of ExtendedGridOrbitals::computeDiagonalElementsDotProduct
// (src/ExtendedGridOrbitals.cc:1085-1106: ss[i] = vel * MPdot(phi_i, psi_i)),
// which feed Orbitals::dotProduct (dotProductDiagonal, :1205-1213) and through
// it the residual norm and AndersonMix (src/AndersonMix.cc:88-110).  One launch
// for all orbitals, double accumulation of double products (MPdot,
// src/linear_algebra/mputils.cc:135-166), fixed two-pass reduction order.
#include "common.cuh"

namespace mgb
{

constexpr int kDotBlocks = 32; // partial sums per orbital

template <typename T>
__global__ void k_dot_cols_partial(long long n, const T* __restrict__ x, long long ldx,
    const T* __restrict__ y, long long ldy, double* __restrict__ partial)
{
    __shared__ double sm[32];
    const int f = blockIdx.y;
    const T* px = x + (long long)f * ldx;
    const T* py = y + (long long)f * ldy;
    double s = 0.;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        s += (double)px[i] * (double)py[i];
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32)
    {
        s = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.;
        for (int o = 16; o > 0; o >>= 1)
            s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) partial[(long long)f * gridDim.x + blockIdx.x] = s;
    }
}

__global__ void k_dot_cols_final(int nblocks, int nfunc, double alpha,
    const double* __restrict__ partial, double* __restrict__ out)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfunc) return;
    double s = 0.;
    for (int b = 0; b < nblocks; b++)
        s += partial[(long long)f * nblocks + b];
    out[f] = alpha * s;
}

void* scratch(int slot, size_t bytes);

template <typename T>
static int dot_cols_t(size_t n, int nfunc, double alpha, const T* x, size_t ldx, const T* y,
    size_t ldy, double* out, cudaStream_t st)
{
    double* partial = (double*)scratch(7, sizeof(double) * kDotBlocks * (size_t)nfunc);
    if (!partial) return MGB_ECUDA;
    for (int f0 = 0; f0 < nfunc; f0 += 65535)
    {
        const int nf = (nfunc - f0 < 65535) ? nfunc - f0 : 65535;
        dim3 grid(kDotBlocks, (unsigned)nf);
        k_dot_cols_partial<T><<<grid, 256, 0, st>>>((long long)n, x + (size_t)f0 * ldx,
            (long long)ldx, y + (size_t)f0 * ldy, (long long)ldy,
            partial + (size_t)f0 * kDotBlocks);
        MGB_LAUNCHED("k_dot_cols_partial");
    }
    k_dot_cols_final<<<(nfunc + 127) / 128, 128, 0, st>>>(kDotBlocks, nfunc, alpha, partial, out);
    MGB_LAUNCHED("k_dot_cols_final");
    return MGB_OK;
}

} // namespace mgb

using namespace mgb;

extern "C" int mgb_dot_cols(int dtype, size_t n, int nfunc, double alpha, const void* x,
    size_t ldx, const void* y, size_t ldy, double* result_dev, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(x && y && result_dev, "mgb_dot_cols: null pointer");
    MGB_REQUIRE(nfunc >= 0 && ldx >= n && ldy >= n, "mgb_dot_cols: bad dimensions");
    if (nfunc == 0) return MGB_OK;
    if (dtype == MGB_F64)
        return dot_cols_t<double>(n, nfunc, alpha, (const double*)x, ldx, (const double*)y, ldy,
            result_dev, as_stream(stream));
    if (dtype == MGB_F32)
        return dot_cols_t<float>(n, nfunc, alpha, (const float*)x, ldx, (const float*)y, ldy,
            result_dev, as_stream(stream));
    set_error("mgb_dot_cols: bad dtype");
    return MGB_EINVAL;
}
