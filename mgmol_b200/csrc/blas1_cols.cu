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

// ---------------------------------------------------------------------------
// Electron density from non-orthogonal orbitals:
// Rho::computeRhoSubdomainUsingBlas3 (src/Rho.cc:359-448)
//     product = Phi1 X            (MPgemmNN, ORBDTYPE)
//     rho[i] += product[j][i] * phi2[j][i]   for j = 0 .. ncols-1, in this order
// for one x-slab of nrows points.  The contraction runs on the tensor path in
// point chunks through a library workspace; the accumulation is one pass over
// the chunk (product and phi2 read once), the j loop sequential per point as in
// the reference (so rho is bit-identical given the same product).
// ---------------------------------------------------------------------------
namespace mgb
{
int gemm_nn_fused(int dtype, size_t m, int n, int k, double alpha, const void* A, size_t lda,
    const double* M, int ldm, double beta, void* Out, size_t ldc, double gamma, const void* D,
    size_t ldd, cudaStream_t st);

template <typename T>
__global__ void k_rho_accumulate(long long n, int nfunc, const T* __restrict__ product,
    long long ldp, const T* __restrict__ phi2, long long ld2, double* __restrict__ rho)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = rho[i];
    for (int j = 0; j < nfunc; j++)
        s += (double)(product[(long long)j * ldp + i] * phi2[(long long)j * ld2 + i]);
    rho[i] = s;
}
} // namespace mgb

extern "C" int mgb_rho_blas3(int dtype, size_t nrows, int nfunc, const void* phi1, size_t ld1,
    const double* X, int ldx, const void* phi2, size_t ld2, double* rho, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(phi1 && phi2 && X && rho, "mgb_rho_blas3: null pointer");
    MGB_REQUIRE(dtype == MGB_F32 || dtype == MGB_F64, "mgb_rho_blas3: bad dtype");
    MGB_REQUIRE(nfunc >= 0 && ld1 >= nrows && ld2 >= nrows && ldx >= nfunc,
        "mgb_rho_blas3: bad dimensions");
    if (nfunc == 0 || nrows == 0) return MGB_OK;
    cudaStream_t st = as_stream(stream);
    const size_t es = dtype == MGB_F64 ? 8 : 4;
    // point chunks of <= 2^20 points (a multiple of 128: whole tiles, aligned)
    size_t chunk = (size_t)1 << 20;
    if (chunk > nrows) chunk = (nrows + 127) / 128 * 128;
    void* product = scratch(9, chunk * (size_t)nfunc * es);
    if (!product) return MGB_ECUDA;
    for (size_t p0 = 0; p0 < nrows; p0 += chunk)
    {
        const size_t m = (nrows - p0 < chunk) ? nrows - p0 : chunk;
        if (int rc = gemm_nn_fused(dtype, m, nfunc, nfunc, 1., (const char*)phi1 + p0 * es, ld1, X,
                ldx, 0., product, chunk, 0., nullptr, 0, st))
            return rc;
        const unsigned blocks = (unsigned)((m + 255) / 256);
        if (dtype == MGB_F64)
            k_rho_accumulate<double><<<blocks, 256, 0, st>>>((long long)m, nfunc,
                (const double*)product, (long long)chunk, (const double*)phi2 + p0,
                (long long)ld2, rho + p0);
        else
            k_rho_accumulate<float><<<blocks, 256, 0, st>>>((long long)m, nfunc,
                (const float*)product, (long long)chunk, (const float*)phi2 + p0, (long long)ld2,
                rho + p0);
        MGB_LAUNCHED("k_rho_accumulate");
    }
    return MGB_OK;
}
