// Dense orbital contractions on the FP64 tensor pipe (DMMA m8n8k4):
//   Gram        S      = alpha * Phi^T Phi          (mgb_syrk_t)
//   projection  Phi^T A                              (mgb_gemm_tn)
//   mixing      Phi * M                              (mgb_gemm_nn)
// as LinearAlgebraUtils::MPsyrk / MPgemm('T','N') / MPgemmNN define them
// (src/linear_algebra/mputils.cc:295-1067): ORBDTYPE operands, products and
// accumulation in double (float operands are widened when fragments are
// loaded, exactly the (double)a*(double)b of MPdot, :147-166), one rounding to
// the output type at the end.
//
// B200 has no tcgen05 FP64 kind; DMMA.8x8x4 is the FP64 tensor instruction
// (mma.sync m16n8k16.f64 lowers to it).  Kernel shape: 128x128 CTA tile, 16
// warps of 32x32, K staged through a 4-deep cp.async ring of 16-wide slabs
// whose rows are padded to 20 elements (fragment loads conflict-free).
// Tall-skinny shapes (K = grid points, M = N = orbitals) are split along K;
// partial tiles are combined by a second kernel in a fixed order, so results
// are run-to-run deterministic.
#include <cstdint>

#include "hpsi.h"

namespace mgb
{

constexpr int BM = 128, BN = 128, KC = 16, PADK = 4, STAGES = 4;
constexpr int PITCH_K = KC + PADK; // elements
constexpr int NTHREADS = 512;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem),
        "r"(src_bytes)
        : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
    asm volatile(
        "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c[0]), "+d"(c[1])
        : "d"(a), "d"(b));
}

// load a [128 rows][KC] slab of a K-major operand (row r at base + r*ld, K
// contiguous) into smem rows of PITCH_K elements; rows >= nrows and K >= kend
// are zero-filled.
template <typename T>
__device__ __forceinline__ void load_kmajor(T* sm, const T* base, long long ld,
    int row0, int nrows, long long k0, long long kend, int tid)
{
    constexpr int CH  = 16 / (int)sizeof(T);      // elements per 16B chunk
    constexpr int CPR = KC / CH;                   // chunks per row
    constexpr int TOT = 128 * CPR;
#pragma unroll
    for (int c = tid; c < TOT; c += NTHREADS)
    {
        const int r  = c / CPR;
        const int cc = c % CPR;
        const long long k = k0 + (long long)cc * CH;
        int valid = 0;
        const int row = row0 + r;
        if (row < nrows && k < kend)
        {
            const long long rem = kend - k;
            valid = rem >= CH ? 16 : (int)rem * (int)sizeof(T);
        }
        const T* src = base + (long long)(row < nrows ? row : 0) * ld + (valid ? k : 0);
        cp_async16(sm + r * PITCH_K + cc * CH, src, valid);
    }
}

// C(m x n) (+)= A^T B over K range of this CTA.  A rows = m index, B rows = n.
template <typename T, bool SYRK>
__global__ void __launch_bounds__(NTHREADS, 1) k_gemm_tn(int m, int n, long long k,
    long long kper, const T* __restrict__ A, long long lda, const T* __restrict__ B,
    long long ldb, double alpha, double beta, double* __restrict__ C, int ldc,
    double* __restrict__ partial, int nsplit)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    T* As = reinterpret_cast<T*>(smraw);
    T* Bs = As + STAGES * 128 * PITCH_K;

    const int tile_n = blockIdx.x, tile_m = blockIdx.y, split = blockIdx.z;
    if (SYRK && tile_n > tile_m) return; // lower triangle of tiles only
    const int m0 = tile_m * BM, n0 = tile_n * BN;
    const long long kb = (long long)split * kper;
    const long long ke = (kb + kper < k) ? kb + kper : k;
    const int nkt      = (int)((ke - kb + KC - 1) / KC);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3; // 4 x 4 warps, 32 x 32 each
    const int fr = lane >> 2, fc = lane & 3;

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
            acc[i][j][0] = acc[i][j][1] = 0.;

    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++)
    {
        if (s < nkt)
        {
            load_kmajor<T>(As + s * 128 * PITCH_K, A, lda, m0, m, kb + (long long)s * KC, ke, tid);
            load_kmajor<T>(Bs + s * 128 * PITCH_K, B, ldb, n0, n, kb + (long long)s * KC, ke, tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < nkt; kt++)
    {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nx = kt + STAGES - 1;
            if (nx < nkt)
            {
                const int s = nx % STAGES;
                load_kmajor<T>(As + s * 128 * PITCH_K, A, lda, m0, m, kb + (long long)nx * KC, ke, tid);
                load_kmajor<T>(Bs + s * 128 * PITCH_K, B, ldb, n0, n, kb + (long long)nx * KC, ke, tid);
            }
            cp_async_commit();
        }
        const T* as = As + (kt % STAGES) * 128 * PITCH_K + (wm * 32 + fr) * PITCH_K + fc;
        const T* bs = Bs + (kt % STAGES) * 128 * PITCH_K + (wn * 32 + fr) * PITCH_K + fc;
#pragma unroll
        for (int kk = 0; kk < KC / 4; kk++)
        {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                a[i] = (double)as[i * 8 * PITCH_K + kk * 4];
                b[i] = (double)bs[i * 8 * PITCH_K + kk * 4];
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
                    dmma(acc[i][j], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: thread holds C(m0 + wm*32 + i*8 + fr, n0 + wn*32 + j*8 + fc*2 + e)
    double* dst   = (nsplit > 1) ? partial + (size_t)split * (size_t)m * n : C;
    const int ldd = (nsplit > 1) ? m : ldc;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++)
            {
                const int mm = m0 + wm * 32 + i * 8 + fr;
                const int nn = n0 + wn * 32 + j * 8 + fc * 2 + e;
                if (mm < m && nn < n)
                {
                    if (nsplit > 1)
                        dst[(size_t)nn * ldd + mm] = acc[i][j][e];
                    else
                    {
                        const double old
                            = (beta == 0.) ? 0. : beta * C[(size_t)nn * ldc + mm];
                        const double val = alpha * acc[i][j][e] + old;
                        C[(size_t)nn * ldc + mm] = val;
                        // tiles above the tile diagonal are filled by mirroring
                        if (SYRK && tile_m > tile_n) C[(size_t)mm * ldc + nn] = val;
                    }
                }
            }
}

// C = alpha * sum_s partial[s] + beta*C in a fixed order; SYRK mirrors the
// lower triangle (LocalMatrices::fillUpperWithLower)
template <bool SYRK>
__global__ void k_splitk_reduce(int m, int n, int nsplit, const double* __restrict__ partial,
    double alpha, double beta, double* __restrict__ C, int ldc)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)m * n) return;
    const int mm = (int)(t % m), nn = (int)(t / m);
    if (SYRK)
    {
        // tiles strictly above the tile diagonal were not computed
        if (nn / BN > mm / BM) return;
    }
    double s = 0.;
    for (int sp = 0; sp < nsplit; sp++)
        s += partial[(size_t)sp * m * n + (size_t)nn * m + mm];
    const double old = (beta == 0.) ? 0. : beta * C[(size_t)nn * ldc + mm];
    const double val = alpha * s + old;
    C[(size_t)nn * ldc + mm] = val;
    if (SYRK && mm / BM > nn / BN) C[(size_t)mm * ldc + nn] = val;
}

// Out(n x npt) = alpha * M^T Phi + beta*Out : rows of Out/Phi are orbitals
// (points contiguous).  A' = M^T tile [128 j][KC l] (K-major: M is column-major
// l + j*ldm), B' = Phi tile [KC l][128 p] (N-major).
constexpr int PITCH_P = 128 + 8;

template <typename T>
__global__ void __launch_bounds__(NTHREADS, 1) k_gemm_nn(long long npt, int n, int k,
    const T* __restrict__ Phi, long long lda, const double* __restrict__ M, int ldm,
    double alpha, double beta, T* __restrict__ Out, long long ldc)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    double* Ms = reinterpret_cast<double*>(smraw);               // [STAGES][128][PITCH_K]
    T* Ps      = reinterpret_cast<T*>(Ms + STAGES * 128 * PITCH_K); // [STAGES][KC][PITCH_P]

    const int tile_j = blockIdx.x;
    const long long p0 = (long long)blockIdx.y * 128;
    const int j0  = tile_j * 128;
    const int nkt = (k + KC - 1) / KC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int fr = lane >> 2, fc = lane & 3;

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
            acc[i][j][0] = acc[i][j][1] = 0.;

    auto load_phi = [&](int s, int kt) {
        constexpr int CH  = 16 / (int)sizeof(T);
        constexpr int CPR = 128 / CH; // chunks per row of 128 points
        constexpr int TOT = KC * CPR;
        T* sm = Ps + s * KC * PITCH_P;
#pragma unroll
        for (int c = tid; c < TOT; c += NTHREADS)
        {
            const int r  = c / CPR;
            const int cc = c % CPR;
            const int l  = kt * KC + r;
            const long long p = p0 + (long long)cc * CH;
            int valid = 0;
            if (l < k && p < npt)
            {
                const long long rem = npt - p;
                valid = rem >= CH ? 16 : (int)rem * (int)sizeof(T);
            }
            const T* src = Phi + (long long)(l < k ? l : 0) * lda + (valid ? p : 0);
            cp_async16(sm + r * PITCH_P + cc * CH, src, valid);
        }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++)
    {
        if (s < nkt)
        {
            load_kmajor<double>(Ms + s * 128 * PITCH_K, M, ldm, j0, n, (long long)s * KC, k, tid);
            load_phi(s, s);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < nkt; kt++)
    {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nx = kt + STAGES - 1;
            if (nx < nkt)
            {
                const int s = nx % STAGES;
                load_kmajor<double>(Ms + s * 128 * PITCH_K, M, ldm, j0, n, (long long)nx * KC, k, tid);
                load_phi(s, nx);
            }
            cp_async_commit();
        }
        const double* as = Ms + (kt % STAGES) * 128 * PITCH_K + (wm * 32 + fr) * PITCH_K + fc;
        const T* bs      = Ps + (kt % STAGES) * KC * PITCH_P + fc * PITCH_P + wn * 32 + fr;
#pragma unroll
        for (int kk = 0; kk < KC / 4; kk++)
        {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                a[i] = as[i * 8 * PITCH_K + kk * 4];
                b[i] = (double)bs[kk * 4 * PITCH_P + i * 8];
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
                    dmma(acc[i][j], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // thread holds Out(j0 + wm*32 + i*8 + fr, p0 + wn*32 + j*8 + fc*2 + {0,1})
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        const int jj = j0 + wm * 32 + i * 8 + fr;
        if (jj >= n) continue;
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            const long long pp = p0 + wn * 32 + j * 8 + fc * 2;
            T* o = Out + (long long)jj * ldc + pp;
#pragma unroll
            for (int e = 0; e < 2; e++)
                if (pp + e < npt)
                {
                    // MPscal(beta) then += (T)buff  (mputils.cc:750-759)
                    T base = (beta == 0.) ? (T)0 : (T)(beta * (double)o[e]);
                    o[e]   = base + (T)(alpha * acc[i][j][e]);
                }
        }
    }
}

// slow, always-applicable fallbacks (unaligned operands): one thread per
// output element, sequential K like MPdot
template <typename T>
__global__ void k_gemm_tn_ref(int m, int n, long long k, const T* A, long long lda,
    const T* B, long long ldb, double alpha, double beta, double* C, int ldc, int sym)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)m * n) return;
    const int mm = (int)(t % m), nn = (int)(t / m);
    if (sym && nn > mm) return;
    double s = 0.;
    for (long long l = 0; l < k; l++)
        s += (double)A[(long long)mm * lda + l] * (double)B[(long long)nn * ldb + l];
    const double old = (beta == 0.) ? 0. : beta * C[(size_t)nn * ldc + mm];
    C[(size_t)nn * ldc + mm] = alpha * s + old;
    if (sym && mm != nn) C[(size_t)mm * ldc + nn] = alpha * s + old;
}
template <typename T>
__global__ void k_gemm_nn_ref(long long npt, int n, int k, const T* Phi, long long lda,
    const double* M, int ldm, double alpha, double beta, T* Out, long long ldc)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int j       = blockIdx.y;
    if (p >= npt) return;
    double s = 0.;
    for (int l = 0; l < k; l++)
        s += (alpha * M[(size_t)j * ldm + l]) * (double)Phi[(long long)l * lda + p];
    T* o   = Out + (long long)j * ldc + p;
    T base = (beta == 0.) ? (T)0 : (T)(beta * (double)*o);
    *o     = base + (T)s;
}

template <typename T>
static int gemm_tn_t(bool syrk, int m, int n, size_t k, double alpha, const T* A,
    size_t lda, const T* B, size_t ldb, double beta, double* C, int ldc, cudaStream_t st)
{
    const bool aligned = (((uintptr_t)A | (uintptr_t)B) & 15) == 0
                         && (lda * sizeof(T)) % 16 == 0 && (ldb * sizeof(T)) % 16 == 0;
    if (!aligned || k < 64)
    {
        const long long tot = (long long)m * n;
        k_gemm_tn_ref<T><<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(m, n,
            (long long)k, A, (long long)lda, B, (long long)ldb, alpha, beta, C, ldc,
            syrk ? 1 : 0);
        MGB_LAUNCHED("k_gemm_tn_ref");
        return MGB_OK;
    }
    const int tm = (m + BM - 1) / BM, tn = (n + BN - 1) / BN;
    const long long tiles = syrk ? (long long)tm * (tm + 1) / 2 : (long long)tm * tn;
    // split K so that about two waves of CTAs exist
    int nsplit = 1;
    if (tiles < 2 * 148)
    {
        nsplit = (int)((2 * 148 + tiles - 1) / tiles);
        const long long maxsplit = (long long)(k / (KC * 64)) > 0 ? (long long)(k / (KC * 64)) : 1;
        if (nsplit > maxsplit) nsplit = (int)maxsplit;
        if (nsplit > 512) nsplit = 512;
    }
    long long kper = ((long long)k + nsplit - 1) / nsplit;
    kper           = (kper + KC - 1) / KC * KC;
    nsplit         = (int)(((long long)k + kper - 1) / kper);
    double* partial = nullptr;
    if (nsplit > 1)
    {
        partial = (double*)scratch(2, (size_t)nsplit * m * n * sizeof(double));
        if (!partial) return MGB_ECUDA;
    }
    const size_t smem = (size_t)2 * STAGES * 128 * PITCH_K * sizeof(T);
    dim3 grid((unsigned)tn, (unsigned)tm, (unsigned)nsplit);
    if (syrk)
    {
        auto kern = k_gemm_tn<T, true>;
        MGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, NTHREADS, smem, st>>>(m, n, (long long)k, kper, A, (long long)lda, B,
            (long long)ldb, alpha, beta, C, ldc, partial, nsplit);
    }
    else
    {
        auto kern = k_gemm_tn<T, false>;
        MGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, NTHREADS, smem, st>>>(m, n, (long long)k, kper, A, (long long)lda, B,
            (long long)ldb, alpha, beta, C, ldc, partial, nsplit);
    }
    MGB_LAUNCHED("k_gemm_tn");
    if (nsplit > 1)
    {
        const long long tot = (long long)m * n;
        if (syrk)
            k_splitk_reduce<true><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(
                m, n, nsplit, partial, alpha, beta, C, ldc);
        else
            k_splitk_reduce<false><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(
                m, n, nsplit, partial, alpha, beta, C, ldc);
        MGB_LAUNCHED("k_splitk_reduce");
    }
    return MGB_OK;
}

template <typename T>
static int gemm_nn_t(size_t m, int n, int k, double alpha, const T* A, size_t lda,
    const double* M, int ldm, double beta, T* Out, size_t ldc, cudaStream_t st)
{
    const bool aligned = (((uintptr_t)A | (uintptr_t)M) & 15) == 0
                         && (lda * sizeof(T)) % 16 == 0 && (ldm * sizeof(double)) % 16 == 0;
    if (!aligned)
    {
        dim3 grid((unsigned)((m + 127) / 128), (unsigned)n);
        k_gemm_nn_ref<T><<<grid, 128, 0, st>>>((long long)m, n, k, A, (long long)lda, M,
            ldm, alpha, beta, Out, (long long)ldc);
        MGB_LAUNCHED("k_gemm_nn_ref");
        return MGB_OK;
    }
    const size_t smem = (size_t)STAGES * 128 * PITCH_K * sizeof(double)
                        + (size_t)STAGES * KC * PITCH_P * sizeof(T);
    const size_t ptiles = (m + 127) / 128;
    MGB_REQUIRE(ptiles <= 65535u * 32768u, "mgb_gemm_nn: too many points");
    // grid.y is limited to 65535: fold the point tiles over several launches
    auto kern = k_gemm_nn<T>;
    MGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (size_t t0 = 0; t0 < ptiles; t0 += 65535)
    {
        const size_t nt = (ptiles - t0 < 65535) ? ptiles - t0 : 65535;
        dim3 grid((unsigned)((n + 127) / 128), (unsigned)nt);
        const size_t poff = t0 * 128;
        kern<<<grid, NTHREADS, smem, st>>>((long long)(m - poff), n, k, A + poff,
            (long long)lda, M, ldm, alpha, beta, Out + poff, (long long)ldc);
        MGB_LAUNCHED("k_gemm_nn");
    }
    return MGB_OK;
}

} // namespace mgb

using namespace mgb;

extern "C"
{

int mgb_gemm_tn(int dtype, int m, int n, size_t k, double alpha, const void* A,
    size_t lda, const void* B, size_t ldb, double beta, double* C, int ldc, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(A && B && C, "mgb_gemm_tn: null pointer");
    MGB_REQUIRE(m >= 0 && n >= 0 && ldc >= m && lda >= k && ldb >= k,
        "mgb_gemm_tn: bad dimensions");
    if (m == 0 || n == 0) return MGB_OK;
    if (dtype == MGB_F64)
        return gemm_tn_t<double>(false, m, n, k, alpha, (const double*)A, lda,
            (const double*)B, ldb, beta, C, ldc, as_stream(stream));
    if (dtype == MGB_F32)
        return gemm_tn_t<float>(false, m, n, k, alpha, (const float*)A, lda,
            (const float*)B, ldb, beta, C, ldc, as_stream(stream));
    set_error("mgb_gemm_tn: bad dtype");
    return MGB_EINVAL;
}

int mgb_syrk_t(int dtype, int n, size_t k, double alpha, const void* A, size_t lda,
    double* C, int ldc, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(A && C, "mgb_syrk_t: null pointer");
    MGB_REQUIRE(n >= 0 && ldc >= n && lda >= k, "mgb_syrk_t: bad dimensions");
    if (n == 0) return MGB_OK;
    if (dtype == MGB_F64)
        return gemm_tn_t<double>(true, n, n, k, alpha, (const double*)A, lda,
            (const double*)A, lda, 0., C, ldc, as_stream(stream));
    if (dtype == MGB_F32)
        return gemm_tn_t<float>(true, n, n, k, alpha, (const float*)A, lda,
            (const float*)A, lda, 0., C, ldc, as_stream(stream));
    set_error("mgb_syrk_t: bad dtype");
    return MGB_EINVAL;
}

int mgb_gemm_nn(int dtype, size_t m, int n, int k, double alpha, const void* A,
    size_t lda, const double* M, int ldm, double beta, void* Cout, size_t ldc, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(A && M && Cout, "mgb_gemm_nn: null pointer");
    MGB_REQUIRE(A != Cout, "mgb_gemm_nn: output must not alias the input block");
    MGB_REQUIRE(n >= 0 && k >= 0 && lda >= m && ldc >= m && ldm >= k,
        "mgb_gemm_nn: bad dimensions");
    if (m == 0 || n == 0) return MGB_OK;
    if (dtype == MGB_F64)
        return gemm_nn_t<double>(m, n, k, alpha, (const double*)A, lda, M, ldm, beta,
            (double*)Cout, ldc, as_stream(stream));
    if (dtype == MGB_F32)
        return gemm_nn_t<float>(m, n, k, alpha, (const float*)A, lda, M, ldm, beta,
            (float*)Cout, ldc, as_stream(stream));
    set_error("mgb_gemm_nn: bad dtype");
    return MGB_EINVAL;
}

} // extern "C"
