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
// warps of 32x32, one CTA per SM, K staged through a 3-deep cp.async ring of
// 32-wide slabs whose rows are padded by 4 elements (fragment loads conflict-
// free).  The tall-skinny reductions (K = grid points, M = N = orbitals) are
// scheduled stream-K: every CTA takes an equal share of the (tile, K) space and
// a fix-up kernel adds the partial tiles in a fixed order, so results are
// run-to-run deterministic; diagonal Gram tiles do half the tensor work.
// ORBDTYPE float operands default to error-compensated 3xTF32 tensor tiles
// (k_gemm_tn_tf32 / k_gemm_nn_tf32); mgb_set_f32_contraction(1) selects the
// DMMA kernels on widened operands instead.
#include <cstdint>
#include <cstdlib>

#include "hpsi.h"
#include "tma_ptx.cuh"

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
template <typename T, int KCv = KC>
__device__ __forceinline__ void load_kmajor(T* sm, const T* base, long long ld,
    int row0, int nrows, long long k0, long long kend, int tid)
{
    constexpr int PITCH_K = KCv + PADK;
    constexpr int CH  = 16 / (int)sizeof(T);      // elements per 16B chunk
    constexpr int CPR = KCv / CH;                  // chunks per row
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

// ---------------------------------------------------------------------------
// C_b(m x n) = alpha * A_b^T B_b + beta * C_b for b < nbatch  (A rows = m index,
// B rows = n index, K contiguous).
//
// Work decomposition ("stream-K"): the (batch, tile) list times the K range is
// one linear space of k-iterations (KC wide), weighted by what an iteration
// costs -- a diagonal tile of the Gram matrix only needs about half the tensor
// work of a full tile (below).  Each of the G CTAs (one per SM) takes an equal
// share of that space, so there is no wave quantisation and no tail; a share
// covers one or more segments (tile, k-iteration range).  A segment that is a
// whole off-diagonal tile writes C directly; every other segment writes its
// 128x128 partial into its own slot and k_tn_fixup adds the slots of a tile in
// CTA order, i.e. the summation order is fixed: results are deterministic.
//
// Diagonal tiles of S = A^T A (SYRK).  View the 128x128 tile as 16x16 subtiles
// of 8x8 (one DMMA each).  Subtile (I, J) and its mirror (J, I) hold transposed
// values, so each is accumulated over only HALF of the k4 steps -- (I > J) on
// the even steps, (I < J) on the odd ones -- and S(I, J) = P(I, J) + P(J, I)^T
// is formed by the fix-up; diagonal subtiles take every step.  Every warp of
// the CTA does 32 or 40 of its 64 DMMAs per slab and the four SM sub-partitions
// stay balanced (each holds one diagonal warp and three off-diagonal ones).
// ---------------------------------------------------------------------------
struct TnWork
{
    long long nkt; // k-iterations per tile
    long long tot; // total cost units
    int ND, NT;    // diagonal-class tiles (listed first), all tiles
    int G, smax;   // CTAs, partial slots per CTA
    int tm, tn, nbatch;
    int cd, cf;    // cost units of one k-iteration: diagonal / full tile
};

__host__ __device__ inline long long tn_tile_start(const TnWork& W, int u)
{
    return u < W.ND ? (long long)u * W.cd * W.nkt
                    : (long long)W.ND * W.cd * W.nkt + (long long)(u - W.ND) * W.cf * W.nkt;
}
__host__ __device__ inline int tn_tile_of(const TnWork& W, long long b)
{
    const long long dspan = (long long)W.ND * W.cd * W.nkt;
    int u = b < dspan ? (int)(b / (W.cd * W.nkt)) : W.ND + (int)((b - dspan) / (W.cf * W.nkt));
    return u < W.NT ? u : W.NT - 1;
}
__host__ __device__ inline void tn_cta_bounds(const TnWork& W, int g, long long& b0, long long& b1)
{
    b0 = W.tot * g / W.G;
    b1 = W.tot * (g + 1) / W.G;
}
// k-iterations [it0, it1) of tile u inside the cost interval [b0, b1)
__host__ __device__ inline void tn_seg(const TnWork& W, int u, long long b0, long long b1,
    long long& it0, long long& it1)
{
    const long long s = tn_tile_start(W, u);
    const int c       = u < W.ND ? W.cd : W.cf;
    const long long r0 = b0 - s, r1 = b1 - s;
    it0 = r0 <= 0 ? 0 : (r0 + c - 1) / c;
    it1 = r1 <= 0 ? 0 : (r1 + c - 1) / c;
    if (it0 > W.nkt) it0 = W.nkt;
    if (it1 > W.nkt) it1 = W.nkt;
}
// tile u -> (batch, tile_m, tile_n)
template <bool SYRK>
__host__ __device__ inline void tn_decode(const TnWork& W, int u, int& batch, int& tile_m,
    int& tile_n)
{
    if (SYRK)
    {
        if (u < W.ND)
        {
            batch  = u / W.tm;
            tile_m = tile_n = u % W.tm;
            return;
        }
        const int noff = W.tm * (W.tm - 1) / 2;
        const int o    = u - W.ND;
        batch          = o / noff;
        const int idx  = o % noff;
        int r          = 1;
        while ((r + 1) * r / 2 <= idx) r++; // row r holds indices r(r-1)/2 .. r(r+1)/2-1
        tile_m = r;
        tile_n = idx - r * (r - 1) / 2;
    }
    else
    {
        const int per = W.tm * W.tn;
        batch         = u / per;
        const int r   = u % per;
        tile_m        = r % W.tm;
        tile_n        = r / W.tm;
    }
}

// one k4 step of a warp: 4 + 4 fragment loads, then the DMMAs MODE selects
//   0 all 16 subtiles, 1 diagonal warp on an even step (i >= j),
//   2 diagonal warp on an odd step (i <= j)
template <typename T, int MODE, int PITCH>
__device__ __forceinline__ void tn_k4(double (&acc)[4][4][2], const T* as, const T* bs)
{
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        a[i] = (double)as[i * 8 * PITCH];
        b[i] = (double)bs[i * 8 * PITCH];
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (MODE == 0 || (MODE == 1 && i >= j) || (MODE == 2 && i <= j))
                dmma(acc[i][j], a[i], b[j]);
}

template <typename T, bool DIAG, int KCv, int ST>
__device__ __forceinline__ void tn_segment(double (&acc)[4][4][2], T* As, T* Bs, const T* A,
    long long lda, const T* B, long long ldb, int m0, int m, int n0, int n, long long it0,
    long long it1, long long kend, int tid, int wm, int wn, int fr, int fc)
{
    constexpr int PITCH = KCv + PADK;
    const int nit = (int)(it1 - it0);
    // the ring is reused by the next segment: everybody must be done reading
    __syncthreads();
#pragma unroll
    for (int s = 0; s < ST - 1; s++)
    {
        if (s < nit)
        {
            load_kmajor<T, KCv>(As + s * 128 * PITCH, A, lda, m0, m, (it0 + s) * KCv, kend, tid);
            load_kmajor<T, KCv>(Bs + s * 128 * PITCH, B, ldb, n0, n, (it0 + s) * KCv, kend, tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < nit; kt++)
    {
        cp_async_wait<ST - 2>();
        __syncthreads();
        {
            const int nx = kt + ST - 1;
            if (nx < nit)
            {
                const int s = nx % ST;
                load_kmajor<T, KCv>(
                    As + s * 128 * PITCH, A, lda, m0, m, (it0 + nx) * KCv, kend, tid);
                load_kmajor<T, KCv>(
                    Bs + s * 128 * PITCH, B, ldb, n0, n, (it0 + nx) * KCv, kend, tid);
            }
            cp_async_commit();
        }
        const T* as = As + (kt % ST) * 128 * PITCH + (wm * 32 + fr) * PITCH + fc;
        const T* bs = Bs + (kt % ST) * 128 * PITCH + (wn * 32 + fr) * PITCH + fc;
        if (!DIAG)
        {
#pragma unroll
            for (int kk = 0; kk < KCv / 4; kk++)
                tn_k4<T, 0, PITCH>(acc, as + kk * 4, bs + kk * 4);
        }
        else if (wm > wn)
        {
            // lower warp block: every subtile has I > J -> even steps only
#pragma unroll
            for (int kk = 0; kk < KCv / 4; kk += 2)
                tn_k4<T, 0, PITCH>(acc, as + kk * 4, bs + kk * 4);
        }
        else if (wm < wn)
        {
#pragma unroll
            for (int kk = 1; kk < KCv / 4; kk += 2)
                tn_k4<T, 0, PITCH>(acc, as + kk * 4, bs + kk * 4);
        }
        else
        {
            // diagonal warp block: I - J = i - j
#pragma unroll
            for (int kk = 0; kk < KCv / 4; kk += 2)
            {
                tn_k4<T, 1, PITCH>(acc, as + kk * 4, bs + kk * 4);
                tn_k4<T, 2, PITCH>(acc, as + (kk + 1) * 4, bs + (kk + 1) * 4);
            }
        }
    }
    cp_async_wait<0>();
}

template <typename T, bool SYRK, int KCv, int ST>
__global__ void __launch_bounds__(NTHREADS, 1) k_gemm_tn(TnWork W, int m, int n, long long k,
    const T* __restrict__ A, long long lda, long long strideA, const T* __restrict__ B,
    long long ldb, long long strideB, double alpha, double beta, double* __restrict__ C,
    int ldc, long long strideC, double* __restrict__ partial)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    T* As = reinterpret_cast<T*>(smraw);
    T* Bs = As + ST * 128 * (KCv + PADK);

    const int g = blockIdx.x;
    long long b0, b1;
    tn_cta_bounds(W, g, b0, b1);
    if (b1 <= b0) return;
    const int u_first = tn_tile_of(W, b0), u_last = tn_tile_of(W, b1 - 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3; // 4 x 4 warps, 32 x 32 each
    const int fr = lane >> 2, fc = lane & 3;

    for (int u = u_first; u <= u_last; u++)
    {
        long long it0, it1;
        tn_seg(W, u, b0, b1, it0, it1);
        if (it0 >= it1) continue;
        int batch, tile_m, tile_n;
        tn_decode<SYRK>(W, u, batch, tile_m, tile_n);
        const int m0 = tile_m * BM, n0 = tile_n * BN;
        const bool diag = SYRK && u < W.ND;
        const T* Ab = A + (long long)batch * strideA;
        const T* Bb = B + (long long)batch * strideB;

        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
                acc[i][j][0] = acc[i][j][1] = 0.;

        if (diag)
            tn_segment<T, true, KCv, ST>(acc, As, Bs, Ab, lda, Bb, ldb, m0, m, n0, n, it0, it1,
                k, tid, wm, wn, fr, fc);
        else
            tn_segment<T, false, KCv, ST>(acc, As, Bs, Ab, lda, Bb, ldb, m0, m, n0, n, it0, it1,
                k, tid, wm, wn, fr, fc);

        // thread holds C(m0 + wm*32 + i*8 + fr, n0 + wn*32 + j*8 + fc*2 + e)
        const bool direct = !diag && it0 == 0 && it1 == W.nkt;
        if (direct)
        {
            double* Cb = C + (long long)batch * strideC;
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
                            const double old
                                = (beta == 0.) ? 0. : beta * Cb[(size_t)nn * ldc + mm];
                            const double val          = alpha * acc[i][j][e] + old;
                            Cb[(size_t)nn * ldc + mm] = val;
                            // LocalMatrices::fillUpperWithLower
                            if (SYRK) Cb[(size_t)mm * ldc + nn] = val;
                        }
                    }
        }
        else
        {
            double* dst = partial + ((size_t)g * W.smax + (size_t)(u - u_first)) * (BM * BN);
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
                {
                    const int ml = wm * 32 + i * 8 + fr;
                    const int nl = wn * 32 + j * 8 + fc * 2;
                    dst[(size_t)nl * BM + ml]       = acc[i][j][0];
                    dst[(size_t)(nl + 1) * BM + ml] = acc[i][j][1];
                }
        }
    }
}

// ---------------------------------------------------------------------------
// ORBDTYPE float on the FP32-class tensor path: error-compensated 3xTF32
// (mma.sync.m16n8k8.tf32).  Each operand is split a = a_hi + a_lo into two
// TF32 numbers and a*b is formed as a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, which
// keeps ~21 bits of every product; the tensor core accumulates in FP32 over one
// 32-point slab only, slab sums are added in FP32 over at most 64 slabs and
// those sums in double, so the long K = npt reduction behaves like the
// reference's double accumulation (MPgemm float path, src/linear_algebra/
// mputils.cc:763-774) to ~1e-6 of |a||b| -- inside the 1e-5 FP32 bar of the
// north star; mgb_set_f32_contraction(1) selects the DMMA kernel above, whose
// products and sums are exactly those double ones.
// Same stream-K decomposition, same fix-up; the half-work trick of diagonal
// Gram tiles pairs 16x16 subtiles here.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4],
    const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 "
                 "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int KC32  = 32;            // K slab of the float kernel
constexpr int P32   = KC32 + PADK;   // smem row pitch (floats): conflict-free fragments
constexpr int ST32  = 4;             // cp.async ring depth

// one k8 step of a warp (32 x 32 = 2 m16 x 4 n8): MODE 0 all blocks; on diagonal
// warp blocks MODE 1 = even step (16x16 blocks with i >= j), MODE 2 = odd (i <= j)
template <int MODE>
__device__ __forceinline__ void tn_k8_tf32(float (&c)[2][4][4], const float* as, const float* bs)
{
    uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            // a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)
            const float v = as[(i * 16 + (r & 1) * 8) * P32 + (r >> 1) * 4];
            ah[i][r]      = to_tf32(v);
            al[i][r]      = to_tf32(v - __uint_as_float(ah[i][r]));
        }
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int r = 0; r < 2; r++)
        {
            // b0 (k = t, n = g)  b1 (k = t+4, n = g)
            const float v = bs[(j * 8) * P32 + r * 4];
            bh[j][r]      = to_tf32(v);
            bl[j][r]      = to_tf32(v - __uint_as_float(bh[j][r]));
        }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            const int J = j >> 1; // 16-wide block column inside the warp tile
            if (MODE == 0 || (MODE == 1 && i >= J) || (MODE == 2 && i <= J))
            {
                mma_tf32(c[i][j], al[i], bh[j]);
                mma_tf32(c[i][j], ah[i], bl[j]);
                mma_tf32(c[i][j], ah[i], bh[j]);
            }
        }
}

constexpr int FOLD2 = 64; // level-1 sums folded into the double sums every FOLD2 slabs

template <bool DIAG>
__device__ __forceinline__ void tn_segment_tf32(double* __restrict__ dst, float* As, float* Bs,
    const float* A, long long lda, const float* B, long long ldb, int m0, int m, int n0, int n,
    long long it0, long long it1, long long kend, int tid, int wm, int wn, int g, int t)
{
    const int nit = (int)(it1 - it0);
    // three levels: c = tensor-core FP32 sums of one slab (32 points), hi = FP32
    // sums of up to FOLD2 slabs, dst = double sums (this segment's partial slot,
    // L2-resident, touched once every FOLD2 slabs)
    float c[2][4][4], hi[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 4; e++)
                c[i][j][e] = hi[i][j][e] = 0.f;
    bool first = true;
    int pending = 0;
    __syncthreads();
#pragma unroll
    for (int s = 0; s < ST32 - 1; s++)
    {
        if (s < nit)
        {
            load_kmajor<float, KC32>(As + s * 128 * P32, A, lda, m0, m, (it0 + s) * KC32, kend, tid);
            load_kmajor<float, KC32>(Bs + s * 128 * P32, B, ldb, n0, n, (it0 + s) * KC32, kend, tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < nit; kt++)
    {
        cp_async_wait<ST32 - 2>();
        __syncthreads();
        {
            const int nx = kt + ST32 - 1;
            if (nx < nit)
            {
                const int s = nx % ST32;
                load_kmajor<float, KC32>(
                    As + s * 128 * P32, A, lda, m0, m, (it0 + nx) * KC32, kend, tid);
                load_kmajor<float, KC32>(
                    Bs + s * 128 * P32, B, ldb, n0, n, (it0 + nx) * KC32, kend, tid);
            }
            cp_async_commit();
        }
        const float* as = As + (kt % ST32) * 128 * P32 + (wm * 32 + g) * P32 + t;
        const float* bs = Bs + (kt % ST32) * 128 * P32 + (wn * 32 + g) * P32 + t;
        if (!DIAG)
        {
#pragma unroll
            for (int kk = 0; kk < KC32 / 8; kk++)
                tn_k8_tf32<0>(c, as + kk * 8, bs + kk * 8);
        }
        else if (wm > wn)
        {
#pragma unroll
            for (int kk = 0; kk < KC32 / 8; kk += 2)
                tn_k8_tf32<0>(c, as + kk * 8, bs + kk * 8);
        }
        else if (wm < wn)
        {
#pragma unroll
            for (int kk = 1; kk < KC32 / 8; kk += 2)
                tn_k8_tf32<0>(c, as + kk * 8, bs + kk * 8);
        }
        else
        {
#pragma unroll
            for (int kk = 0; kk < KC32 / 8; kk += 2)
            {
                tn_k8_tf32<1>(c, as + kk * 8, bs + kk * 8);
                tn_k8_tf32<2>(c, as + (kk + 1) * 8, bs + (kk + 1) * 8);
            }
        }
        // level 1: slab sums into the FP32 running sums
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int e = 0; e < 4; e++)
                {
                    hi[i][j][e] = __fadd_rn(hi[i][j][e], c[i][j][e]);
                    c[i][j][e]  = 0.f;
                }
        if (++pending == FOLD2 || kt == nit - 1)
        {
            // level 2: into the double sums
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int e = 0; e < 4; e++)
                    {
                        const int ml = wm * 32 + i * 16 + g + 8 * (e >> 1);
                        const int nl = wn * 32 + j * 8 + 2 * t + (e & 1);
                        double* d    = dst + (size_t)nl * BM + ml;
                        *d           = (first ? 0. : *d) + (double)hi[i][j][e];
                        hi[i][j][e]  = 0.f;
                    }
            first   = false;
            pending = 0;
        }
    }
    cp_async_wait<0>();
}

template <bool SYRK>
__global__ void __launch_bounds__(NTHREADS, 1) k_gemm_tn_tf32(TnWork W, int m, int n,
    long long k, const float* __restrict__ A, long long lda, long long strideA,
    const float* __restrict__ B, long long ldb, long long strideB, double alpha, double beta,
    double* __restrict__ C, int ldc, long long strideC, double* __restrict__ partial)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    float* As = reinterpret_cast<float*>(smraw);
    float* Bs = As + ST32 * 128 * P32;

    const int gcta = blockIdx.x;
    long long b0, b1;
    tn_cta_bounds(W, gcta, b0, b1);
    if (b1 <= b0) return;
    const int u_first = tn_tile_of(W, b0), u_last = tn_tile_of(W, b1 - 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, t = lane & 3;

    for (int u = u_first; u <= u_last; u++)
    {
        long long it0, it1;
        tn_seg(W, u, b0, b1, it0, it1);
        if (it0 >= it1) continue;
        int batch, tile_m, tile_n;
        tn_decode<SYRK>(W, u, batch, tile_m, tile_n);
        const int m0 = tile_m * BM, n0 = tile_n * BN;
        const bool diag = SYRK && u < W.ND;
        const float* Ab = A + (long long)batch * strideA;
        const float* Bb = B + (long long)batch * strideB;
        double* dst = partial + ((size_t)gcta * W.smax + (size_t)(u - u_first)) * (BM * BN);
        if (diag)
            tn_segment_tf32<true>(dst, As, Bs, Ab, lda, Bb, ldb, m0, m, n0, n, it0, it1, k, tid,
                wm, wn, g, t);
        else
            tn_segment_tf32<false>(dst, As, Bs, Ab, lda, Bb, ldb, m0, m, n0, n, it0, it1, k, tid,
                wm, wn, g, t);

        // a whole off-diagonal tile: finish it here (each thread re-reads the
        // sums it wrote itself); everything else is left to the fix-up
        if (!diag && it0 == 0 && it1 == W.nkt)
        {
            double* Cb = C + (long long)batch * strideC;
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int e = 0; e < 4; e++)
                    {
                        const int ml = wm * 32 + i * 16 + g + 8 * (e >> 1);
                        const int nl = wn * 32 + j * 8 + 2 * t + (e & 1);
                        const int mm = m0 + ml, nn = n0 + nl;
                        if (mm < m && nn < n)
                        {
                            const double old
                                = (beta == 0.) ? 0. : beta * Cb[(size_t)nn * ldc + mm];
                            const double val = alpha * dst[(size_t)nl * BM + ml] + old;
                            Cb[(size_t)nn * ldc + mm] = val;
                            if (SYRK) Cb[(size_t)mm * ldc + nn] = val;
                        }
                    }
        }
    }
}

// Adds the partial slots of every tile that was not written directly, in CTA
// order; for diagonal Gram tiles it also folds the mirrored half-sums,
// S(r, c) = P(r, c) + P(c, r) unless both lie in the same 8x8 subtile.
// grid: (tiles, 16 column blocks of 8), block 128 x 2.
template <bool SYRK>
__global__ void k_tn_fixup(TnWork W, int m, int n, const double* __restrict__ partial,
    double alpha, double beta, double* __restrict__ C, int ldc, long long strideC, int sub_shift)
{
    // which slots hold partials of this tile: resolved once per block (the
    // interval arithmetic is 64-bit division heavy), in CTA order
    __shared__ long long slot_of[256]; // element offset of the slot, or -1
    __shared__ int nslots, is_direct;
    const int u = blockIdx.x;
    const long long s0 = tn_tile_start(W, u);
    const long long s1 = s0 + (long long)(u < W.ND ? W.cd : W.cf) * W.nkt;
    int gf = (int)(s0 * W.G / W.tot), gl = (int)((s1 - 1) * W.G / W.tot);
    if (gf > 0) gf--;
    if (gl < W.G - 1) gl++;
    const bool diag = SYRK && u < W.ND;
    const int tl    = threadIdx.y * blockDim.x + threadIdx.x;
    if (tl == 0)
    {
        nslots    = gl - gf + 1;
        is_direct = 0;
    }
    __syncthreads();
    if (tl <= gl - gf) // G <= 256 CTAs (one per SM)
    {
        const int g = gf + tl;
        long long b0, b1, it0 = 0, it1 = 0;
        tn_cta_bounds(W, g, b0, b1);
        long long off = -1;
        if (b1 > b0)
        {
            tn_seg(W, u, b0, b1, it0, it1);
            if (it0 < it1)
            {
                if (!diag && it0 == 0 && it1 == W.nkt)
                    is_direct = 1; // written by the main kernel
                else
                    off = ((long long)g * W.smax + (u - tn_tile_of(W, b0))) * (BM * BN);
            }
        }
        slot_of[tl] = off;
    }
    __syncthreads();
    if (is_direct) return;
    int batch, tile_m, tile_n;
    tn_decode<SYRK>(W, u, batch, tile_m, tile_n);
    const int ml = threadIdx.x;
    const int mm = tile_m * BM + ml;
    double* Cb   = C + (long long)batch * strideC;
    const int ns = nslots;
    for (int nl = blockIdx.y * 8 + threadIdx.y; nl < blockIdx.y * 8 + 8; nl += blockDim.y)
    {
        const int nn = tile_n * BN + nl;
        if (diag && nl > ml) continue; // lower triangle of the diagonal tile
        // mirrored half-sums are folded across subtiles (8x8 DMMA / 16x16 TF32)
        // sub_shift < 0 (tcgen05 kernel): the tile holds P with S = (P + P^T) / 2
        const bool fold = diag && (sub_shift < 0 || (ml >> sub_shift) != (nl >> sub_shift));
        const double fs = (diag && sub_shift < 0) ? 0.5 : 1.;
        double s = 0.;
        bool any = false;
        for (int i = 0; i < ns; i++)
        {
            const long long off = slot_of[i];
            if (off < 0) continue;
            const double* src = partial + off;
            s += src[(size_t)nl * BM + ml];
            if (fold) s += src[(size_t)ml * BM + nl];
            any = true;
        }
        if (!any || mm >= m || nn >= n) continue;
        const double old = (beta == 0.) ? 0. : beta * Cb[(size_t)nn * ldc + mm];
        const double val = alpha * (fs * s) + old;
        Cb[(size_t)nn * ldc + mm] = val;
        if (SYRK) Cb[(size_t)mm * ldc + nn] = val;
    }
}

// Out(n x npt) = alpha * M^T Phi + beta*Out : rows of Out/Phi are orbitals
// (points contiguous).  A' = M^T tile [128 j][KC l] (K-major: M is column-major
// l + j*ldm), B' = Phi tile [KC l][128 p] (N-major).
// row pitch of the point-major Phi slab: the four k rows a fragment load touches
// must fall into different bank groups (8-byte elements: pitch = 4 mod 16;
// 4-byte: pitch = 8 mod 32)
template <typename T>
struct PitchP
{
    static constexpr int value = sizeof(T) == 8 ? 128 + 4 : 128 + 8;
};

// Persistent: one CTA per SM walks a contiguous range of (point tile, orbital
// tile) items; the cp.async ring runs across item boundaries, so the operand
// stream never drains while a tile's results are written out.
template <typename T, int KCv, int ST, bool HASD>
__global__ void __launch_bounds__(NTHREADS, 1) k_gemm_nn(long long npt, int n, int k,
    const T* __restrict__ Phi, long long lda, const double* __restrict__ M, int ldm,
    double alpha, double beta, T* __restrict__ Out, long long ldc, long long nitems,
    int jtiles, double gamma, const T* __restrict__ D, long long ldd)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    constexpr int PITCH_K = KCv + PADK;
    constexpr int PITCH_P = PitchP<T>::value;
    constexpr int KC      = KCv;
    constexpr int STAGES  = ST;
    double* Ms = reinterpret_cast<double*>(smraw);               // [STAGES][128][PITCH_K]
    T* Ps      = reinterpret_cast<T*>(Ms + STAGES * 128 * PITCH_K); // [STAGES][KC][PITCH_P]

    const long long i0 = nitems * blockIdx.x / gridDim.x;
    const long long i1 = nitems * (blockIdx.x + 1) / gridDim.x;
    if (i1 <= i0) return;
    const int nkt         = (k + KC - 1) / KC;
    const long long total = (i1 - i0) * nkt;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int fr = lane >> 2, fc = lane & 3;
    // pairs of results go out as one 2-element store when every row allows it
    const bool vec2 = (((uintptr_t)Out) % (2 * sizeof(T)) == 0) && (ldc % 2 == 0);

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
            acc[i][j][0] = acc[i][j][1] = 0.;

    // producer position (item, k-slab, ring stage) of the next load to issue
    long long p_item = i0, p_issued = 0;
    int p_kt = 0, p_stage = 0;
    auto issue = [&]() {
        if (p_issued < total)
        {
            const int j0       = (int)(p_item % jtiles) * 128;
            const long long p0 = (p_item / jtiles) * 128;
            load_kmajor<double, KCv>(
                Ms + p_stage * 128 * PITCH_K, M, ldm, j0, n, (long long)p_kt * KC, k, tid);
            constexpr int CH  = 16 / (int)sizeof(T);
            constexpr int CPR = 128 / CH; // chunks per row of 128 points
            constexpr int TOT = KC * CPR;
            T* sm = Ps + p_stage * KC * PITCH_P;
#pragma unroll
            for (int c = tid; c < TOT; c += NTHREADS)
            {
                const int r  = c / CPR;
                const int cc = c % CPR;
                const int l  = p_kt * KC + r;
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
            p_issued++;
            if (++p_kt == nkt)
            {
                p_kt = 0;
                p_item++;
            }
            if (++p_stage == STAGES) p_stage = 0;
        }
        cp_async_commit();
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++)
        issue();

    long long item = i0;
    int kt = 0, stage = 0;
    for (long long it = 0; it < total; it++)
    {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        issue();
        const double* as = Ms + stage * 128 * PITCH_K + (wm * 32 + fr) * PITCH_K + fc;
        const T* bs      = Ps + stage * KC * PITCH_P + fc * PITCH_P + wn * 32 + fr;
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
        if (++stage == STAGES) stage = 0;
        if (++kt == nkt)
        {
            // thread holds Out(j0 + wm*32 + i*8 + fr, p0 + wn*32 + j*8 + fc*2 + {0,1})
            const int j0       = (int)(item % jtiles) * 128;
            const long long p0 = (item / jtiles) * 128;
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                const int jj = j0 + wm * 32 + i * 8 + fr;
                // the subtrahend of the four pairs of this row: all loads issued
                // before any is consumed
                T dv[4][2];
                if (HASD)
                {
#pragma unroll
                    for (int j = 0; j < 4; j++)
#pragma unroll
                        for (int e = 0; e < 2; e++)
                        {
                            const long long pp = p0 + wn * 32 + j * 8 + fc * 2 + e;
                            dv[j][e] = (jj < n && pp < npt) ? __ldg(D + (long long)jj * ldd + pp)
                                                            : (T)0;
                        }
                }
#pragma unroll
                for (int j = 0; j < 4; j++)
                {
                    const long long pp = p0 + wn * 32 + j * 8 + fc * 2;
                    T* o = Out + (long long)jj * ldc + pp;
                    // MPscal(beta) then += (T)buff  (mputils.cc:750-759)
                    T r[2];
#pragma unroll
                    for (int e = 0; e < 2; e++)
                    {
                        T base = (T)0;
                        if (beta != 0. && jj < n && pp + e < npt)
                            base = (T)(beta * (double)o[e]);
                        r[e]         = base + (T)(alpha * acc[i][j][e]);
                        acc[i][j][e] = 0.;
                        // fused "Out.axpy(gamma, D)" (MPaxpy: y += (T)(gamma * (double)x),
                        // mputils.cc:222-244) on the freshly rounded product
                        if (HASD) r[e] += (T)(gamma * (double)dv[j][e]);
                    }
                    if (jj < n)
                    {
                        if (vec2 && pp + 1 < npt)
                        {
                            if (sizeof(T) == 8)
                                *reinterpret_cast<double2*>(o) = make_double2(r[0], r[1]);
                            else
                                *reinterpret_cast<float2*>(o) = make_float2(r[0], r[1]);
                        }
                        else
                        {
                            if (pp < npt) o[0] = r[0];
                            if (pp + 1 < npt) o[1] = r[1];
                        }
                    }
                }
            }
            kt = 0;
            item++;
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------
// Phi M for ORBDTYPE float on the FP32-class tensor path (3xTF32, see
// k_gemm_tn_tf32): Out^T tile (128 orbitals x 128 points) = Mf^T tile x Phi tile
// with Mf = (float)(alpha M) prepared once per call.  The reduction is only
// K = numst long; the tensor core's FP32 sums are folded into a second FP32
// accumulator every 64 terms.  Persistent CTAs like k_gemm_nn.
// ---------------------------------------------------------------------------
__global__ void k_scale_to_float(long long n, double alpha, const double* __restrict__ in,
    float* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)(alpha * in[i]);
}

constexpr int PP32 = 128 + 8; // point-major Phi slab pitch (floats)

template <bool HASD>
__global__ void __launch_bounds__(NTHREADS, 1) k_gemm_nn_tf32(long long npt, int n, int k,
    const float* __restrict__ Phi, long long lda, const float* __restrict__ Mf, int ldm,
    double beta, float* __restrict__ Out, long long ldc, long long nitems, int jtiles,
    double gamma, const float* __restrict__ D, long long ldd)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    float* Ms = reinterpret_cast<float*>(smraw);     // [ST32][128 j][P32]   (K-major)
    float* Ps = Ms + ST32 * 128 * P32;               // [ST32][KC32 l][PP32] (point-major)

    const long long i0 = nitems * blockIdx.x / gridDim.x;
    const long long i1 = nitems * (blockIdx.x + 1) / gridDim.x;
    if (i1 <= i0) return;
    const int nkt         = (k + KC32 - 1) / KC32;
    const long long total = (i1 - i0) * nkt;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, t = lane & 3;
    const bool vec2 = (((uintptr_t)Out) % 8 == 0) && (ldc % 2 == 0);

    float c[2][4][4], hi[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 4; e++)
                c[i][j][e] = hi[i][j][e] = 0.f;

    long long p_item = i0, p_issued = 0;
    int p_kt = 0, p_stage = 0;
    auto issue = [&]() {
        if (p_issued < total)
        {
            const int j0       = (int)(p_item % jtiles) * 128;
            const long long p0 = (p_item / jtiles) * 128;
            load_kmajor<float, KC32>(
                Ms + p_stage * 128 * P32, Mf, ldm, j0, n, (long long)p_kt * KC32, k, tid);
            constexpr int CPR = 128 / 4; // 16-byte chunks per row of 128 points
            constexpr int TOT = KC32 * CPR;
            float* sm = Ps + p_stage * KC32 * PP32;
#pragma unroll
            for (int cidx = tid; cidx < TOT; cidx += NTHREADS)
            {
                const int r  = cidx / CPR;
                const int cc = cidx % CPR;
                const int l  = p_kt * KC32 + r;
                const long long p = p0 + (long long)cc * 4;
                int valid = 0;
                if (l < k && p < npt)
                {
                    const long long rem = npt - p;
                    valid = rem >= 4 ? 16 : (int)rem * 4;
                }
                const float* src = Phi + (long long)(l < k ? l : 0) * lda + (valid ? p : 0);
                cp_async16(sm + r * PP32 + cc * 4, src, valid);
            }
            p_issued++;
            if (++p_kt == nkt)
            {
                p_kt = 0;
                p_item++;
            }
            if (++p_stage == ST32) p_stage = 0;
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < ST32 - 1; s++)
        issue();

    long long item = i0;
    int kt = 0, stage = 0;
    for (long long it = 0; it < total; it++)
    {
        cp_async_wait<ST32 - 2>();
        __syncthreads();
        issue();
        const float* as = Ms + stage * 128 * P32 + (wm * 32 + g) * P32 + t;
        const float* bs = Ps + stage * KC32 * PP32 + t * PP32 + wn * 32 + g;
#pragma unroll
        for (int kk = 0; kk < KC32 / 8; kk++)
        {
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int r = 0; r < 4; r++)
                {
                    const float v = as[(i * 16 + (r & 1) * 8) * P32 + kk * 8 + (r >> 1) * 4];
                    ah[i][r]      = to_tf32(v);
                    al[i][r]      = to_tf32(v - __uint_as_float(ah[i][r]));
                }
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
                uint32_t bh[2], bl[2];
#pragma unroll
                for (int r = 0; r < 2; r++)
                {
                    // b0 (k = t, n = g)  b1 (k = t + 4, n = g)
                    const float v = bs[(kk * 8 + r * 4) * PP32 + j * 8];
                    bh[r]         = to_tf32(v);
                    bl[r]         = to_tf32(v - __uint_as_float(bh[r]));
                }
#pragma unroll
                for (int i = 0; i < 2; i++)
                {
                    mma_tf32(c[i][j], al[i], bh);
                    mma_tf32(c[i][j], ah[i], bl);
                    mma_tf32(c[i][j], ah[i], bh);
                }
            }
        }
        // slab sums into the second-level FP32 sums every two slabs (64 terms)
        if ((kt & 1) == 1 || kt == nkt - 1)
        {
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int e = 0; e < 4; e++)
                    {
                        hi[i][j][e] = __fadd_rn(hi[i][j][e], c[i][j][e]);
                        c[i][j][e]  = 0.f;
                    }
        }
        if (++stage == ST32) stage = 0;
        if (++kt == nkt)
        {
            // thread holds Out(j0 + wm*32 + i*16 + g + 8*(e>>1), p0 + wn*32 + j*8 + 2t + (e&1))
            const int j0       = (int)(item % jtiles) * 128;
            const long long p0 = (item / jtiles) * 128;
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int h = 0; h < 2; h++)
                {
                    const int jj = j0 + wm * 32 + i * 16 + g + 8 * h;
#pragma unroll
                    for (int j = 0; j < 4; j++)
                    {
                        const long long pp = p0 + wn * 32 + j * 8 + 2 * t;
                        float* o = Out + (long long)jj * ldc + pp;
                        float r[2];
#pragma unroll
                        for (int e = 0; e < 2; e++)
                        {
                            float base = 0.f;
                            if (beta != 0. && jj < n && pp + e < npt)
                                base = (float)(beta * (double)o[e]);
                            r[e] = base + hi[i][j][2 * h + e];
                            hi[i][j][2 * h + e] = 0.f;
                            if (HASD && jj < n && pp + e < npt)
                                r[e] += (float)(gamma * (double)D[(long long)jj * ldd + pp + e]);
                        }
                        if (jj < n)
                        {
                            if (vec2 && pp + 1 < npt)
                                *reinterpret_cast<float2*>(o) = make_float2(r[0], r[1]);
                            else
                            {
                                if (pp < npt) o[0] = r[0];
                                if (pp + 1 < npt) o[1] = r[1];
                            }
                        }
                    }
                }
            kt = 0;
            item++;
        }
    }
    cp_async_wait<0>();
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

// tcgen05 (kind::tf32, TMEM) kernels of the float contractions, inside namespace mgb
#include "tn_umma.cuh"
#include "nn_umma.cuh"

// float operands: 0 = 3xTF32 tensor tiles (tcgen05 where the kernel takes the shape, else
// mma.sync), 1 = DMMA (double products), 2 = 3xTF32 on mma.sync only
static int g_f32_mode = 0;
#define g_f32_exact (g_f32_mode == 1)

static int num_sms()
{
    static int n = 0;
    if (n == 0)
    {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess
            || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess
            || n <= 0)
            n = 148;
    }
    return n;
}

template <typename T>
static int gemm_tn_t(bool syrk, int m, int n, size_t k, double alpha, const T* A,
    size_t lda, size_t strideA, const T* B, size_t ldb, size_t strideB, double beta, double* C,
    int ldc, size_t strideC, int nbatch, cudaStream_t st)
{
    const bool aligned = (((uintptr_t)A | (uintptr_t)B) & 15) == 0
                         && (lda * sizeof(T)) % 16 == 0 && (ldb * sizeof(T)) % 16 == 0
                         && (nbatch == 1
                             || ((strideA * sizeof(T)) % 16 == 0 && (strideB * sizeof(T)) % 16 == 0));
    if (!aligned || k < 64)
    {
        const long long tot = (long long)m * n;
        for (int b = 0; b < nbatch; b++)
        {
            k_gemm_tn_ref<T><<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(m, n, (long long)k,
                A + b * strideA, (long long)lda, B + b * strideB, (long long)ldb, alpha, beta,
                C + b * strideC, ldc, syrk ? 1 : 0);
            MGB_LAUNCHED("k_gemm_tn_ref");
        }
        return MGB_OK;
    }
    TnWork W;
    W.tm     = (m + BM - 1) / BM;
    W.tn     = (n + BN - 1) / BN;
    W.nbatch = nbatch;
    const bool tf32 = sizeof(T) == 4 && !g_f32_exact;
    // tcgen05 kernel: TMA needs 16-byte pitches (checked above); MGB_TN_UMMA=0 is a tuning hook
    bool use_umma = tf32 && g_f32_mode == 0;
    if (const char* env = getenv("MGB_TN_UMMA")) use_umma = use_umma && atoi(env) != 0;
    const int sub_shift = use_umma ? -1 : tf32 ? 4 : 3;
    // slab width / ring depth: 16 x 4 or 32 x 3 (MGB_TN_KC, tuning hook)
    int kcv = 32;
    if (const char* env = getenv("MGB_TN_KC")) kcv = atoi(env) == 16 ? 16 : 32;
    if (tf32) kcv = KC32;
    // tcgen05 kernel: one k-iteration = one TMA box of KO 32-point slabs (5-D maps: whole 8-row
    // groups and whole slabs only; any other shape takes one slab per box)
    int ko = 1;
    if (use_umma)
    {
        ko = (m % 8 == 0 && n % 8 == 0 && k % 32 == 0) ? 2 : 1;
        if (const char* env = getenv("MGB_UMMA_KO"))
            if (ko > 1 && atoi(env) >= 1 && atoi(env) <= 2) ko = atoi(env);
        kcv = 32 * ko;
    }
    const int st_ = kcv == 32 ? 3 : 4;
    W.nkt    = (long long)((k + kcv - 1) / kcv);
    // cost of one k-iteration: a diagonal Gram tile issues 136 of the 256 DMMAs
    // of a full tile but stages the same operands (10/16 measured best on B200)
    W.cf = 16;
    W.cd = 10;
    if (sizeof(T) == 4 && !g_f32_exact) W.cd = 12; // 3xTF32: 16x16 pairing, measured best
    // tcgen05: 2 of 3 MMAs, one operand box, no low tile of B -- but a slab of this kernel is
    // bound by its hand-offs more than by its MMAs: 14/16 measured best (N = 256, 512, 1024)
    if (use_umma) W.cd = 14;
    if (const char* env = getenv("MGB_SYRK_DIAG_COST"))
    {
        const int c = atoi(env);
        if (c >= 1 && c <= 16) W.cd = c;
    }
    if (syrk)
    {
        W.ND = W.tm * nbatch;
        W.NT = W.tm * (W.tm + 1) / 2 * nbatch;
    }
    else
    {
        W.ND = 0;
        W.NT = W.tm * W.tn * nbatch;
    }
    W.tot = ((long long)W.ND * W.cd + (long long)(W.NT - W.ND) * W.cf) * W.nkt;
    // one CTA per SM, but at least ~32 full-tile iterations each
    long long G = W.tot / ((long long)W.cf * (512 / kcv));
    if (G < 1) G = 1;
    if (G > num_sms()) G = num_sms();
    if (G > 256) G = 256; // k_tn_fixup resolves one CTA per thread of its block
    W.G = (int)G;
    // tcgen05 kernel, tiles of equal cost: give every tile the same number of CTAs when that
    // idles at most 1/16 of them -- CTA g then works on tile g / ng and K range g % ng, and the NT
    // CTAs of a K range start on the same operand boxes together (reuse through L2)
    if (use_umma && !syrk && W.NT >= 2 && W.G >= W.NT)
    {
        const int gm = W.NT * (W.G / W.NT);
        if ((W.G - gm) * 16 <= W.G) W.G = gm;
    }
    // partial slots per CTA = the most tiles one CTA's share touches
    int smax = 1;
    for (int g = 0; g < W.G; g++)
    {
        long long b0, b1;
        tn_cta_bounds(W, g, b0, b1);
        if (b1 <= b0) continue;
        const int span = tn_tile_of(W, b1 - 1) - tn_tile_of(W, b0) + 1;
        if (span > smax) smax = span;
    }
    W.smax = smax;
    double* partial
        = (double*)scratch(2, (size_t)W.G * W.smax * BM * BN * sizeof(double));
    if (!partial) return MGB_ECUDA;
    if (use_umma)
    {
        // clusters of 2 x 2 tiles sharing their operand boxes by TMA multicast (see tn_umma.cuh):
        // the plan must give every tile the same CTAs (G = NT * ng) and all clusters must be
        // resident at once
        // Off by default: at 128^3 x 512 it cuts the DRAM reads from 19.6 to 14.0 GB (ncu) and a
        // CTA's box time by 12 %, but only 32 clusters of four are resident on the 148 SMs
        // (128 CTAs instead of 144), and the kernel is not bound by DRAM in the first place
        // (tensor pipe 61-63 % active either way): 6.05 vs 6.14 ms.  MGB_TN_MC=1 selects it.
        bool mc = false;
        if (const char* env = getenv("MGB_TN_MC"))
            mc = atoi(env) != 0 && !syrk && ko == 2 && nbatch == 1 && W.tm % 2 == 0 && W.tn % 2 == 0
                 && W.G % W.NT == 0;
        if (mc)
        {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim          = dim3((unsigned)W.G);
            cfg.blockDim         = dim3(umma::NTHR);
            cfg.dynamicSmemBytes = umma::SMEM;
            cudaLaunchAttribute at[1];
            at[0].id               = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 4;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs    = at;
            cfg.numAttrs = 1;
            auto kern = umma::k_gemm_tn_umma<false, true, 2, true>;
            MGB_CUDA(cudaFuncSetAttribute(
                kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)umma::SMEM));
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess)
            {
                (void)cudaGetLastError();
                nclusters = 0;
            }
            const int quads = W.NT / 4;
            const int ng    = nclusters / quads < W.G / W.NT ? nclusters / quads : W.G / W.NT;
            if (ng < 1 || ng * 16 < (W.G / W.NT) * 13) // would idle more than ~1/5 of the SMs
                mc = false;
            else if (ng != W.G / W.NT)
            {
                W.G = W.NT * ng;
                // same share per CTA as before -> same slot count (one tile per CTA)
            }
        }
        CUtensorMap mapA, mapB;
        if (int rc = umma::make_map_kmajor(&mapA, (const float*)A, k, m, lda, strideA, nbatch, ko,
                mc ? 8 : 16))
            return rc;
        if (int rc = umma::make_map_kmajor(&mapB, (const float*)B, k, n, ldb, strideB, nbatch, ko,
                mc ? 8 : 16))
            return rc;
        // chb = boxes per tensor-core accumulation chunk (1: the truncation bias of the FP32 sums
        // stays at 1e-6; MGB_UMMA_CHB: tuning hook)
        int fold = 32, trunc = 1, chb = 1;
        if (const char* env = getenv("MGB_UMMA_FOLD")) fold = atoi(env) > 0 ? atoi(env) : fold;
        if (const char* env = getenv("MGB_UMMA_CHB")) chb = atoi(env) > 0 ? atoi(env) : chb;
        if (const char* env = getenv("MGB_UMMA_TRUNC")) trunc = atoi(env);
#define MGB_UMMA_LAUNCH(SY, TR, KOV)                                                      \
    {                                                                                     \
        auto kern = umma::k_gemm_tn_umma<SY, TR, KOV>;                                    \
        MGB_CUDA(cudaFuncSetAttribute(                                                    \
            kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)umma::SMEM));         \
        kern<<<W.G, umma::NTHR, umma::SMEM, st>>>(mapA, mapB, W, m, n, alpha, beta, C,   \
            ldc, (long long)strideC, partial, fold, chb);                                 \
    }
#define MGB_UMMA_KO(SY, TR)                                                               \
    {                                                                                     \
        if (ko == 2)                                                                      \
            MGB_UMMA_LAUNCH(SY, TR, 2)                                                    \
        else                                                                              \
            MGB_UMMA_LAUNCH(SY, TR, 1)                                                    \
    }
        if (mc)
        {
            auto kern = umma::k_gemm_tn_umma<false, true, 2, true>;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim          = dim3((unsigned)W.G);
            cfg.blockDim         = dim3(umma::NTHR);
            cfg.dynamicSmemBytes = umma::SMEM;
            cfg.stream           = st;
            cudaLaunchAttribute at[1];
            at[0].id               = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 4;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs    = at;
            cfg.numAttrs = 1;
            MGB_CUDA(cudaLaunchKernelEx(&cfg, kern, mapA, mapB, W, m, n, alpha, beta, C, ldc,
                (long long)strideC, partial, fold, chb));
        }
        else if (syrk)
        {
            if (trunc)
                MGB_UMMA_KO(true, true)
            else
                MGB_UMMA_KO(true, false)
        }
        else
        {
            if (trunc)
                MGB_UMMA_KO(false, true)
            else
                MGB_UMMA_KO(false, false)
        }
#undef MGB_UMMA_KO
#undef MGB_UMMA_LAUNCH
        MGB_LAUNCHED("k_gemm_tn_umma");
    }
    else if (tf32)
    {
        const size_t smem32 = (size_t)2 * ST32 * 128 * P32 * sizeof(float);
        if (syrk)
        {
            auto kern = k_gemm_tn_tf32<true>;
            MGB_CUDA(cudaFuncSetAttribute(
                kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32));
            kern<<<W.G, NTHREADS, smem32, st>>>(W, m, n, (long long)k, (const float*)A,
                (long long)lda, (long long)strideA, (const float*)B, (long long)ldb,
                (long long)strideB, alpha, beta, C, ldc, (long long)strideC, partial);
        }
        else
        {
            auto kern = k_gemm_tn_tf32<false>;
            MGB_CUDA(cudaFuncSetAttribute(
                kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32));
            kern<<<W.G, NTHREADS, smem32, st>>>(W, m, n, (long long)k, (const float*)A,
                (long long)lda, (long long)strideA, (const float*)B, (long long)ldb,
                (long long)strideB, alpha, beta, C, ldc, (long long)strideC, partial);
        }
        MGB_LAUNCHED("k_gemm_tn_tf32");
    }
    else
    {
    const size_t smem = (size_t)2 * st_ * 128 * (kcv + PADK) * sizeof(T);
#define MGB_TN_LAUNCH(SY, KV, STV)                                                        \
    {                                                                                     \
        auto kern = k_gemm_tn<T, SY, KV, STV>;                                            \
        MGB_CUDA(cudaFuncSetAttribute(                                                    \
            kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
        kern<<<W.G, NTHREADS, smem, st>>>(W, m, n, (long long)k, A, (long long)lda,      \
            (long long)strideA, B, (long long)ldb, (long long)strideB, alpha, beta, C,   \
            ldc, (long long)strideC, partial);                                            \
    }
    if (syrk)
    {
        if (kcv == 32)
            MGB_TN_LAUNCH(true, 32, 3)
        else
            MGB_TN_LAUNCH(true, 16, 4)
    }
    else
    {
        if (kcv == 32)
            MGB_TN_LAUNCH(false, 32, 3)
        else
            MGB_TN_LAUNCH(false, 16, 4)
    }
#undef MGB_TN_LAUNCH
    MGB_LAUNCHED("k_gemm_tn");
    }
    dim3 fgrid((unsigned)W.NT, 16), fblock(128, 2);
    if (syrk)
        k_tn_fixup<true><<<fgrid, fblock, 0, st>>>(
            W, m, n, partial, alpha, beta, C, ldc, (long long)strideC, sub_shift);
    else
        k_tn_fixup<false><<<fgrid, fblock, 0, st>>>(
            W, m, n, partial, alpha, beta, C, ldc, (long long)strideC, sub_shift);
    MGB_LAUNCHED("k_tn_fixup");
    return MGB_OK;
}

template <typename T>
__global__ void k_axpy_cols(long long m, double gamma, const T* __restrict__ D, long long ldd,
    T* __restrict__ Out, long long ldc)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    const int j = blockIdx.y;
    Out[(long long)j * ldc + p] += (T)(gamma * (double)D[(long long)j * ldd + p]);
}

template <typename T>
static int gemm_nn_t(size_t m, int n, int k, double alpha, const T* A, size_t lda,
    const double* M, int ldm, double beta, T* Out, size_t ldc, cudaStream_t st,
    double gamma = 0., const T* D = nullptr, size_t ldd = 0)
{
    const bool aligned = (((uintptr_t)A | (uintptr_t)M) & 15) == 0
                         && (lda * sizeof(T)) % 16 == 0 && (ldm * sizeof(double)) % 16 == 0;
    if (!aligned)
    {
        dim3 grid((unsigned)((m + 127) / 128), (unsigned)n);
        k_gemm_nn_ref<T><<<grid, 128, 0, st>>>((long long)m, n, k, A, (long long)lda, M,
            ldm, alpha, beta, Out, (long long)ldc);
        MGB_LAUNCHED("k_gemm_nn_ref");
        if (D)
        {
            k_axpy_cols<T><<<grid, 128, 0, st>>>(
                (long long)m, gamma, D, (long long)ldd, Out, (long long)ldc);
            MGB_LAUNCHED("k_axpy_cols");
        }
        return MGB_OK;
    }
    const long long ptiles = (long long)((m + 127) / 128);
    const int jtiles       = (n + 127) / 128;
    const long long nitems = ptiles * jtiles;
    const unsigned grid    = (unsigned)(nitems < num_sms() ? nitems : num_sms());
    if (sizeof(T) == 4 && !g_f32_exact)
    {
        // FP32-class tensor tiles: Mf = (float)(alpha M), K-major with a 16-byte
        // aligned leading dimension
        const int ldf = (k + 3) / 4 * 4;
        float* Mf     = (float*)scratch(8, (size_t)ldf * n * sizeof(float) + 256);
        if (!Mf) return MGB_ECUDA;
        if (ldf == ldm)
        {
            const long long cnt = (long long)ldm * n;
            k_scale_to_float<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(cnt, alpha, M, Mf);
        }
        else
        {
            MGB_CUDA(cudaMemsetAsync(Mf, 0, (size_t)ldf * n * sizeof(float), st));
            for (int j = 0; j < n; j++) // rare: odd numst
                k_scale_to_float<<<(k + 255) / 256, 256, 0, st>>>(
                    k, alpha, M + (size_t)j * ldm, Mf + (size_t)j * ldf);
        }
        MGB_LAUNCHED("k_scale_to_float");
        bool use_umma = g_f32_mode == 0 && ((uintptr_t)A & 15) == 0;
        if (const char* env = getenv("MGB_NN_UMMA")) use_umma = use_umma && atoi(env) != 0;
        if (use_umma)
        {
            // tcgen05 kernel, one CTA per SM: the low parts of Mf, three tensor maps
            int chk = 2; // k-blocks per tensor-core accumulation chunk (MGB_NN_CHK: tuning hook)
            if (const char* env = getenv("MGB_NN_CHK")) chk = atoi(env) > 0 ? atoi(env) : chk;
            float* Ml = (float*)scratch(10, (size_t)ldf * n * sizeof(float) + 256);
            if (!Ml) return MGB_ECUDA;
            const long long cnt = (long long)ldf * n;
            umma::k_tf32_low_parts<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(cnt, Mf, Ml);
            MGB_LAUNCHED("k_tf32_low_parts");
            CUtensorMap mapPhi, mapMh, mapMl;
            if (int rc = umma::make_map_2d(&mapPhi, (const float*)A, m, (size_t)k, lda, 128, 32, false))
                return rc;
            if (int rc = umma::make_map_2d(&mapMh, Mf, (size_t)k, (size_t)n, (size_t)ldf, 32, 128, true))
                return rc;
            if (int rc = umma::make_map_2d(&mapMl, Ml, (size_t)k, (size_t)n, (size_t)ldf, 32, 128, true))
                return rc;
            // clusters of two CTAs that share the coefficient tiles by TMA multicast: one point-
            // tile pair per cluster at a time (MGB_NN_MC=0: independent CTAs)
            int mc = ptiles >= 2 ? 1 : 0;
            if (const char* env = getenv("MGB_NN_MC")) mc = mc && atoi(env) != 0;
            unsigned grid_mc = (unsigned)(num_sms() & ~1);
            if ((long long)grid_mc > 2 * ((ptiles + 1) / 2)) grid_mc = (unsigned)(2 * ((ptiles + 1) / 2));
#define MGB_NN_UMMA(HD)                                                                   \
    if (mc)                                                                               \
    {                                                                                     \
        auto kern = umma::k_gemm_nn_umma<HD, true>;                                       \
        MGB_CUDA(cudaFuncSetAttribute(                                                    \
            kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)umma::NN_SMEM));      \
        cudaLaunchConfig_t cfg = {};                                                      \
        cfg.gridDim          = dim3(grid_mc);                                             \
        cfg.blockDim         = dim3(umma::NTHR);                                          \
        cfg.dynamicSmemBytes = umma::NN_SMEM;                                             \
        cfg.stream           = st;                                                        \
        cudaLaunchAttribute at[1];                                                        \
        at[0].id               = cudaLaunchAttributeClusterDimension;                     \
        at[0].val.clusterDim.x = 2;                                                       \
        at[0].val.clusterDim.y = 1;                                                       \
        at[0].val.clusterDim.z = 1;                                                       \
        cfg.attrs    = at;                                                                \
        cfg.numAttrs = 1;                                                                 \
        MGB_CUDA(cudaLaunchKernelEx(&cfg, kern, mapPhi, mapMh, mapMl, (long long)m, n, k, \
            beta, (float*)Out, (long long)ldc, nitems, jtiles, gamma, (const float*)D,    \
            (long long)ldd, chk));                                                        \
    }                                                                                     \
    else                                                                                  \
    {                                                                                     \
        auto kern = umma::k_gemm_nn_umma<HD, false>;                                      \
        MGB_CUDA(cudaFuncSetAttribute(                                                    \
            kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)umma::NN_SMEM));      \
        kern<<<grid, umma::NTHR, umma::NN_SMEM, st>>>(mapPhi, mapMh, mapMl, (long long)m, \
            n, k, beta, (float*)Out, (long long)ldc, nitems, jtiles, gamma,               \
            (const float*)D, (long long)ldd, chk);                                        \
    }
            if (D)
                MGB_NN_UMMA(true)
            else
                MGB_NN_UMMA(false)
#undef MGB_NN_UMMA
            MGB_LAUNCHED("k_gemm_nn_umma");
            return MGB_OK;
        }
        const size_t smem32
            = (size_t)ST32 * 128 * P32 * sizeof(float) + (size_t)ST32 * KC32 * PP32 * sizeof(float);
        if (D)
        {
            auto kern = k_gemm_nn_tf32<true>;
            MGB_CUDA(cudaFuncSetAttribute(
                kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32));
            kern<<<grid, NTHREADS, smem32, st>>>((long long)m, n, k, (const float*)A,
                (long long)lda, Mf, ldf, beta, (float*)Out, (long long)ldc, nitems, jtiles, gamma,
                (const float*)D, (long long)ldd);
        }
        else
        {
            auto kern = k_gemm_nn_tf32<false>;
            MGB_CUDA(cudaFuncSetAttribute(
                kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32));
            kern<<<grid, NTHREADS, smem32, st>>>((long long)m, n, k, (const float*)A,
                (long long)lda, Mf, ldf, beta, (float*)Out, (long long)ldc, nitems, jtiles, gamma,
                (const float*)D, (long long)ldd);
        }
        MGB_LAUNCHED("k_gemm_nn_tf32");
        return MGB_OK;
    }
    int kcv = 32;
    if (const char* env = getenv("MGB_NN_KC")) kcv = atoi(env) == 16 ? 16 : 32;
    const int st_     = kcv == 32 ? 3 : 4;
    const size_t smem = (size_t)st_ * 128 * (kcv + PADK) * sizeof(double)
                        + (size_t)st_ * kcv * PitchP<T>::value * sizeof(T);
#define MGB_NN_LAUNCH(KV, STV, HD)                                                        \
    {                                                                                     \
        auto kern = k_gemm_nn<T, KV, STV, HD>;                                            \
        MGB_CUDA(cudaFuncSetAttribute(                                                    \
            kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
        kern<<<grid, NTHREADS, smem, st>>>((long long)m, n, k, A, (long long)lda, M, ldm, \
            alpha, beta, Out, (long long)ldc, nitems, jtiles, gamma, D, (long long)ldd); \
    }
    if (kcv == 32)
    {
        if (D)
            MGB_NN_LAUNCH(32, 3, true)
        else
            MGB_NN_LAUNCH(32, 3, false)
    }
    else
    {
        if (D)
            MGB_NN_LAUNCH(16, 4, true)
        else
            MGB_NN_LAUNCH(16, 4, false)
    }
#undef MGB_NN_LAUNCH
    MGB_LAUNCHED("k_gemm_nn");
    return MGB_OK;
}

int gemm_nn_fused(int dtype, size_t m, int n, int k, double alpha, const void* A, size_t lda,
    const double* M, int ldm, double beta, void* Out, size_t ldc, double gamma, const void* D,
    size_t ldd, cudaStream_t st)
{
    if (dtype == MGB_F64)
        return gemm_nn_t<double>(m, n, k, alpha, (const double*)A, lda, M, ldm, beta,
            (double*)Out, ldc, st, gamma, (const double*)D, ldd);
    return gemm_nn_t<float>(m, n, k, alpha, (const float*)A, lda, M, ldm, beta, (float*)Out,
        ldc, st, gamma, (const float*)D, ldd);
}

} // namespace mgb

using namespace mgb;

extern "C"
{

// Host-only view of the stream-K decomposition (no device needed): the segments
// (cta, tile, first k-iteration, one-past-last) in launch order.  Test hook.
int mgb_debug_tn_plan(int syrk, int m, int n, size_t k, int nbatch, int kc, int ncta, int cd,
    long long* segs, int max_segs, int* nsegs, long long* nkt_out, int* ntiles_out, int* ndiag_out)
{
    MGB_REQUIRE(segs && nsegs && m > 0 && n > 0 && k > 0 && nbatch > 0 && kc > 0 && ncta > 0,
        "mgb_debug_tn_plan: bad arguments");
    TnWork W;
    W.tm     = (m + BM - 1) / BM;
    W.tn     = (n + BN - 1) / BN;
    W.nbatch = nbatch;
    W.nkt    = (long long)((k + kc - 1) / kc);
    W.cf     = 16;
    W.cd     = cd;
    if (syrk)
    {
        W.ND = W.tm * nbatch;
        W.NT = W.tm * (W.tm + 1) / 2 * nbatch;
    }
    else
    {
        W.ND = 0;
        W.NT = W.tm * W.tn * nbatch;
    }
    W.tot  = ((long long)W.ND * W.cd + (long long)(W.NT - W.ND) * W.cf) * W.nkt;
    W.G    = ncta;
    W.smax = 0;
    int cnt = 0;
    for (int g = 0; g < W.G; g++)
    {
        long long b0, b1;
        tn_cta_bounds(W, g, b0, b1);
        if (b1 <= b0) continue;
        const int uf = tn_tile_of(W, b0), ul = tn_tile_of(W, b1 - 1);
        for (int u = uf; u <= ul; u++)
        {
            long long it0, it1;
            tn_seg(W, u, b0, b1, it0, it1);
            if (it0 >= it1) continue;
            if (cnt < max_segs)
            {
                segs[4 * cnt + 0] = g;
                segs[4 * cnt + 1] = u;
                segs[4 * cnt + 2] = it0;
                segs[4 * cnt + 3] = it1;
            }
            cnt++;
        }
    }
    *nsegs = cnt;
    if (nkt_out) *nkt_out = W.nkt;
    if (ntiles_out) *ntiles_out = W.NT;
    if (ndiag_out) *ndiag_out = W.ND;
    return MGB_OK;
}

int mgb_set_f32_contraction(int mode)
{
    MGB_REQUIRE(mode >= 0 && mode <= 2, "mgb_set_f32_contraction: mode %d", mode);
    g_f32_mode = mode;
    return MGB_OK;
}

int mgb_gemm_tn(int dtype, int m, int n, size_t k, double alpha, const void* A,
    size_t lda, const void* B, size_t ldb, double beta, double* C, int ldc, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(A && B && C, "mgb_gemm_tn: null pointer");
    MGB_REQUIRE(m >= 0 && n >= 0 && ldc >= m && lda >= k && ldb >= k,
        "mgb_gemm_tn: bad dimensions");
    if (m == 0 || n == 0) return MGB_OK;
    if (dtype == MGB_F64)
        return gemm_tn_t<double>(false, m, n, k, alpha, (const double*)A, lda, 0,
            (const double*)B, ldb, 0, beta, C, ldc, 0, 1, as_stream(stream));
    if (dtype == MGB_F32)
        return gemm_tn_t<float>(false, m, n, k, alpha, (const float*)A, lda, 0,
            (const float*)B, ldb, 0, beta, C, ldc, 0, 1, as_stream(stream));
    set_error("mgb_gemm_tn: bad dtype");
    return MGB_EINVAL;
}

int mgb_gemm_tn_slabs(int dtype, int m, int n, size_t k, int nslabs, double alpha,
    const void* A, size_t lda, const void* B, size_t ldb, double beta, double* C, int ldc,
    void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(A && B && C, "mgb_gemm_tn_slabs: null pointer");
    MGB_REQUIRE(m >= 0 && n >= 0 && nslabs >= 1 && ldc >= m && lda >= k * nslabs
                    && ldb >= k * nslabs,
        "mgb_gemm_tn_slabs: bad dimensions");
    if (m == 0 || n == 0) return MGB_OK;
    const size_t sc = (size_t)ldc * n;
    if (dtype == MGB_F64)
        return gemm_tn_t<double>(false, m, n, k, alpha, (const double*)A, lda, k,
            (const double*)B, ldb, k, beta, C, ldc, sc, nslabs, as_stream(stream));
    if (dtype == MGB_F32)
        return gemm_tn_t<float>(false, m, n, k, alpha, (const float*)A, lda, k,
            (const float*)B, ldb, k, beta, C, ldc, sc, nslabs, as_stream(stream));
    set_error("mgb_gemm_tn_slabs: bad dtype");
    return MGB_EINVAL;
}

int mgb_syrk_t_slabs(int dtype, int n, size_t k, int nslabs, double alpha, const void* A,
    size_t lda, double* C, int ldc, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(A && C, "mgb_syrk_t_slabs: null pointer");
    MGB_REQUIRE(n >= 0 && nslabs >= 1 && ldc >= n && lda >= k * nslabs,
        "mgb_syrk_t_slabs: bad dimensions");
    if (n == 0) return MGB_OK;
    const size_t sc = (size_t)ldc * n;
    if (dtype == MGB_F64)
        return gemm_tn_t<double>(true, n, n, k, alpha, (const double*)A, lda, k,
            (const double*)A, lda, k, 0., C, ldc, sc, nslabs, as_stream(stream));
    if (dtype == MGB_F32)
        return gemm_tn_t<float>(true, n, n, k, alpha, (const float*)A, lda, k,
            (const float*)A, lda, k, 0., C, ldc, sc, nslabs, as_stream(stream));
    set_error("mgb_syrk_t_slabs: bad dtype");
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
        return gemm_tn_t<double>(true, n, n, k, alpha, (const double*)A, lda, 0,
            (const double*)A, lda, 0, 0., C, ldc, 0, 1, as_stream(stream));
    if (dtype == MGB_F32)
        return gemm_tn_t<float>(true, n, n, k, alpha, (const float*)A, lda, 0,
            (const float*)A, lda, 0, 0., C, ldc, 0, 1, as_stream(stream));
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
