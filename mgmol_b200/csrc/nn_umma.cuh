// Out(npt x n, float) = Phi(npt x k) * Mf(k x n) + beta Out + gamma D for ORBDTYPE float on the
// 5th-generation tensor cores (ExtendedGridOrbitals::multiplyByMatrix, src/ExtendedGridOrbitals.cc:
// 448-498; arithmetic of MPgemmNN, src/linear_algebra/mputils.cc:848-948).  Included by
// contractions.cu after tn_umma.cuh, whose helpers (descriptors, tcgen05 wrappers, the 3xTF32
// split) it shares; Mf = (float)(alpha M) and its low parts Ml = Mf - trunc(Mf) are prepared by
// the host wrapper (K-major: element (l, j) at l + j*ldf, i.e. the reference's column-major M).
//
// An item is 128 grid points x 128 output orbitals.  In tcgen05 terms D[128 points x 128] =
// A[128 points x K = k] * B[128 j x K]^T:
//   A  the Phi tile.  In memory the POINTS are contiguous (an MN-major operand), but A comes from
//      TMEM, where layout is not an issue: the TMA lands the {128 points, 32 orbitals} box as 32
//      rows of 512 B, split thread p reads element (p, l) for its 32 l's (conflict-free: lanes
//      are consecutive points) and stores a_hi, a_lo into TMEM lane p with tcgen05.st.
//   B  the 128 x 32 tiles of Mf and Ml, K-major under the 128-byte swizzle, straight from the
//      TMA (L2-resident, evict_last); b_hi is the raw tile (the tensor core ignores the 13 low
//      mantissa bits).  Having the split warps derive the low tile instead saves a third of the
//      L2 -> SM traffic but was slower (1.91 vs 1.74 ms at 128^3 x 256: the split warps are the
//      tighter resource).
// Per 32-orbital k-block: one 48 KB stage (Phi box, Mf tile, Ml tile), 12 MMAs (4 k8 steps x
// {a_lo b_hi, a_hi b_lo, a_hi b_hi}).  The tensor core's truncating FP32 sums run over CHK = 2
// k-blocks (4: 7 % faster, twice the truncation bias); chunk sums are added round-to-nearest in FP32 registers (K = numst is short: the
// reference's double sum differs by < 1e-6 of |a||m|).  Persistent CTAs walk the items with the
// ring running across item boundaries; the two accumulators alternate by chunk, so an item's
// epilogue (beta Out, gamma D, coalesced stores: a warp stores 32 consecutive points of one
// orbital) overlaps the next item's MMAs.
//
// MC (cluster of two CTAs).  The coefficient tiles are two thirds of what a CTA pulls through
// the L2 -> SM fabric (12.9 GB of TMA traffic for a 4.3 GB block at 128^3 x 256, ncu), and every
// CTA that works on the same column tile wants the same ones.  Two CTAs of a cluster take two
// neighbouring point tiles, walk the column tiles and k-blocks in the same order, and split the
// coefficient loads: rank 0 loads the Mf tile, rank 1 the Ml tile, each with .multicast::cluster
// into both CTAs' stage; a stage is released by both MMA warps (multicast commit, count 2).
//
// Roles (14 warps): 0 TMA producer, 1 MMA issuer (warp-converged, one elected lane), 2-5 split,
// 6-13 drain / epilogue (warp w: TMEM lane quarter w % 4, 64 of the 128 columns).
#pragma once

namespace umma
{
constexpr int NN_ST    = 4;                 // ring depth: smem stage s <-> TMEM A stage s
constexpr int NN_STAGE = 3 * TILE_B;        // Phi box, Mf tile, Ml tile
constexpr size_t NN_SMEM = (size_t)NN_ST * NN_STAGE + 1024;

__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0,
    int c1, uint16_t mask, uint64_t policy)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
                 ".multicast::cluster.L2::cache_hint [%0], [%1, {%4, %5}], [%2], %3, %6;" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "l"(policy)
                 : "memory");
}
// The items of a CTA.  !MC: items blockIdx.x, + gridDim.x, ... of the (point tile, column tile)
// list, column tile fastest.  MC: cluster c = blockIdx.x / 2 takes the point-tile pairs c, c + NC,
// ...; rank r = blockIdx.x % 2 the r-th tile of the pair (clamped to the last tile when the count
// is odd: that CTA computes a duplicate and stores nothing), all column tiles in turn.
template <bool MC>
struct NnItems
{
    long long ptiles, nitems;
    int jtiles;
    long long cur, step; // !MC: item; MC: pair index
    int jt;
    int rank;
    __device__ __forceinline__ NnItems(long long nitems_, int jtiles_) : nitems(nitems_), jtiles(jtiles_)
    {
        ptiles = nitems / jtiles;
        if (MC)
        {
            cur  = blockIdx.x >> 1;
            step = gridDim.x >> 1;
            rank = blockIdx.x & 1;
        }
        else
        {
            cur  = blockIdx.x;
            step = gridDim.x;
            rank = 0;
        }
        jt = 0;
    }
    __device__ __forceinline__ bool valid() const { return MC ? 2 * cur < ptiles : cur < nitems; }
    __device__ __forceinline__ long long ptile() const
    {
        if (!MC) return cur / jtiles;
        const long long p = 2 * cur + rank;
        return p < ptiles ? p : ptiles - 1;
    }
    __device__ __forceinline__ bool stores() const { return !MC || 2 * cur + rank < ptiles; }
    __device__ __forceinline__ int jtile() const { return MC ? jt : (int)(cur % jtiles); }
    __device__ __forceinline__ void next()
    {
        if (MC)
        {
            if (++jt == jtiles)
            {
                jt = 0;
                cur += step;
            }
        }
        else
            cur += step;
    }
};

template <bool HASD, bool MC>
__global__ void __launch_bounds__(NTHR, 1) k_gemm_nn_umma(const __grid_constant__ CUtensorMap mapPhi,
    const __grid_constant__ CUtensorMap mapMh, const __grid_constant__ CUtensorMap mapMl,
    long long npt, int n, int k, double beta, float* __restrict__ Out, long long ldc,
    long long nitems, int jtiles, double gamma, const float* __restrict__ D, long long ldd, int NN_CHK)
{
    extern __shared__ uint8_t smraw[];
    __shared__ __align__(8) uint64_t full[NN_ST], conv[NN_ST], empty[NN_ST], accf[2], acce[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* ring = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    const int nkb = (k + 31) / 32;                 // k-blocks per item
    const int nch = (nkb + NN_CHK - 1) / NN_CHK;   // chunks per item
    if (!NnItems<MC>(nitems, jtiles).valid()) return; // never in a cluster launch (grid <= pairs)

    if (tid == 0)
    {
        for (int s = 0; s < NN_ST; s++)
        {
            mbar_init(&full[s], 1);
            mbar_init(&conv[s], 4);
            mbar_init(&empty[s], MC ? 2 : 1);
        }
        for (int b = 0; b < 2; b++)
        {
            mbar_init(&accf[b], 1);
            mbar_init(&acce[b], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_base_s)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    // the partner's barriers must exist before anything is multicast into its stages
    if (MC) cluster_sync_all();
    fence_after();
    const uint32_t tmem   = tmem_base_s;
    const uint32_t tmem_a = tmem + 2 * ACC_COLS; // NN_ST stages of 64 columns

    if (warp == 0)
    {
        // ---------------- TMA producer ----------------
        if (lane == 0)
        {
            const uint64_t pol_phi = policy_evict_normal(), pol_m = policy_evict_last();
            uint32_t gs = 0;
            for (NnItems<MC> I(nitems, jtiles); I.valid(); I.next())
            {
                const int j0       = I.jtile() * 128;
                const long long p0 = I.ptile() * 128;
                for (int kb = 0; kb < nkb; kb++, gs++)
                {
                    const int s = gs % NN_ST;
                    if (gs >= NN_ST) mbar_wait(&empty[s], ((gs / NN_ST) & 1) ^ 1);
                    uint8_t* st = ring + (size_t)s * NN_STAGE;
                    mbar_arrive_expect_tx(&full[s], NN_STAGE);
                    tma_load_2d(st, &mapPhi, &full[s], (int)p0, kb * 32, pol_phi);
                    if (!MC)
                    {
                        tma_load_2d(st + TILE_B, &mapMh, &full[s], kb * 32, j0, pol_m);
                        tma_load_2d(st + 2 * TILE_B, &mapMl, &full[s], kb * 32, j0, pol_m);
                    }
                    else if (I.rank == 0)
                        tma_load_2d_mc(st + TILE_B, &mapMh, &full[s], kb * 32, j0, (uint16_t)3, pol_m);
                    else
                        tma_load_2d_mc(st + 2 * TILE_B, &mapMl, &full[s], kb * 32, j0, (uint16_t)3, pol_m);
                }
            }
        }
    }
    else if (warp == 1)
    {
        // ---------------- MMA issuer (warp-converged, see tn_umma.cuh) ----------------
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
        const uint32_t ring_u = __shfl_sync(0xffffffffu, smem_u32(ring), 0);
        uint32_t gs = 0, gc = 0;
        for (NnItems<MC> I(nitems, jtiles); I.valid(); I.next())
        {
            for (int c = 0; c < nch; c++, gc++)
            {
                const int ab = gc & 1;
                if (gc >= 2) mbar_wait(&acce[ab], ((gc >> 1) & 1) ^ 1);
                const uint32_t tacc = tmem_u + (uint32_t)(ab * ACC_COLS);
                const int kb1       = (c + 1) * NN_CHK < nkb ? (c + 1) * NN_CHK : nkb;
                for (int kb = c * NN_CHK; kb < kb1; kb++, gs++)
                {
                    const int s = gs % NN_ST;
                    mbar_wait(&conv[s], (gs / NN_ST) & 1);
                    fence_after();
                    const uint32_t sa  = ring_u + (uint32_t)(s * NN_STAGE);
                    const uint64_t dbh = smem_desc(sa + TILE_B, 1024);
                    const uint64_t dbl = smem_desc(sa + 2 * TILE_B, 1024);
                    const uint32_t ah  = tmem_u + 2 * ACC_COLS + (uint32_t)(s * 64);
                    if (elect_one())
                    {
#pragma unroll
                        for (int ks = 0; ks < 4; ks++)
                        {
                            const uint64_t adv = (uint64_t)(ks * 2); // 32 bytes along K
                            const uint32_t ak  = ah + (uint32_t)(ks * 8);
                            mma_tf32_ts(tacc, ak + 32, dbh + adv, (kb > c * NN_CHK || ks > 0) ? 1u : 0u);
                            mma_tf32_ts(tacc, ak, dbl + adv, 1u);
                            mma_tf32_ts(tacc, ak, dbh + adv, 1u);
                        }
                        if (MC)
                            commit_mc(&empty[s], (uint16_t)3);
                        else
                            commit(&empty[s]);
                        if (kb == kb1 - 1) commit(&accf[ab]);
                    }
                    __syncwarp();
                }
            }
        }
    }
    else if (warp < 6)
    {
        // ---------------- split warps: Phi box -> a_hi, a_lo in TMEM ----------------
        const int p = (warp & 3) * 32 + lane; // point of the tile = TMEM lane
        uint32_t gs = 0;
        for (NnItems<MC> I(nitems, jtiles); I.valid(); I.next())
        {
            for (int kb = 0; kb < nkb; kb++, gs++)
            {
                const int s = gs % NN_ST;
                mbar_wait(&full[s], (gs / NN_ST) & 1);
                fence_after();
                const float* box = reinterpret_cast<const float*>(ring + (size_t)s * NN_STAGE);
                uint32_t h[32], l[32];
#pragma unroll
                for (int q = 0; q < 32; q++)
                {
                    const float x = box[q * 128 + p];
                    h[q]          = __float_as_uint(x) & 0xFFFFE000u;
                    l[q]          = __float_as_uint(x - __uint_as_float(h[q]));
                }
                const uint32_t taddr = tmem_a + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(s * 64);
                st32(taddr, h);
                st32(taddr + 32, l);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&conv[s]);
            }
        }
    }
    else
    {
        // ---------------- drain / epilogue warps ----------------
        const int q    = warp & 3;
        const int half = (warp - 6) >> 2;
        const int pl   = q * 32 + lane;
        uint32_t gc    = 0;
        for (NnItems<MC> I(nitems, jtiles); I.valid(); I.next())
        {
            const int j0       = I.jtile() * 128 + half * 64;
            const long long pp = I.ptile() * 128 + pl;
            float hi[64];
#pragma unroll
            for (int j = 0; j < 64; j++) hi[j] = 0.f;
            for (int c = 0; c < nch; c++, gc++)
            {
                const int ab = gc & 1;
                mbar_wait(&accf[ab], (gc >> 1) & 1);
                fence_after();
                const uint32_t taddr
                    = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * ACC_COLS + half * 64);
#pragma unroll
                for (int h2 = 0; h2 < 2; h2++)
                {
                    uint32_t v[32];
                    ld32(taddr + 32 * h2, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        hi[32 * h2 + j] = __fadd_rn(hi[32 * h2 + j], __uint_as_float(v[j]));
                }
                fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acce[ab]);
            }
            if (pp < npt && I.stores())
            {
                // the epilogue sits between two drains of the same warps: keep it short (the
                // first version spent 22 instructions per column on addressing and tests)
                float* o       = Out + (long long)j0 * ldc + pp;
                const float* d = HASD ? D + (long long)j0 * ldd + pp : nullptr;
                if (n - j0 >= 64 && beta == 0.)
                {
#pragma unroll
                    for (int j = 0; j < 64; j++)
                    {
                        float r = hi[j];
                        if (HASD) r += (float)(gamma * (double)d[(long long)j * ldd]);
                        o[(long long)j * ldc] = r;
                    }
                }
                else
                {
#pragma unroll
                    for (int j = 0; j < 64; j++)
                        if (j0 + j < n)
                        {
                            float r = hi[j];
                            if (beta != 0.) r += (float)(beta * (double)o[(long long)j * ldc]);
                            if (HASD) r += (float)(gamma * (double)d[(long long)j * ldd]);
                            o[(long long)j * ldc] = r;
                        }
                }
            }
        }
    }

    fence_before();
    __syncthreads();
    // nobody leaves while the partner may still multicast into this CTA or arrive on its barriers
    if (MC) cluster_sync_all();
    if (warp == 1)
    {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u)
                     : "memory");
    }
}

// Ml = Mf - trunc_tf32(Mf): the low parts of the coefficient matrix, once per call
__global__ void k_tf32_low_parts(long long cnt, const float* __restrict__ in, float* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) out[i] = in[i] - __uint_as_float(__float_as_uint(in[i]) & 0xFFFFE000u);
}

// 2-D maps: Phi (points contiguous; box {128 points, 32 orbitals}, no swizzle) and the K-major
// coefficient tiles (box {32 l, 128 j}, 128-byte swizzle); out-of-range parts are zero-filled
static int make_map_2d(CUtensorMap* mp, const float* base, size_t d0, size_t d1, size_t ld, int b0,
    int b1, bool swizzle)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc)
    {
        set_error("cuTensorMapEncodeTiled entry point not found");
        return MGB_ECUDA;
    }
    cuuint64_t dims[2]    = { (cuuint64_t)d0, (cuuint64_t)d1 };
    cuuint64_t strides[1] = { (cuuint64_t)ld * 4 };
    cuuint32_t box[2]     = { (cuuint32_t)b0, (cuuint32_t)b1 };
    cuuint32_t estr[2]    = { 1, 1 };
    CUresult r = enc(mp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box,
        estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
    {
        set_error("cuTensorMapEncodeTiled failed (%d): 2-D map %zu x %zu ld %zu", (int)r, d0, d1, ld);
        return MGB_ECUDA;
    }
    return MGB_OK;
}

} // namespace umma
