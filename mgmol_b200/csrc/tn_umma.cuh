// ORBDTYPE float contractions C = alpha A^T B + beta C on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM), included by contractions.cu after the
// stream-K plan (TnWork, tn_*): same (tile, K) decomposition, same partial-slot format and the
// same fix-up kernel as the DMMA / mma.sync kernels -- only the inner loop differs.
//
// Arithmetic (src/linear_algebra/mputils.cc:848-948 sums (double)a*(double)b): error-compensated
// 3xTF32.  Every operand is split a = a_hi + a_lo into two TF32 numbers and a*b is formed as
// a_lo b_hi + a_hi b_lo + a_hi b_hi (the dropped a_lo b_lo is 2^-22 of |a||b|).  The tensor
// core sums in FP32 -- truncating, measured: a sum of S slabs of positive terms comes out low by
// ~5e-7 S -- over one chunk only (one TMA box = 2 slabs by default); chunk sums are added in FP32
// registers (round to nearest) over at most FOLD chunks and those in double (the segment's
// partial slot, L2-resident), so the K = 10^6..10^7 reduction behaves like a double sum.
//
// One CTA per SM, 14 warps, each with one role:
//   warp 0      TMA producer: raw 128 x 32 float tiles of A and B (K-major, 128-byte swizzle =
//               the canonical UMMA K-major SWIZZLE_128B atoms), KO consecutive slabs per box
//   warps 2-5   split: one thread per row of A (= TMEM lane) reads the raw row (de-swizzled
//               16-byte chunks, conflict-free) and stores a_hi, a_lo with tcgen05.st into a ring
//               of TMEM columns; B's low parts go to a twin tile in shared memory at the SAME
//               offsets (the swizzle never has to be undone); B's high part is the raw tile
//               itself (TRUNC: the tensor core ignores the 13 low mantissa bits, so b_hi =
//               trunc(b) costs nothing and b_lo = b - trunc(b) is exact) or is rounded in place
//   warp 1      walks the loops converged; one elected lane issues the 12 tcgen05.mma of a slab
//               (4 k8 steps x 3 products; A from TMEM, B from shared memory) and the commits that
//               hand the stages back; owns the TMEM allocation
//   warps 6-13  drain: tcgen05.ld of a finished chunk accumulator (two 128-column accumulators
//               alternate, so the drain overlaps the next chunk's MMAs), the two upper summation
//               levels, the partial slot / direct write of C
//
// What bounds it (B200, 128^3 x 256, measured with the kernel's pieces switched off one by one,
// gpurun_out/r02_umma_dbg*.log):
//   * A in TMEM: with both operands in shared memory a 128 x 128 x 8 TF32 MMA reads 8 KB per
//     64 cycles, the whole shared-memory bandwidth of the SM.
//   * Address translation: a tile's 128 rows are 128 orbitals, `ld` floats apart -- for a 128^3
//     block 8 MB, i.e. every 128-byte row of a TMA box lies in its own 2 MB page.  With one slab
//     per box the kernel ran at 1.08 us per slab whatever else it did (0.46 us with ld = 1 MB,
//     where two rows share a page; padding ld changed nothing).  The boxes are therefore 5-D
//     {32 k, 8 rows, KO slabs, 16 row groups, slab}: per row group the TMA walks 8 rows x KO
//     consecutive 128-byte pieces, so a translation serves KO requests, and the box lands as
//     [row group][ko][8 rows][128 B] -- each slab's tile is a regular SWIZZLE_128B tile with
//     its 8-row atoms KO KB apart (the descriptor's stride byte offset).
//   * L2 -> SM: 32 KB per slab and SM is ~0.45 us at the fabric's per-SM share; the MMA floor of
//     a slab (12 x 64 cycles) is 0.40 us at 1.9 GHz.
//   * TMEM: tcgen05.ld moves ~64 B/clk per SM, so draining a 64 KB accumulator takes ~1000
//     cycles, and drains and MMAs take turns: a box is 1536 cycles of MMAs + ~1000 of drain, which
//     is what the kernel runs at (tensor pipe 61-63 % active, ncu).  Longer chunks would halve
//     that share and double the truncation bias.
#pragma once

namespace umma
{
constexpr int TILE_B   = 128 * 128;   // bytes of one 128-row x 32-float tile
constexpr int NTHR     = 448;         // 14 warps
constexpr int ACC_COLS = 128;         // TMEM columns of one accumulator
// split stages: B low in shared memory, A high / low in TMEM (4 with a 2-box raw ring: slower)
constexpr int CST      = 2;
constexpr size_t SMEM  = (size_t)224 * 1024 + 1024;
// raw ring: NSS boxes of KO slabs, A and B each (NSS * KO * 32 KB) + CST low tiles of 16 KB
template <int KO>
struct Ring
{
    static constexpr int NSS = (224 - 16 * CST) / (32 * KO);
    static_assert(KO == 1 || KO == 2, "KO");
    static_assert(NSS >= 2 && CST <= 4, "ring");
};

// UMMA::SmemDescriptor of a K-major SWIZZLE_128B tile (8-row x 128-byte atoms, `sbo` bytes
// apart): start address [0,14) >> 4, leading byte offset [16,30) = 1, stride byte offset
// [32,46) = sbo >> 4, version [46,48) = 1, layout type [61,64) = 2.  Pinned on the device by
// tools/microbench/umma_tf32_probe.cu.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// UMMA::InstrDescriptor: D = F32 [4,6), A = B = TF32 [7,10) [10,13), both K-major, N >> 3 at
// [17,23), M >> 4 at [24,29): M = N = 128
constexpr uint32_t IDESC
    = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// one lane of a converged warp
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t"
                 ".reg .pred P;\n\t"
                 "elect.sync _|P, 0xffffffff;\n\t"
                 "selp.u32 %0, 1, 0, P;\n\t"
                 "}\n"
                 : "=r"(pred));
    return pred != 0;
}

// clusters: a commit that arrives on the same barrier of every CTA in `mask`, a multicast TMA
// load, and the cluster-wide barrier
__device__ __forceinline__ void commit_mc(uint64_t* bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
                 "[%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tma_load_5d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0,
    int c1, int c2, int c3, int c4, uint16_t mask, uint64_t policy)
{
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
                 ".multicast::cluster.L2::cache_hint [%0], [%1, {%4, %5, %6, %7, %8}], [%2], %3, %9;" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4),
                 "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}


// D[tmem_c] (+)= A[tmem_a] * B[db]^T, A from TMEM (lane = row, one k per 32-bit column)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t db, uint32_t acc)
{
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
                 "}\n" ::"r"(tmem_c),
                 "r"(tmem_a), "l"(db), "r"(IDESC), "r"(acc)
                 : "memory");
}
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ float rna_tf32(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// split one raw tile: low parts to `lo`, same offsets; !TRUNC also rounds the raw tile in place

__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}

// B's low parts: raw tile (8-row atoms `sbo` bytes apart) -> dense low tile (atoms 1 KB apart),
// same offsets inside an atom; !TRUNC also rounds the raw tile in place.  128 threads.
template <bool TRUNC>
__device__ __forceinline__ void split_tile(uint8_t* raw, uint32_t sbo, uint8_t* lo, int ctid)
{
#pragma unroll
    for (int j = 0; j < TILE_B / 16 / 128; j++)
    {
        const int c    = ctid + 128 * j;          // 16-byte chunk of the tile
        const int atom = c >> 6, off = (c & 63) * 16;
        uint8_t* src   = raw + (size_t)atom * sbo + off;
        float4 v       = *reinterpret_cast<const float4*>(src);
        float4 l;
        if (TRUNC)
        {
            l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
            l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
            l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
            l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
        }
        else
        {
            float4 h;
            h.x = rna_tf32(v.x), h.y = rna_tf32(v.y), h.z = rna_tf32(v.z), h.w = rna_tf32(v.w);
            l.x = rna_tf32(v.x - h.x), l.y = rna_tf32(v.y - h.y);
            l.z = rna_tf32(v.z - h.z), l.w = rna_tf32(v.w - h.w);
            *reinterpret_cast<float4*>(src) = h;
        }
        *reinterpret_cast<float4*>(lo + atom * 1024 + off) = l;
    }
}

// row `row` of a raw tile -> a_hi, a_lo in TMEM lane `row`, columns [col, col+32) and
// [col+32, col+64).  TRUNC: a_hi = a with the 13 low mantissa bits cleared (what the tensor core
// would read anyway), a_lo = a - a_hi exactly: two instructions per element -- the split warps
// are one warp per SM sub-partition (cvt.rna.tf32 is a four-instruction sequence)
// TWICE (diagonal Gram tiles): a_lo is stored doubled
template <bool TRUNC, bool TWICE>
__device__ __forceinline__ void split_row_to_tmem(const uint8_t* raw, uint32_t sbo, int row, uint32_t taddr)
{
    const uint8_t* rb = raw + (size_t)(row >> 3) * sbo + (row & 7) * 128;
    uint32_t h[32], l[32];
#pragma unroll
    for (int c = 0; c < 8; c++)
    {
        const float4 v = *reinterpret_cast<const float4*>(rb + ((c ^ (row & 7)) << 4));
        const float x[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int e = 0; e < 4; e++)
        {
            if (TRUNC)
            {
                h[4 * c + e]  = __float_as_uint(x[e]) & 0xFFFFE000u;
                const float d = x[e] - __uint_as_float(h[4 * c + e]);
                l[4 * c + e]  = __float_as_uint(TWICE ? d + d : d);
            }
            else
            {
                const float hh = rna_tf32(x[e]);
                h[4 * c + e]   = __float_as_uint(hh);
                const float d  = rna_tf32(x[e] - hh);
                l[4 * c + e]   = __float_as_uint(TWICE ? d + d : d);
            }
        }
    }
    st32(taddr, h);
    st32(taddr + 32, l);
}

// Drain warps, one (tile, K range) segment: chunk accumulators out of TMEM as the MMA warp
// finishes them (two alternate), FP32 sums of up to FOLD chunks in registers, double sums in the
// segment's partial slot.  Thread = row ml of the tile, 64 columns (half).
template <bool SYRK>
__device__ __forceinline__ void drain_segment(double* __restrict__ dst, uint32_t tmem, uint64_t* accf,
    uint64_t* acce, uint32_t& gc, int nit, int CH, int FOLD, int q, int half, int ml, int lane)
{
    float hi[64];
#pragma unroll
    for (int j = 0; j < 64; j++) hi[j] = 0.f;
    bool first  = true;
    int pending = 0;
    for (int c0 = 0; c0 < nit; c0 += CH, gc++)
    {
        const int b = gc & 1;
        mbar_wait(&accf[b], (gc >> 1) & 1);
        fence_after();
        const uint32_t taddr
            = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * ACC_COLS + half * 64);
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
        if (lane == 0) mbar_arrive(&acce[b]);
        if (++pending == FOLD || c0 + CH >= nit)
        {
            // the slot is L2-resident: 8 independent loads in flight, not one round trip per
            // element (16 would cost the registers that the per-chunk path needs)
#pragma unroll
            for (int j0 = 0; j0 < 64; j0 += 8)
            {
                double old[8];
#pragma unroll
                for (int j = 0; j < 8; j++)
                    old[j] = first ? 0. : __ldcg(dst + (size_t)(half * 64 + j0 + j) * BM + ml);
#pragma unroll
                for (int j = 0; j < 8; j++)
                    __stcg(dst + (size_t)(half * 64 + j0 + j) * BM + ml, old[j] + (double)hi[j0 + j]);
            }
#pragma unroll
            for (int j = 0; j < 64; j++) hi[j] = 0.f;
            first   = false;
            pending = 0;
        }
    }
}

// a whole off-diagonal tile: finish it in the main kernel (each thread re-reads the sums it wrote
// itself); everything else is left to the fix-up
template <bool SYRK>
__device__ __forceinline__ void write_whole_tile(const double* __restrict__ dst, double* __restrict__ Cb,
    int ldc, double alpha, double beta, int tile_m, int tile_n, int m, int n, int half, int ml)
{
    const int mm = tile_m * BM + ml;
    for (int j = 0; j < 64; j++)
    {
        const int nl = half * 64 + j;
        const int nn = tile_n * BN + nl;
        if (mm < m && nn < n)
        {
            const double old = (beta == 0.) ? 0. : beta * Cb[(size_t)nn * ldc + mm];
            const double val = alpha * __ldcg(dst + (size_t)nl * BM + ml) + old;
            Cb[(size_t)nn * ldc + mm] = val;
            if (SYRK) Cb[(size_t)mm * ldc + nn] = val;
        }
    }
}

// mapA / mapB: KO == 1: 3-D (k, row, slab) maps with a {32, 128, 1} box (any shape); KO > 1:
// 5-D (k in slab, row in group, slab, row group, slab-batch) maps with a {32, 8, KO, 16, 1} box
// (rows a multiple of 8, K a multiple of 32).  One k-iteration of the plan W is one box = KO
// slabs = one chunk of the tensor core's FP32 sums.
//
// Diagonal tiles of S = A^T A: with the truncating split A = H + L both for the TMEM and the
// shared-memory operand, P = H H^T + (2L) H^T needs two MMAs per k8 step instead of three and no
// low tile of B; the fix-up forms S = (P + P^T) / 2 = H H^T + L H^T + H L^T.
//
// Operand reuse is through L2 only: when the plan gives every tile the same number of CTAs (G =
// NT * ng: CTA g works on tile g / ng, K range g % ng; gemm_tn_t arranges that when it idles few
// CTAs), the NT CTAs of one K range start together and read the same operand boxes, each box
// being wanted by tm or tn of them.  They drift apart by more than what L2 holds, though: ncu
// shows 6.9 GB of DRAM reads for 4.3 GB of operands at 128^3 x 256, and the kernel then sits at
// the HBM roofline of that redundant traffic.  Keeping the CTAs of a K range in step with
// progress counters in global memory does bring the traffic down to 4.8 GB (measured), but any
// polling of global memory from inside this kernel -- by the producer lane or by a warp of its
// own -- waits behind the SM's outstanding TMA traffic and made the kernel 2-4x slower
// (gpurun_out/r02_umma_lock*.log); the remedy left is a cluster with remote mbarrier arrivals
// (or TMA multicast), not built.
// MC (KO == 2, no SYRK, tile grid with even tm and tn, G = NT * ng): clusters of four CTAs = the
// 2 x 2 tiles {m0, m1} x {n0, n1} over the same K range.  Every operand box is wanted by two of
// them: rank r loads one half (8 of the 16 row groups) of its A box and of its B box and
// multicasts each to the CTA that shares it (A: r ^ 2, B: r ^ 1), so a box crosses the L2 -> SM
// fabric -- and comes from DRAM -- once per cluster instead of once per tile.  mapA / mapB then
// carry {32, 8, KO, 8, 1} boxes; a stage is released by the three CTAs that write into it
// (multicast commit, barrier count 3).
template <bool SYRK, bool TRUNC, int KO, bool MC = false>
__global__ void __launch_bounds__(NTHR, 1) k_gemm_tn_umma(const __grid_constant__ CUtensorMap mapA,
    const __grid_constant__ CUtensorMap mapB, TnWork W, int m, int n, double alpha, double beta,
    double* __restrict__ C, int ldc, long long strideC, double* __restrict__ partial, int FOLD,
    int CHB)
{
    constexpr int NSS      = Ring<KO>::NSS;
    constexpr int NSLOT    = NSS * KO;        // slab slots of the raw ring
    constexpr uint32_t SBO = KO * 1024;       // bytes between the 8-row atoms of one slab's tile
    constexpr int OPER_B   = KO * TILE_B;     // one operand's box
    extern __shared__ uint8_t smraw[];
    __shared__ __align__(8) uint64_t full[NSS], empty[NSS], conv[NSLOT], cempty[CST], accf[2], acce[2];
    __shared__ uint32_t tmem_base_s;

    // the CTA's index in the plan; in a cluster launch rank r of cluster (quad q, K range j)
    // is tile (2 qm + (r & 1), 2 qn + (r >> 1)), i.e. plan index tile * ng + j
    int gcta       = blockIdx.x;
    const int rank = MC ? (int)(blockIdx.x & 3) : 0;
    if (MC)
    {
        const int ng = W.G / W.NT, cid = blockIdx.x >> 2;
        const int q = cid / ng, j = cid % ng, hm = W.tm / 2;
        gcta = ((2 * (q % hm) + (rank & 1)) + W.tm * (2 * (q / hm) + (rank >> 1))) * ng + j;
    }
    long long b0, b1;
    tn_cta_bounds(W, gcta, b0, b1);
    if (b1 <= b0) return; // never in a cluster launch
    const int u_first = tn_tile_of(W, b0), u_last = tn_tile_of(W, b1 - 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tiles must sit on 1024-byte boundaries (swizzle atom; descriptor base offset 0)
    uint8_t* ring  = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    uint8_t* lring = ring + (size_t)NSS * 2 * OPER_B; // CST tiles of B's low parts

    if (tid == 0)
    {
        for (int s = 0; s < NSS; s++)
        {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], MC ? 3 : 1);
        }
        for (int s = 0; s < NSLOT; s++) mbar_init(&conv[s], 4);
        for (int s = 0; s < CST; s++) mbar_init(&cempty[s], 1);
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
    // the partners' barriers must exist before anything is multicast into their stages
    if (MC) cluster_sync_all();
    fence_after();
    const uint32_t tmem   = tmem_base_s;
    const uint32_t tmem_a = tmem + 2 * ACC_COLS; // CST stages of 64 columns
    // CTAs that write into my stages and into whose stages I write: myself, the one sharing my
    // A box (rank ^ 2) and the one sharing my B box (rank ^ 1)
    const uint16_t mc_mask = (uint16_t)((1u << rank) | (1u << (rank ^ 1)) | (1u << (rank ^ 2)));

    // Every role walks the same list of segments (tile, boxes [it0, it1)); gb counts the boxes
    // of this CTA, gs = gb * KO + ko its slabs.
    if (warp == 0)
    {
        // ---------------- TMA producer ----------------
        if (lane == 0)
        {
            const uint64_t pol = policy_evict_normal();
            uint32_t gb        = 0;
            for (int u = u_first; u <= u_last; u++)
            {
                long long it0, it1;
                tn_seg(W, u, b0, b1, it0, it1);
                if (it0 >= it1) continue;
                int batch, tile_m, tile_n;
                tn_decode<SYRK>(W, u, batch, tile_m, tile_n);
                const bool diag = SYRK && u < W.ND;
                for (long long it = it0; it < it1; it++, gb++)
                {
                    const int ss = gb % NSS;
                    if (gb >= NSS) mbar_wait(&empty[ss], ((gb / NSS) & 1) ^ 1);
                    uint8_t* st = ring + (size_t)ss * 2 * OPER_B;
                    mbar_arrive_expect_tx(&full[ss], diag ? OPER_B : 2 * OPER_B);
                    if (KO == 1)
                    {
                        tma_load_3d(st, &mapA, &full[ss], (int)(it * 32), tile_m * BM, batch, pol);
                        if (!diag)
                            tma_load_3d(st + OPER_B, &mapB, &full[ss], (int)(it * 32), tile_n * BN, batch,
                                pol);
                    }
                    else if (!MC)
                    {
                        tma_load_5d(st, &mapA, &full[ss], 0, 0, (int)(it * KO), tile_m * (BM / 8), batch,
                            pol);
                        if (!diag)
                            tma_load_5d(st + OPER_B, &mapB, &full[ss], 0, 0, (int)(it * KO),
                                tile_n * (BN / 8), batch, pol);
                    }
                    else
                    {
                        // my half of the A box to me and to rank ^ 2, of the B box to me and rank ^ 1
                        const int ha = rank >> 1, hb = rank & 1;
                        tma_load_5d_mc(st + ha * (OPER_B / 2), &mapA, &full[ss], 0, 0, (int)(it * KO),
                            tile_m * (BM / 8) + 8 * ha, batch, (uint16_t)((1u << rank) | (1u << (rank ^ 2))),
                            pol);
                        tma_load_5d_mc(st + OPER_B + hb * (OPER_B / 2), &mapB, &full[ss], 0, 0,
                            (int)(it * KO), tile_n * (BN / 8) + 8 * hb, batch,
                            (uint16_t)((1u << rank) | (1u << (rank ^ 1))), pol);
                    }
                }
            }
        }
    }
    else if (warp == 1)
    {
        // ---------------- MMA issuer ----------------
        // The whole warp walks the loops converged and one elected lane issues: the operands of
        // tcgen05.mma live in uniform registers, and ptxas only keeps them there when it can see
        // that they are the same in every lane (issued from inside `if (lane == 0)` every MMA cost
        // an ELECT + 4 R2UR.BROADCAST waterfall, ~90 cycles of the issuing thread).
        const uint32_t tmem_u  = __shfl_sync(0xffffffffu, tmem, 0);
        const uint32_t ring_u  = __shfl_sync(0xffffffffu, smem_u32(ring), 0);
        const uint32_t lring_u = __shfl_sync(0xffffffffu, smem_u32(lring), 0);
        uint32_t gb = 0, gc = 0; // boxes, chunks (CHB boxes, never across segments)
        for (int u = u_first; u <= u_last; u++)
        {
            long long it0, it1;
            tn_seg(W, u, b0, b1, it0, it1);
            if (it0 >= it1) continue;
            const bool diag = SYRK && u < W.ND;
            for (long long it = it0; it < it1; it++, gb++)
            {
                const int cpos   = (int)((it - it0) % CHB);        // box of the chunk
                const bool clast = cpos == CHB - 1 || it == it1 - 1;
                const int ab = gc & 1, ss = gb % NSS;
                if (cpos == 0 && gc >= 2) mbar_wait(&acce[ab], ((gc >> 1) & 1) ^ 1);
                const uint32_t tacc = tmem_u + (uint32_t)(ab * ACC_COLS);
#pragma unroll
                for (int ko = 0; ko < KO; ko++)
                {
                    const uint32_t gs = gb * KO + ko;
                    const int cs = gs % CST, slot = ss * KO + ko;
                    mbar_wait(&conv[slot], (gb / NSS) & 1);
                    fence_after();
                    const uint32_t sa  = ring_u + (uint32_t)(ss * 2 * OPER_B + ko * 1024);
                    const uint64_t dbh = smem_desc(diag ? sa : sa + OPER_B, SBO);
                    const uint64_t dbl = smem_desc(lring_u + (uint32_t)(cs * TILE_B), 1024);
                    const uint32_t ah  = tmem_u + 2 * ACC_COLS + (uint32_t)(cs * 64);
                    if (elect_one())
                    {
#pragma unroll
                        for (int ks = 0; ks < 4; ks++)
                        {
                            const uint64_t adv = (uint64_t)(ks * 2); // 32 bytes along K
                            const uint32_t ak  = ah + (uint32_t)(ks * 8);
                            mma_tf32_ts(tacc, ak + 32, dbh + adv, (cpos > 0 || ko > 0 || ks > 0) ? 1u : 0u);
                            if (!diag) mma_tf32_ts(tacc, ak, dbl + adv, 1u);
                            mma_tf32_ts(tacc, ak, dbh + adv, 1u);
                        }
                        commit(&cempty[cs]);
                        if (ko == KO - 1)
                        {
                            if (MC)
                                commit_mc(&empty[ss], mc_mask);
                            else
                                commit(&empty[ss]);
                            if (clast) commit(&accf[ab]);
                        }
                    }
                    __syncwarp();
                }
                if (clast) gc++;
            }
        }
    }
    else if (warp < 6)
    {
        // ---------------- split warps ----------------
        const int ctid = tid - 64;
        const int row  = (warp & 3) * 32 + lane; // the TMEM lane quarter this warp may write
        uint32_t gb    = 0;
        for (int u = u_first; u <= u_last; u++)
        {
            long long it0, it1;
            tn_seg(W, u, b0, b1, it0, it1);
            if (it0 >= it1) continue;
            const bool diag = SYRK && u < W.ND;
            for (long long it = it0; it < it1; it++, gb++)
            {
                const int ss = gb % NSS;
                mbar_wait(&full[ss], (gb / NSS) & 1);
#pragma unroll
                for (int ko = 0; ko < KO; ko++)
                {
                    const uint32_t gs = gb * KO + ko;
                    const int cs = gs % CST, slot = ss * KO + ko;
                    if (gs >= CST) mbar_wait(&cempty[cs], ((gs / CST) & 1) ^ 1);
                    fence_after();
                    uint8_t* st          = ring + (size_t)ss * 2 * OPER_B + ko * 1024;
                    const uint32_t taddr = tmem_a + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cs * 64);
                    if (diag)
                        split_row_to_tmem<true, true>(st, SBO, row, taddr);
                    else
                    {
                        split_row_to_tmem<TRUNC, false>(st, SBO, row, taddr);
                        // B's low parts; !TRUNC also rounds B's raw tile in place
                        split_tile<TRUNC>(st + OPER_B, SBO, lring + (size_t)cs * TILE_B, ctid);
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    // generic-proxy writes -> visible to the tensor core (async proxy)
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&conv[slot]);
                }
            }
        }
    }
    else
    {
        // ---------------- drain warps ----------------
        const int dw   = warp - 6;
        const int q    = warp & 3;       // the TMEM lane quarter this warp may read
        const int half = dw >> 2;        // which 64 of the 128 columns
        const int ml   = q * 32 + lane;  // row of the tile = TMEM lane
        uint32_t gc    = 0;
        for (int u = u_first; u <= u_last; u++)
        {
            long long it0, it1;
            tn_seg(W, u, b0, b1, it0, it1);
            if (it0 >= it1) continue;
            int batch, tile_m, tile_n;
            tn_decode<SYRK>(W, u, batch, tile_m, tile_n);
            const bool diag = SYRK && u < W.ND;
            const int nit   = (int)(it1 - it0);
            double* dst = partial + ((size_t)gcta * W.smax + (size_t)(u - u_first)) * (BM * BN);
            drain_segment<SYRK>(dst, tmem, accf, acce, gc, nit, CHB, FOLD, q, half, ml, lane);
            if (!diag && it0 == 0 && it1 == W.nkt)
                write_whole_tile<SYRK>(dst, C + (long long)batch * strideC, ldc, alpha, beta, tile_m,
                    tile_n, m, n, half, ml);
        }
    }

    fence_before();
    __syncthreads();
    // nobody leaves while a partner may still multicast into this CTA or arrive on its barriers
    if (MC) cluster_sync_all();
    if (warp == 1)
    {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u)
                     : "memory");
    }
}

// K-major operand (rows = orbitals at base + r*ld, K contiguous, slabs `stride` apart).
// ko == 1: 3-D tensor (k, row, slab), box 32 x 128 x 1; rows and K beyond the extents are
// zero-filled.  ko > 1: 5-D tensor (k in slab, row in group of 8, slab, row group, slab-batch),
// box {32, 8, ko, 16, 1}: needs rows % 8 == 0 and K % 32 == 0 (whole groups / slabs only; slabs
// and groups beyond the extents are zero-filled).  Both under the 128-byte swizzle.
static int make_map_kmajor(CUtensorMap* mp, const float* base, size_t K, int rows, size_t ld,
    size_t stride, int nbatch, int ko, int groups = 16)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc)
    {
        set_error("cuTensorMapEncodeTiled entry point not found");
        return MGB_ECUDA;
    }
    const cuuint64_t bstride = (cuuint64_t)(nbatch > 1 ? stride : ld * (size_t)rows) * 4;
    CUresult r;
    if (ko == 1)
    {
        cuuint64_t dims[3]    = { (cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)nbatch };
        cuuint64_t strides[2] = { (cuuint64_t)ld * 4, bstride };
        cuuint32_t box[3]     = { 32, 128, 1 };
        cuuint32_t estr[3]    = { 1, 1, 1 };
        r = enc(mp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box,
            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    else
    {
        cuuint64_t dims[5]    = { 32, 8, (cuuint64_t)(K / 32), (cuuint64_t)(rows / 8), (cuuint64_t)nbatch };
        cuuint64_t strides[4] = { (cuuint64_t)ld * 4, 128, (cuuint64_t)ld * 4 * 8, bstride };
        cuuint32_t box[5]     = { 32, 8, (cuuint32_t)ko, (cuuint32_t)groups, 1 };
        cuuint32_t estr[5]    = { 1, 1, 1, 1, 1 };
        r = enc(mp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box,
            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS)
    {
        set_error("cuTensorMapEncodeTiled failed (%d): K-major operand K %zu rows %d ld %zu ko %d",
            (int)r, K, rows, ld, ko);
        return MGB_ECUDA;
    }
    return MGB_OK;
}

} // namespace umma
