// ORBDTYPE float contractions C = alpha A^T B + beta C on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM), included by contractions.cu after the
// stream-K plan (TnWork, tn_*): same (tile, K) decomposition, same partial-slot format and the
// same fix-up kernel as the DMMA / mma.sync kernels -- only the inner loop differs.
//
// Arithmetic (src/linear_algebra/mputils.cc:848-948 sums (double)a*(double)b): error-compensated
// 3xTF32.  Every operand is split a = a_hi + a_lo into two TF32 numbers and a*b is formed as
// a_lo b_hi + a_hi b_lo + a_hi b_hi (the dropped a_lo b_lo is 2^-22 of |a||b|).  The tensor
// core sums in FP32 over one chunk of CH 32-point slabs only; chunk sums are added in FP32
// registers (round to nearest) over at most FOLD chunks and those in double (the segment's
// partial slot, L2-resident), so the K = 10^6..10^7 reduction behaves like a double sum.
//
// One CTA per SM, 14 warps, each with one role:
//   warp 0      TMA producer: the raw 128 x 32 float tiles of A and B (K-major, 128-byte swizzle
//               = the canonical UMMA K-major SWIZZLE_128B layout) into a 5-stage ring (the ring
//               depth hides the load latency; 160 KB in flight per SM)
//   warps 2-5   split: read a raw tile, write the low part to a twin tile (2-stage ring: it
//               only lives from the split to the end of the slab's MMAs) at the SAME
//               offsets (the swizzle never has to be undone); the high part is either the raw
//               tile itself (TRUNC: the tensor core ignores the 13 low mantissa bits, so
//               a_hi = trunc(a) costs nothing and a_lo = a - trunc(a) is exact) or is rounded in
//               place (cvt.rna)
//   warp 1      one lane issues the 12 tcgen05.mma of a slab (4 k8 steps x 3 products) and
//               commits the stage back to the producer; owns the TMEM allocation
//   warps 6-13  drain: tcgen05.ld of a finished chunk accumulator (two 128-column accumulators
//               alternate, so the drain overlaps the next chunk's MMAs), the two upper summation
//               levels, the partial slot / direct write of C
#pragma once

namespace umma
{
constexpr int UST      = 5;           // ring depth of the raw tiles
constexpr int LST      = 2;           // ring depth of the low-part tiles
constexpr int TILE_B   = 128 * 128;   // bytes of one 128-row x 32-float tile
constexpr int STAGE_B  = 2 * TILE_B;  // A, B (raw stage) / A low, B low (low stage)
constexpr int NTHR     = 448;         // 14 warps
constexpr int ACC_COLS = 128;         // TMEM columns of one accumulator
constexpr size_t SMEM  = (size_t)(UST + LST) * STAGE_B + 1024;

// UMMA::SmemDescriptor of a K-major SWIZZLE_128B tile (8-row x 128-byte atoms, 1024 B apart):
// start address [0,14) >> 4, leading byte offset [16,30) = 1, stride byte offset [32,46) =
// 1024 >> 4, version [46,48) = 1, layout type [61,64) = 2.  Pinned on the device by
// tools/microbench/umma_tf32_probe.cu.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// UMMA::InstrDescriptor: D = F32 [4,6), A = B = TF32 [7,10) [10,13), both K-major, N >> 3 at
// [17,23), M >> 4 at [24,29): M = N = 128
constexpr uint32_t IDESC
    = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t acc)
{
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
                 "}\n" ::"r"(tmem_c),
                 "l"(da), "l"(db), "r"(IDESC), "r"(acc)
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
template <bool TRUNC>
__device__ __forceinline__ void split_tile(uint8_t* raw, uint8_t* lo, int ctid)
{
#pragma unroll
    for (int j = 0; j < TILE_B / 16 / 128; j++)
    {
        const int off = (ctid + 128 * j) * 16;
        float4 v      = *reinterpret_cast<const float4*>(raw + off);
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
            *reinterpret_cast<float4*>(raw + off) = h;
        }
        *reinterpret_cast<float4*>(lo + off) = l;
    }
}

// Which 32-point slab a CTA loads at position p of tile u's K range.  The stream-K plan hands a
// CTA a CONTIGUOUS range of positions; read as slabs, the P CTAs that share a tile would each walk
// their own far-apart region of K, so every 128-byte piece a TMA box fetches (one per orbital, 8 MB
// apart) opens its own DRAM page (measured: 3.2 TB/s of DRAM reads at 128^3 x 256 whatever the
// kernel did on chip).  Positions are therefore mapped to slabs round-robin: position p lies in
// row r = p / (nkt / P) of a P-row table and is the t-th entry of that row, slab = t * P + r --
// at any moment the P CTAs of a tile (and those of the tiles that share its operands) read P
// neighbouring slabs, i.e. a few KB of every orbital.  A bijection of [0, nkt); only the TMA
// producer needs it.  The summation order stays fixed (deterministic), just not ascending in K.
struct SlabPerm
{
    int P, q, rem, split;
    __device__ __forceinline__ SlabPerm(const TnWork& W, int u, int on)
    {
        const int c = u < W.ND ? W.cd : W.cf;
        long long pp = (long long)W.G * c * W.nkt / W.tot;
        P = on ? (int)(pp < 1 ? 1 : pp) : 1;
        if ((long long)P > W.nkt) P = (int)W.nkt;
        q     = (int)(W.nkt / P);
        rem   = (int)(W.nkt % P);
        split = rem * (q + 1);
    }
    __device__ __forceinline__ int operator()(int p) const
    {
        if (P == 1) return p;
        int r, t;
        if (p < split)
        {
            r = p / (q + 1);
            t = p - r * (q + 1);
        }
        else
        {
            const int o = p - split;
            r = o / q;
            t = o - r * q;
            r += rem;
        }
        return t * P + r;
    }
};

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
            // the slot is L2-resident: 16 independent loads in flight, not one round trip
            // per element
#pragma unroll
            for (int j0 = 0; j0 < 64; j0 += 16)
            {
                double old[16];
#pragma unroll
                for (int j = 0; j < 16; j++)
                    old[j] = first ? 0. : __ldcg(dst + (size_t)(half * 64 + j0 + j) * BM + ml);
#pragma unroll
                for (int j = 0; j < 16; j++)
                {
                    __stcg(dst + (size_t)(half * 64 + j0 + j) * BM + ml, old[j] + (double)hi[j0 + j]);
                    hi[j0 + j] = 0.f;
                }
            }
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

template <bool SYRK, bool TRUNC>
__global__ void __launch_bounds__(NTHR, 1) k_gemm_tn_umma(const __grid_constant__ CUtensorMap mapA,
    const __grid_constant__ CUtensorMap mapB, TnWork W, int m, int n, double alpha, double beta,
    double* __restrict__ C, int ldc, long long strideC, double* __restrict__ partial, int CH,
    int FOLD, int PERM)
{
    extern __shared__ uint8_t smraw[];
    __shared__ __align__(8) uint64_t full[UST], conv[UST], empty[UST], lempty[LST], accf[2], acce[2];
    __shared__ uint32_t tmem_base_s;

    const int gcta = blockIdx.x;
    long long b0, b1;
    tn_cta_bounds(W, gcta, b0, b1);
    if (b1 <= b0) return;
    const int u_first = tn_tile_of(W, b0), u_last = tn_tile_of(W, b1 - 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tiles must sit on 1024-byte boundaries (swizzle atom; descriptor base offset 0)
    uint8_t* ring  = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    uint8_t* lring = ring + (size_t)UST * STAGE_B;

    if (tid == 0)
    {
        for (int s = 0; s < UST; s++)
        {
            mbar_init(&full[s], 1);
            mbar_init(&conv[s], 4);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < LST; s++) mbar_init(&lempty[s], 1);
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
                     "r"((uint32_t)(2 * ACC_COLS))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0)
    {
        // ---------------- TMA producer ----------------
        if (lane == 0)
        {
            const uint64_t pol = policy_evict_normal();
            uint32_t gs        = 0;
            for (int u = u_first; u <= u_last; u++)
            {
                long long it0, it1;
                tn_seg(W, u, b0, b1, it0, it1);
                if (it0 >= it1) continue;
                int batch, tile_m, tile_n;
                tn_decode<SYRK>(W, u, batch, tile_m, tile_n);
                const bool diag = SYRK && u < W.ND;
                const SlabPerm perm(W, u, PERM);
                for (long long it = it0; it < it1; it++, gs++)
                {
                    const int s       = gs % UST;
                    const uint32_t ph = (gs / UST) & 1;
                    if (gs >= UST) mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = ring + (size_t)s * STAGE_B;
                    const int k0 = perm((int)it) * 32;
                    mbar_arrive_expect_tx(&full[s], diag ? TILE_B : 2 * TILE_B);
                    tma_load_3d(st, &mapA, &full[s], k0, tile_m * BM, batch, pol);
                    if (!diag) tma_load_3d(st + TILE_B, &mapB, &full[s], k0, tile_n * BN, batch, pol);
                }
            }
        }
    }
    else if (warp == 1)
    {
        // ---------------- MMA issuer ----------------
        if (lane == 0)
        {
            uint32_t gs = 0, gc = 0;
            for (int u = u_first; u <= u_last; u++)
            {
                long long it0, it1;
                tn_seg(W, u, b0, b1, it0, it1);
                if (it0 >= it1) continue;
                const bool diag = SYRK && u < W.ND;
                const int nit   = (int)(it1 - it0);
                for (int c0 = 0; c0 < nit; c0 += CH, gc++)
                {
                    const int b = gc & 1;
                    if (gc >= 2) mbar_wait(&acce[b], ((gc >> 1) & 1) ^ 1);
                    fence_after();
                    const uint32_t tacc = tmem + (uint32_t)(b * ACC_COLS);
                    const int c1        = c0 + CH < nit ? c0 + CH : nit;
                    for (int kt = c0; kt < c1; kt++, gs++)
                    {
                        const int s = gs % UST;
                        mbar_wait(&conv[s], (gs / UST) & 1);
                        fence_after();
                        const int ls      = gs % LST;
                        const uint32_t sa = smem_u32(ring + (size_t)s * STAGE_B);
                        const uint32_t sl = smem_u32(lring + (size_t)ls * STAGE_B);
                        const uint64_t dah = smem_desc(sa), dal = smem_desc(sl);
                        const uint64_t dbh = diag ? dah : smem_desc(sa + TILE_B);
                        const uint64_t dbl = diag ? dal : smem_desc(sl + TILE_B);
#pragma unroll
                        for (int ks = 0; ks < 4; ks++)
                        {
                            // 32 bytes further along K inside the swizzle row
                            const uint64_t adv = (uint64_t)(ks * 2);
                            mma_tf32(tacc, dal + adv, dbh + adv, (kt > c0 || ks > 0) ? 1u : 0u);
                            mma_tf32(tacc, dah + adv, dbl + adv, 1u);
                            mma_tf32(tacc, dah + adv, dbh + adv, 1u);
                        }
                        commit(&empty[s]);
                        commit(&lempty[ls]);
                    }
                    commit(&accf[b]);
                }
            }
        }
    }
    else if (warp < 6)
    {
        // ---------------- split warps ----------------
        const int ctid = tid - 64;
        uint32_t gs    = 0;
        for (int u = u_first; u <= u_last; u++)
        {
            long long it0, it1;
            tn_seg(W, u, b0, b1, it0, it1);
            if (it0 >= it1) continue;
            const bool diag = SYRK && u < W.ND;
            for (long long it = it0; it < it1; it++, gs++)
            {
                const int s = gs % UST, ls = gs % LST;
                mbar_wait(&full[s], (gs / UST) & 1);
                if (gs >= LST) mbar_wait(&lempty[ls], ((gs / LST) & 1) ^ 1);
                uint8_t* st = ring + (size_t)s * STAGE_B;
                uint8_t* lo = lring + (size_t)ls * STAGE_B;
                split_tile<TRUNC>(st, lo, ctid);
                if (!diag) split_tile<TRUNC>(st + TILE_B, lo + TILE_B, ctid);
                // generic-proxy writes -> visible to the tensor core (async proxy)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&conv[s]);
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
            drain_segment<SYRK>(dst, tmem, accf, acce, gc, nit, CH, FOLD, q, half, ml, lane);
            if (!diag && it0 == 0 && it1 == W.nkt)
                write_whole_tile<SYRK>(dst, C + (long long)batch * strideC, ldc, alpha, beta, tile_m,
                    tile_n, m, n, half, ml);
        }
    }

    fence_before();
    __syncthreads();
    if (warp == 1)
    {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                     "r"((uint32_t)(2 * ACC_COLS))
                     : "memory");
    }
}

// ---------------------------------------------------------------------------
// "TS" variant: the A operand of every tcgen05.mma comes from TMEM instead of shared memory.
// With both operands in shared memory a 128 x 128 x 8 TF32 MMA reads 8 KB per 64 cycles -- the
// whole shared-memory bandwidth of the SM -- and the split warps' and TMA's traffic come on top
// (measured: the SS kernel above runs at ~1/3 of the tensor rate, and slower again when the
// split also rewrites the high parts).  Here the split warps read A's raw tile row by row (one
// thread per row = TMEM lane, de-swizzled 16-byte chunks, conflict-free) and store a_hi and a_lo
// with tcgen05.st into a 4-stage ring of TMEM columns (2 x 32 columns per stage, next to the two
// 128-column accumulators: all 512 columns are used); only B's tiles are read from shared memory
// by the tensor core.  Rings: 5 raw stages (A, B; 32 KB each) and 4 split stages (B low, 16 KB +
// the TMEM columns of A).
// ---------------------------------------------------------------------------
constexpr int CST = 4; // split stages: B low in shared memory, A high / low in TMEM

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

// row `row` of a raw K-major SWIZZLE_128B tile -> a_hi, a_lo in TMEM lane `row`, columns
// [col, col+32) and [col+32, col+64).  TRUNC: a_hi = a with the 13 low mantissa bits cleared (what
// the tensor core would read anyway), a_lo = a - a_hi exactly: two instructions per element -- the
// split warps are one warp per SM sub-partition and their instruction count bounds the kernel
// (cvt.rna.tf32 is a four-instruction sequence)
template <bool TRUNC>
__device__ __forceinline__ void split_row_to_tmem(const uint8_t* raw, int row, uint32_t taddr)
{
    const uint8_t* rb = raw + (row >> 3) * 1024 + (row & 7) * 128;
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
                h[4 * c + e] = __float_as_uint(x[e]) & 0xFFFFE000u;
                l[4 * c + e] = __float_as_uint(x[e] - __uint_as_float(h[4 * c + e]));
            }
            else
            {
                const float hh = rna_tf32(x[e]);
                h[4 * c + e]   = __float_as_uint(hh);
                l[4 * c + e]   = __float_as_uint(rna_tf32(x[e] - hh));
            }
        }
    }
    st32(taddr, h);
    st32(taddr + 32, l);
}

template <bool SYRK, bool TRUNC>
__global__ void __launch_bounds__(NTHR, 1) k_gemm_tn_umma_ts(const __grid_constant__ CUtensorMap mapA,
    const __grid_constant__ CUtensorMap mapB, TnWork W, int m, int n, double alpha, double beta,
    double* __restrict__ C, int ldc, long long strideC, double* __restrict__ partial, int CH,
    int FOLD, int PERM, int DBG)
{
    extern __shared__ uint8_t smraw[];
    __shared__ __align__(8) uint64_t full[UST], conv[UST], empty[UST], cempty[CST], accf[2], acce[2];
    __shared__ uint32_t tmem_base_s;

    const int gcta = blockIdx.x;
    long long b0, b1;
    tn_cta_bounds(W, gcta, b0, b1);
    if (b1 <= b0) return;
    const int u_first = tn_tile_of(W, b0), u_last = tn_tile_of(W, b1 - 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* ring  = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    uint8_t* lring = ring + (size_t)UST * STAGE_B; // CST tiles of B's low parts

    if (tid == 0)
    {
        for (int s = 0; s < UST; s++)
        {
            mbar_init(&full[s], 1);
            mbar_init(&conv[s], 4);
            mbar_init(&empty[s], 1);
        }
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
    fence_after();
    const uint32_t tmem   = tmem_base_s;
    const uint32_t tmem_a = tmem + 2 * ACC_COLS; // CST stages of 64 columns

    if (warp == 0)
    {
        // ---------------- TMA producer ----------------
        if (lane == 0)
        {
            const uint64_t pol = policy_evict_normal();
            uint32_t gs        = 0;
            for (int u = u_first; u <= u_last; u++)
            {
                long long it0, it1;
                tn_seg(W, u, b0, b1, it0, it1);
                if (it0 >= it1) continue;
                int batch, tile_m, tile_n;
                tn_decode<SYRK>(W, u, batch, tile_m, tile_n);
                const bool diag = SYRK && u < W.ND;
                const SlabPerm perm(W, u, PERM);
                for (long long it = it0; it < it1; it++, gs++)
                {
                    const int s       = gs % UST;
                    const uint32_t ph = (gs / UST) & 1;
                    if (gs >= UST) mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = ring + (size_t)s * STAGE_B;
                    // DBG (timing experiments only, results are wrong): 1 = keep K inside an
                    // L2-resident window, 2 = no MMAs, 4 = no split work
                    const int k0 = ((DBG & 1) ? (perm((int)it) & 255) : perm((int)it)) * 32;
                    mbar_arrive_expect_tx(&full[s], diag ? TILE_B : 2 * TILE_B);
                    tma_load_3d(st, &mapA, &full[s], k0, tile_m * BM, batch, pol);
                    if (!diag) tma_load_3d(st + TILE_B, &mapB, &full[s], k0, tile_n * BN, batch, pol);
                }
            }
        }
    }
    else if (warp == 1)
    {
        // ---------------- MMA issuer ----------------
        if (lane == 0)
        {
            uint32_t gs = 0, gc = 0;
            for (int u = u_first; u <= u_last; u++)
            {
                long long it0, it1;
                tn_seg(W, u, b0, b1, it0, it1);
                if (it0 >= it1) continue;
                const bool diag = SYRK && u < W.ND;
                const int nit   = (int)(it1 - it0);
                for (int c0 = 0; c0 < nit; c0 += CH, gc++)
                {
                    const int b = gc & 1;
                    if (gc >= 2) mbar_wait(&acce[b], ((gc >> 1) & 1) ^ 1);
                    fence_after();
                    const uint32_t tacc = tmem + (uint32_t)(b * ACC_COLS);
                    const int c1        = c0 + CH < nit ? c0 + CH : nit;
                    for (int kt = c0; kt < c1; kt++, gs++)
                    {
                        const int s = gs % UST, cs = gs % CST;
                        mbar_wait(&conv[s], (gs / UST) & 1);
                        fence_after();
                        const uint32_t sa  = smem_u32(ring + (size_t)s * STAGE_B);
                        const uint64_t dbh = smem_desc(diag ? sa : sa + TILE_B);
                        const uint64_t dbl = smem_desc(smem_u32(lring + (size_t)cs * TILE_B));
                        const uint32_t ah  = tmem_a + (uint32_t)(cs * 64);
                        if (!(DBG & 2))
#pragma unroll
                        for (int ks = 0; ks < 4; ks++)
                        {
                            const uint64_t adv = (uint64_t)(ks * 2); // 32 bytes along K
                            const uint32_t ak  = ah + (uint32_t)(ks * 8);
                            mma_tf32_ts(tacc, ak + 32, dbh + adv, (kt > c0 || ks > 0) ? 1u : 0u);
                            mma_tf32_ts(tacc, ak, dbl + adv, 1u);
                            mma_tf32_ts(tacc, ak, dbh + adv, 1u);
                        }
                        commit(&empty[s]);
                        commit(&cempty[cs]);
                    }
                    commit(&accf[b]);
                }
            }
        }
    }
    else if (warp < 6)
    {
        // ---------------- split warps ----------------
        const int ctid = tid - 64;
        const int row  = (warp & 3) * 32 + lane; // the TMEM lane quarter this warp may write
        uint32_t gs    = 0;
        for (int u = u_first; u <= u_last; u++)
        {
            long long it0, it1;
            tn_seg(W, u, b0, b1, it0, it1);
            if (it0 >= it1) continue;
            const bool diag = SYRK && u < W.ND;
            for (long long it = it0; it < it1; it++, gs++)
            {
                const int s = gs % UST, cs = gs % CST;
                mbar_wait(&full[s], (gs / UST) & 1);
                if (gs >= CST) mbar_wait(&cempty[cs], ((gs / CST) & 1) ^ 1);
                fence_after();
                uint8_t* st = ring + (size_t)s * STAGE_B;
                if (!(DBG & 4))
                {
                split_row_to_tmem<TRUNC>(st, row, tmem_a + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cs * 64));
                // B's low parts; !TRUNC rounds B's raw tile in place -- except on a diagonal
                // tile, where B is A's raw tile, which other threads are still reading by rows
                if (diag)
                    split_tile<true>(st, lring + (size_t)cs * TILE_B, ctid);
                else
                    split_tile<TRUNC>(st + TILE_B, lring + (size_t)cs * TILE_B, ctid);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&conv[s]);
            }
        }
    }
    else
    {
        // ---------------- drain warps ----------------
        const int dw   = warp - 6;
        const int q    = warp & 3;
        const int half = dw >> 2;
        const int ml   = q * 32 + lane;
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
            drain_segment<SYRK>(dst, tmem, accf, acce, gc, nit, CH, FOLD, q, half, ml, lane);
            if (!diag && it0 == 0 && it1 == W.nkt)
                write_whole_tile<SYRK>(dst, C + (long long)batch * strideC, ldc, alpha, beta, tile_m,
                    tile_n, m, n, half, ml);
        }
    }

    fence_before();
    __syncthreads();
    if (warp == 1)
    {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u)
                     : "memory");
    }
}

// K-major operand (rows = orbitals at base + r*ld, K contiguous, slabs `stride` apart) as a
// 3-D tensor (k, row, slab) with a 32 x 128 x 1 box under the 128-byte swizzle; rows and K
// beyond the extents are zero-filled
static int make_map_kmajor(CUtensorMap* mp, const float* base, size_t K, int rows, size_t ld,
    size_t stride, int nbatch)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc)
    {
        set_error("cuTensorMapEncodeTiled entry point not found");
        return MGB_ECUDA;
    }
    cuuint64_t dims[3]    = { (cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)nbatch };
    cuuint64_t strides[2] = { (cuuint64_t)ld * 4,
        (cuuint64_t)(nbatch > 1 ? stride : ld * (size_t)rows) * 4 };
    cuuint32_t box[3]  = { 32, 128, 1 };
    cuuint32_t estr[3] = { 1, 1, 1 };
    CUresult r = enc(mp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims,
        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
    {
        set_error("cuTensorMapEncodeTiled failed (%d): K-major operand K %zu rows %d ld %zu",
            (int)r, K, rows, ld);
        return MGB_ECUDA;
    }
    return MGB_OK;
}

} // namespace umma
