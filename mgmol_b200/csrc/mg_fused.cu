// Fused multigrid kernels for the float block preconditioner
// (Preconditioning<float>::mg, src/Preconditioning.cc:155-216) on NO-GHOST
// float blocks [function][x][y][z]:
//
//   k_mg_jacobi    one damped-Jacobi sweep  v' = v - omega (A v - f)  in ONE pass
//                  (GridFuncVector::jacobi, src/pb/GridFuncVector.cc:2416-2425:
//                  trade_boundaries + applyLap + "-=" + axpy = 7 sweeps over
//                  the block in the reference, 3 here), optionally emitting
//                  the residual w = A v - f that the restriction consumes.
//   k_mg_restrict  full-weighting restriction (MGkernelRestrict3D,
//                  src/pb/MGkernels.cc:203-283)
//   k_mg_prolong_correct   v -= P e  (MGkernelExtend3D, src/pb/MGkernels.cc:
//                  28-201, followed by the "v -= w" of Preconditioning.cc:206)
//   k_mg_convert / k_mg_scale   entry conversions.
//
// Boundaries.  The reference keeps ghost layers and refreshes them with
// trade_boundaries before every stencil; here the periodic wrap is folded into
// the tile addressing (TMA box coordinates in y, plane index in x, column
// index in z) and Dirichlet-0 ghosts are the TMA out-of-bounds zero fill.  The
// reference's Dirichlet trade also zeroes the first interior layer on the low
// side (src/pb/GridFunc.cc:2192-2336); every array that the reference trades
// before its next use is therefore STORED with that layer already zero
// (zero_low), which is equivalent and costs nothing.
//
// Arithmetic.  Blocks are float (MGPRECONDTYPE).  The reference accumulates
// each stencil in double and rounds once to float.  The conversion pipe of
// sm_100a sustains only 16 F2F/clk/SM (tools/microbench/pipes.cu), too few for
// an HBM-bound float sweep, so the stencil is evaluated in float in DIFFERENCE
// form, A v = sum_k c_k (v_k - v_0) (the coefficients of all Laplacians here
// sum to zero): neighbouring values subtract almost exactly, so the rounding
// error is relative to the size of the second differences, not to c_0 |v|, and
// the result agrees with the double-accumulated reference to a few float ulps
// of (A v).  Products with double constants (gamma, -omega) are formed with a
// two-float split of the constant and an FMA error term, i.e. correctly
// rounded except in ~2^-24 of the cases.  Prolongation follows the reference's
// float operand order exactly.  Parity tests use a 1e-6 relative tolerance on
// the V-cycle output (well inside the 1e-5 FP32 bar).
#include <cuda.h>

#include <cstring>

#include "hpsi.h"
#include "masks.h"
#include "mg_fused.h"
#include "tma_ptx.cuh"

namespace mgb
{

enum
{
    kLapMehr = 0,
    kLap2    = 1,
    kLap4    = 2
};

struct alignas(64) JacobiParams
{
    CUtensorMap in_mid, in_halo; // source block (z, y, x, function), float
    CUtensorMap pw_mid, pw_halo; // west / east neighbour's source block (peer
    CUtensorMap pe_mid, pe_halo; // memory over NVLink) on an x-split domain
    int split_x, have_w, have_e;
    // P3: any px x py x pz decomposition, every halo read in place.  nbmaps[dx*36 + dy*12 +
    // dz*4 + kind] (dx, dy, dz in 0..2 = -1, 0, +1): tensor maps over the source block of the
    // rank at coord + (dx,dy,dz) (myself across a direction that is not split); kind 0: box
    // {nz, TY}, 1: {nz, G}, 2: {4, TY}, 3: {4, G} (16 bytes of z: the z-halo columns)
    const CUtensorMap* nbmaps;
    int zsplit;                  // z-halo columns come from boxes
    int end_lo[3], end_hi[3];    // non-periodic direction and this rank holds the domain's end
    int zlo_off, zarr_mid, zarr_hi, zarr_bytes; // z-halo arrays inside a tile
    const int* map_w;            // color slot of my color's orbital on that rank
    const int* map_e;            // (-1: absent), or null = same slot
    const float* f;              // right-hand side
    float* vout;                 // float output (or null)
    double* vout64;              // double output (exit of precond_mg, or null)
    float* wout;                 // residual A v - f (or null)
    long long ld_f, ld_out, ld_w;
    int nx, ny, nz, nfunc;
    int TY, YG, NB, S, XC;
    int nzv, tpo;
    int row_bytes, off_mid, off_hi, tile_bytes, stage_bytes;
    int per[3];      // periodic direction
    int zero_low[3]; // store zeros on the low layer of this direction
    float sh, sl;    // two-float split of the input scaling (SCALE variant)
    float oh, ol;    // two-float split of -omega
    float c[8];      // stencil coefficients
    MaskView mask;   // app_mask(level) applied to the stored v' and w (MASK)
};

constexpr int kMgBarBytes = 1024;

// RN_float((ch + cl) * x) up to a 2^-24 chance of a one-ulp difference
__device__ __forceinline__ float mul_split(float ch, float cl, float x)
{
    const float p = __fmul_rn(ch, x);
    const float e = __fmaf_rn(ch, x, -p);
    return __fadd_rn(p, __fmaf_rn(cl, x, e));
}

__device__ __forceinline__ void lds4(uint32_t a, float (&o)[4])
{
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3])
                 : "r"(a));
}
__device__ __forceinline__ void lds2(uint32_t a, float& x, float& y)
{
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(a));
}
__device__ __forceinline__ float lds1(uint32_t a)
{
    float x;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a));
    return x;
}

// second difference (a - c) + (b - c)
__device__ __forceinline__ float dd(float a, float b, float c)
{
    return __fadd_rn(__fsub_rn(a, c), __fsub_rn(b, c));
}

// One Jacobi sweep.  grid.x: y tile, grid.y: block of NB functions, grid.z: x
// chunk.  Warp 0 is the TMA producer; the consumers keep the 2G+1 planes a
// stencil needs resident in a ring of S stages and read every tap from shared
// memory.
template <int LAP, int RY, bool SCALE, int MAXT, bool MASK, bool P3>
__global__ void __launch_bounds__(MAXT, 1) k_mg_jacobi(const __grid_constant__ JacobiParams P)
{
    constexpr int G  = (LAP == kLap4) ? 2 : 1;
    constexpr int NW = RY + 2 * G; // rows walked by one thread

    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full  = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 16;

    const int tid     = threadIdx.x;
    const int y0      = blockIdx.x * P.TY;
    const int orb0    = blockIdx.y * P.NB;
    const int xb      = blockIdx.z * P.XC;
    const int xe      = min(xb + P.XC, P.nx);
    const int nplanes = (xe - xb) + 2 * G;
    const int S       = P.S;

    if (tid == 0)
    {
        const int ncw = (blockDim.x - 32) >> 5;
        for (int s = 0; s < S; s++)
        {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], ncw);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();

    if (tid < 32)
    {
        // ---------------------------- TMA producer ----------------------------
        // (P3: every lane issues, one (function, box) each)
        if (!P3 && tid != 0) return;
        unsigned char* stages = smem + kMgBarBytes;
        const uint64_t pol    = policy_evict_first();
        int norb              = P.nfunc - orb0;
        if (norb > P.NB) norb = P.NB;
        int ylo = y0 - G, yhi = y0 + P.TY;
        if (P.per[1])
        {
            if (ylo < 0) ylo += P.ny;
            if (yhi >= P.ny) yhi -= P.ny;
        }
        uint32_t tx = (uint32_t)norb * (uint32_t)((P.TY + 2 * G) * P.row_bytes);
        // only the 19-point operator looks at the z neighbours of halo rows
        constexpr bool ZALL = (LAP == kLapMehr);
        if (P3 && P.zsplit) tx += (uint32_t)norb * (uint32_t)((ZALL ? P.TY + 2 * G : P.TY) * 32);
        int stage = 0;
        uint32_t par = 0;
        for (int it = 0; it < nplanes; it++)
        {
            int xc = xb - G + it;
            if constexpr (P3)
            {
                // the rank holding plane xc and the coordinate there; beyond a Dirichlet
                // end the box stays out of range on my own block (zero fill)
                const int dx  = (xc < 0 && !P.end_lo[0]) ? 0 : ((xc >= P.nx && !P.end_hi[0]) ? 2 : 1);
                const int xcp = dx == 0 ? xc + P.nx : (dx == 2 ? xc - P.nx : xc);
                mbar_wait(&empty[stage], par ^ 1u);
                if (tid == 0) mbar_arrive_expect_tx(&full[stage], tx);
                unsigned char* sb = stages + (size_t)stage * P.stage_bytes;
                // one lane per (function, box): rows below the tile, the tile's rows, rows
                // above, each with the 32-byte z-halo entries of its rows
                for (int l = tid; l < 3 * norb; l += 32)
                {
                    const int o = l / 3, b = l - 3 * o;
                    unsigned char* tb = sb + (size_t)o * P.tile_bytes;
                    int dyi = 1, yc = y0, to = P.off_mid, zo = P.zarr_mid;
                    if (b == 0)
                    {
                        const bool nbr = y0 - G < 0 && !P.end_lo[1];
                        dyi = nbr ? 0 : 1;
                        yc  = nbr ? P.ny - G : y0 - G;
                        to = 0, zo = 0;
                    }
                    else if (b == 2)
                    {
                        const bool nbr = y0 + P.TY >= P.ny && !P.end_hi[1];
                        dyi = nbr ? 2 : 1;
                        yc  = nbr ? 0 : y0 + P.TY;
                        to = P.off_hi, zo = P.zarr_hi;
                    }
                    const CUtensorMap* m = P.nbmaps + dx * 36 + dyi * 12;
                    const int kind       = (b == 1) ? 0 : 1;
                    tma_load_4d(tb + to, m + 4 + kind, &full[stage], 0, yc, xcp, orb0 + o, pol);
                    if (P.zsplit && (ZALL || b == 1))
                        tma_load_4d(tb + P.zlo_off + zo, m + 2 + kind, &full[stage], 0, yc, xcp,
                            orb0 + o, pol);
                }
                if (++stage == S)
                {
                    stage = 0;
                    par ^= 1u;
                }
                continue;
            }
            const CUtensorMap* mm = &P.in_mid;
            const CUtensorMap* mh = &P.in_halo;
            const int* cmap       = nullptr;
            if (P.split_x)
            {
                // planes outside my box: the neighbour's block, read in place
                if (xc < 0 && P.have_w)
                {
                    mm   = &P.pw_mid;
                    mh   = &P.pw_halo;
                    cmap = P.map_w;
                    xc += P.nx;
                }
                else if (xc >= P.nx && P.have_e)
                {
                    mm   = &P.pe_mid;
                    mh   = &P.pe_halo;
                    cmap = P.map_e;
                    xc -= P.nx;
                }
            }
            else if (P.per[0])
            {
                xc %= P.nx;
                if (xc < 0) xc += P.nx;
            }
            // Dirichlet: xc outside [0, nx) -> the whole box is out of bounds
            // and arrives zero-filled
            mbar_wait(&empty[stage], par ^ 1u);
            mbar_arrive_expect_tx(&full[stage], tx);
            unsigned char* sb = stages + (size_t)stage * P.stage_bytes;
            for (int o = 0; o < norb; o++)
            {
                unsigned char* tb = sb + (size_t)o * P.tile_bytes;
                int fo = orb0 + o;
                if (cmap)
                {
                    // gid-addressed: the neighbour may hold this orbital in
                    // another color slot, or not at all (-> zero fill)
                    fo = cmap[fo];
                    if (fo < 0) fo = P.nfunc;
                }
                tma_load_4d(tb + P.off_mid, mm, &full[stage], 0, y0, xc, fo, pol);
                tma_load_4d(tb, mh, &full[stage], 0, ylo, xc, fo, pol);
                tma_load_4d(tb + P.off_hi, mh, &full[stage], 0, yhi, xc, fo, pol);
            }
            if (++stage == S)
            {
                stage = 0;
                par ^= 1u;
            }
        }
        return;
    }

    // -------------------------------- consumers --------------------------------
    const int ct   = tid - 32;
    const int lane = tid & 31;
    const int grp  = ct / P.tpo;
    const int lt   = ct - grp * P.tpo;
    const int orb  = orb0 + grp;
    const bool active = (lt < P.YG * P.nzv) && (orb < P.nfunc);
    const int zv   = active ? lt % P.nzv : 0;
    const int yg   = active ? lt / P.nzv : 0;
    const int z0   = zv * 4;
    const int rr0  = yg * RY;
    const int rb   = P.row_bytes;

    // byte offsets (inside a tile) of the rows this thread walks: tile rows
    // rr0-G .. rr0+RY+G-1, the first G possibly in the low-halo box, the last G
    // possibly in the high-halo box
    uint32_t rowoff[NW];
#pragma unroll
    for (int i = 0; i < NW; i++)
    {
        const int t = rr0 - G + i;
        int o;
        if (t < 0)
            o = (t + G) * rb;
        else if (t < P.TY)
            o = P.off_mid + t * rb;
        else
            o = P.off_hi + (t - P.TY) * rb;
        rowoff[i] = (uint32_t)o;
    }
    // z neighbours: G values left of z0, G values right of z0+3
    int zl = z0 - G, zr = z0 + 4;
    float ml = 1.f, mr = 1.f;
    if (zl < 0)
    {
        zl += P.nz;
        if (!P.per[2]) ml = 0.f;
    }
    if (zr >= P.nz)
    {
        zr -= P.nz;
        if (!P.per[2]) mr = 0.f;
    }
    const uint32_t zoff  = (uint32_t)(z0 * 4);
    const uint32_t zloff = (uint32_t)(zl * 4);
    const uint32_t zroff = (uint32_t)(zr * 4);
    // P3 with z split: the G columns left of z = 0 / right of z = nz-1 sit in the tile's
    // z-halo arrays (16 bytes per tile row); per walked row the offset of its entry
    bool isL = false, isR = false;
    uint32_t zrow[P3 ? NW : 1];
    uint32_t zLb = 0, zRb = 0;
    if constexpr (P3)
    {
        if (P.zsplit)
        {
            isL = z0 == 0;
            isR = z0 + 4 == P.nz;
            ml = mr = 1.f; // a Dirichlet end delivers zero columns (out-of-range box)
        }
#pragma unroll
        for (int i = 0; i < NW; i++)
        {
            const int t = rr0 - G + i;
            zrow[i]     = (uint32_t)(t < 0 ? (t + G) * 32
                                           : (t < P.TY ? P.zarr_mid + t * 32
                                                       : P.zarr_hi + (t - P.TY) * 32));
        }
        zLb = (uint32_t)(P.zlo_off + 16 - G * 4);
        zRb = (uint32_t)(P.zlo_off + 16);
    }
    // address of the left / right z neighbours of walked row i in the tile at `base`
    auto zl_addr = [&](uint32_t base, int i) -> uint32_t {
        if constexpr (P3)
            if (isL) return base + zLb + zrow[i];
        return base + rowoff[i] + zloff;
    };
    auto zr_addr = [&](uint32_t base, int i) -> uint32_t {
        if constexpr (P3)
            if (isR) return base + zRb + zrow[i];
        return base + rowoff[i] + zroff;
    };

    const uint32_t stage0  = smem_u32(smem + kMgBarBytes);
    const uint32_t grp_off = (uint32_t)(grp * P.tile_bytes);
    const long long plane  = (long long)P.ny * P.nz;
    const long long pt0    = (long long)(y0 + rr0) * P.nz + z0;

    const float sh = P.sh, sl = P.sl, oh = P.oh, ol = P.ol;
    auto ldrow = [&](uint32_t a, float(&o)[4]) { lds4(a, o); };
    auto ld1   = [&](uint32_t a, float m) { return lds1(a) * m; };

    // right-hand side of the first output plane (see the prefetch below)
    float4 fnext[RY];
    if (!SCALE && active)
    {
        const float* fp = P.f + (long long)orb * P.ld_f + (long long)xb * plane + pt0;
#pragma unroll
        for (int r = 0; r < RY; r++)
            fnext[r] = __ldg(reinterpret_cast<const float4*>(fp + (long long)r * P.nz));
    }

    int st_new = 0;       // stage of the plane arriving in this iteration
    uint32_t par_new = 0;
    int st_old = 0;       // stage of the oldest plane still needed

    for (int it = 0; it < nplanes; it++)
    {
        mbar_wait(&full[st_new], par_new);
        if (it >= 2 * G)
        {
            const int q = xb + it - 2 * G; // output plane
            if (active)
            {
                // tile bases of planes q-G .. q+G
                uint32_t tb[2 * G + 1];
                {
                    int s = st_old;
#pragma unroll
                    for (int d = 0; d <= 2 * G; d++)
                    {
                        tb[d] = stage0 + (uint32_t)s * (uint32_t)P.stage_bytes + grp_off;
                        if (++s == S) s = 0;
                    }
                }
                const long long o0 = (long long)q * plane + pt0;
                const bool zx      = P.zero_low[0] && q == 0;

                // localization mask of this plane (slab q / sub0 of this
                // color): requested now, consumed when the rows are stored
                int mo = -1;
                float4 mv[MASK ? RY : 1];
                if constexpr (MASK)
                {
                    const int iloc = q / P.mask.sub0;
                    mo             = __ldg(P.mask.off + (long long)orb * P.mask.subdivx + iloc);
                    if (mo >= 0)
                    {
                        const float* mp = P.mask.pool + (long long)mo * P.mask.slab
                                          + (long long)(q - iloc * P.mask.sub0) * plane + pt0;
#pragma unroll
                        for (int r = 0; r < RY; r++)
                            mv[r] = __ldg(reinterpret_cast<const float4*>(mp + (long long)r * P.nz));
                    }
                }

                // Right-hand side: the values of this plane were requested one
                // iteration ago (fnext) and the next plane's are requested
                // now, so the global loads are in flight during a whole plane
                // of shared-memory work.  (SCALE: the tile IS f -- the centre
                // values the stencil reads anyway.)
                float4 fv[RY];
                if constexpr (!SCALE)
                {
#pragma unroll
                    for (int r = 0; r < RY; r++)
                        fv[r] = fnext[r];
                    if (it + 1 < nplanes)
                    {
                        const float* fp = P.f + (long long)orb * P.ld_f + o0 + plane;
#pragma unroll
                        for (int r = 0; r < RY; r++)
                            fnext[r] = __ldg(reinterpret_cast<const float4*>(fp + (long long)r * P.nz));
                    }
                }

                auto finish = [&](int r, const float(&cen0)[4], const float(&av0)[4]) {
                    // SCALE: the stencil ran on the tile's f values; v = s f at
                    // the centre and A v = s (A f) (A is linear: one product with
                    // the two-float constant per point instead of one per tap;
                    // differs from rounding every v_k by < |c|_1 2^-24 |v|, i.e.
                    // by less than the rounding of v' itself after the damping)
                    float fr[4], cen[4], av[4];
#pragma unroll
                    for (int e = 0; e < 4; e++)
                    {
                        if constexpr (SCALE)
                        {
                            fr[e]  = cen0[e];
                            cen[e] = mul_split(sh, sl, cen0[e]);
                            av[e]  = mul_split(sh, sl, av0[e]);
                        }
                        else
                        {
                            fr[e]  = (e == 0) ? fv[r].x : (e == 1) ? fv[r].y : (e == 2) ? fv[r].z
                                                                                       : fv[r].w;
                            cen[e] = cen0[e];
                            av[e]  = av0[e];
                        }
                    }
                    float vn[4], wn[4];
#pragma unroll
                    for (int e = 0; e < 4; e++)
                    {
                        wn[e] = __fsub_rn(av[e], fr[e]);                        // w -= f
                        vn[e] = __fadd_rn(cen[e], mul_split(oh, ol, wn[e])); // v += -omega w
                    }
                    if constexpr (MASK)
                    {
                        // v.app_mask(level) after the sweep, work.app_mask(level)
                        // before the restriction (Preconditioning.cc:176,184)
                        if (mo == -2)
                        {
#pragma unroll
                            for (int e = 0; e < 4; e++)
                                vn[e] = wn[e] = 0.f;
                        }
                        else if (mo >= 0)
                        {
                            const float mm[4] = { mv[r].x, mv[r].y, mv[r].z, mv[r].w };
#pragma unroll
                            for (int e = 0; e < 4; e++)
                            {
                                vn[e] = mask_apply(P.mask.op, vn[e], mm[e]);
                                wn[e] = mask_apply(P.mask.op, wn[e], mm[e]);
                            }
                        }
                    }
                    const bool zrow = zx || (P.zero_low[1] && (y0 + rr0 + r) == 0);
                    if (zrow)
                    {
#pragma unroll
                        for (int e = 0; e < 4; e++)
                            vn[e] = wn[e] = 0.f;
                    }
                    if (P.zero_low[2] && z0 == 0) vn[0] = wn[0] = 0.f;
                    const long long oo = o0 + (long long)r * P.nz;
                    if (P.vout)
                        *reinterpret_cast<float4*>(P.vout + (long long)orb * P.ld_out + oo)
                            = make_float4(vn[0], vn[1], vn[2], vn[3]);
                    if (P.vout64)
                    {
                        double* d = P.vout64 + (long long)orb * P.ld_out + oo;
                        *reinterpret_cast<double2*>(d) = make_double2((double)vn[0], (double)vn[1]);
                        *reinterpret_cast<double2*>(d + 2)
                            = make_double2((double)vn[2], (double)vn[3]);
                    }
                    if (P.wout)
                        *reinterpret_cast<float4*>(P.wout + (long long)orb * P.ld_w + oo)
                            = make_float4(wn[0], wn[1], wn[2], wn[3]);
                };

                if constexpr (LAP == kLap4)
                {
                    const float c1x = P.c[0], c2x = P.c[1], c1y = P.c[2], c2y = P.c[3],
                                c1z = P.c[4], c2z = P.c[5];
                    float W[5][4];
                    ldrow(tb[2] + rowoff[0] + zoff, W[0]);
                    ldrow(tb[2] + rowoff[1] + zoff, W[1]);
                    ldrow(tb[2] + rowoff[2] + zoff, W[2]);
                    ldrow(tb[2] + rowoff[3] + zoff, W[3]);
#pragma unroll
                    for (int r = 0; r < RY; r++)
                    {
                        ldrow(tb[2] + rowoff[r + 4] + zoff, W[4]);
                        const uint32_t ro = rowoff[r + 2];
                        float L2, L1, R1, R2;
                        lds2(zl_addr(tb[2], r + 2), L2, L1);
                        lds2(zr_addr(tb[2], r + 2), R1, R2);
                        L2 *= ml;
                        L1 *= ml;
                        R1 *= mr;
                        R2 *= mr;
                        float Xm2[4], Xm1[4], Xp1[4], Xp2[4];
                        ldrow(tb[0] + ro + zoff, Xm2);
                        ldrow(tb[1] + ro + zoff, Xm1);
                        ldrow(tb[3] + ro + zoff, Xp1);
                        ldrow(tb[4] + ro + zoff, Xp2);
                        float av[4];
#pragma unroll
                        for (int e = 0; e < 4; e++)
                        {
                            const float c   = W[2][e];
                            const float zm1 = (e == 0) ? L1 : W[2][e > 0 ? e - 1 : 0];
                            const float zm2 = (e == 0) ? L2 : (e == 1) ? L1 : W[2][e > 1 ? e - 2 : 0];
                            const float zp1 = (e == 3) ? R1 : W[2][e < 3 ? e + 1 : 0];
                            const float zp2 = (e == 3) ? R2 : (e == 2) ? R1 : W[2][e < 2 ? e + 2 : 0];
                            const float tx = fmaf(c2x, dd(Xm2[e], Xp2[e], c), c1x * dd(Xm1[e], Xp1[e], c));
                            const float ty = fmaf(c2y, dd(W[0][e], W[4][e], c), c1y * dd(W[1][e], W[3][e], c));
                            const float tz = fmaf(c2z, dd(zm2, zp2, c), c1z * dd(zm1, zp1, c));
                            av[e]          = tx + (ty + tz);
                        }
                        finish(r, W[2], av);
#pragma unroll
                        for (int e = 0; e < 4; e++)
                        {
                            W[0][e] = W[1][e];
                            W[1][e] = W[2][e];
                            W[2][e] = W[3][e];
                            W[3][e] = W[4][e];
                        }
                    }
                }
                else if constexpr (LAP == kLap2)
                {
                    const float c1x = P.c[0], c1y = P.c[1], c1z = P.c[2];
                    float W[3][4];
                    ldrow(tb[1] + rowoff[0] + zoff, W[0]);
                    ldrow(tb[1] + rowoff[1] + zoff, W[1]);
#pragma unroll
                    for (int r = 0; r < RY; r++)
                    {
                        ldrow(tb[1] + rowoff[r + 2] + zoff, W[2]);
                        const uint32_t ro = rowoff[r + 1];
                        const float L = ld1(zl_addr(tb[1], r + 1), ml);
                        const float R = ld1(zr_addr(tb[1], r + 1), mr);
                        float Xm[4], Xp[4];
                        ldrow(tb[0] + ro + zoff, Xm);
                        ldrow(tb[2] + ro + zoff, Xp);
                        float av[4];
#pragma unroll
                        for (int e = 0; e < 4; e++)
                        {
                            const float c  = W[1][e];
                            const float zm = (e == 0) ? L : W[1][e > 0 ? e - 1 : 0];
                            const float zp = (e == 3) ? R : W[1][e < 3 ? e + 1 : 0];
                            float a        = c1x * dd(Xm[e], Xp[e], c);
                            a              = fmaf(c1y, dd(W[0][e], W[2][e], c), a);
                            a              = fmaf(c1z, dd(zm, zp, c), a);
                            av[e]          = a;
                        }
                        finish(r, W[1], av);
#pragma unroll
                        for (int e = 0; e < 4; e++)
                        {
                            W[0][e] = W[1][e];
                            W[1][e] = W[2][e];
                        }
                    }
                }
                else
                {
                    // Mehrstellen 19-point A (src/pb/FDkernels.cc:447-520)
                    const float cx = P.c[0], cy = P.c[1], cz = P.c[2], cxy = P.c[3],
                                cyz = P.c[4], cxz = P.c[5];
                    // per plane d: rows (r-1, r, r+1) with their z neighbours
                    float V[3][3][4], L[3][3], R[3][3];
                    auto load = [&](int d, int slot, int i) {
                        const uint32_t a = tb[d] + rowoff[i];
                        ldrow(a + zoff, V[d][slot]);
                        L[d][slot] = ld1(zl_addr(tb[d], i), ml);
                        R[d][slot] = ld1(zr_addr(tb[d], i), mr);
                    };
#pragma unroll
                    for (int d = 0; d < 3; d++)
                    {
                        load(d, 0, 0);
                        load(d, 1, 1);
                    }
#pragma unroll
                    for (int r = 0; r < RY; r++)
                    {
#pragma unroll
                        for (int d = 0; d < 3; d++)
                            load(d, 2, r + 2);
                        float av[4];
#pragma unroll
                        for (int e = 0; e < 4; e++)
                        {
                            const float c = V[1][1][e];
                            auto zm = [&](int d, int s) {
                                return (e == 0) ? L[d][s] : V[d][s][e > 0 ? e - 1 : 0];
                            };
                            auto zp = [&](int d, int s) {
                                return (e == 3) ? R[d][s] : V[d][s][e < 3 ? e + 1 : 0];
                            };
                            const float sx  = dd(V[0][1][e], V[2][1][e], c);
                            const float sy  = dd(V[1][0][e], V[1][2][e], c);
                            const float sz  = dd(zm(1, 1), zp(1, 1), c);
                            const float sxz = dd(zm(0, 1), zp(0, 1), c) + dd(zm(2, 1), zp(2, 1), c);
                            const float syz = dd(zm(1, 0), zp(1, 0), c) + dd(zm(1, 2), zp(1, 2), c);
                            const float sxy = dd(V[0][0][e], V[2][0][e], c)
                                              + dd(V[0][2][e], V[2][2][e], c);
                            float a = cx * sx;
                            a       = fmaf(cy, sy, a);
                            a       = fmaf(cz, sz, a);
                            float b = cxy * sxy;
                            b       = fmaf(cyz, syz, b);
                            b       = fmaf(cxz, sxz, b);
                            av[e]   = a + b;
                        }
                        finish(r, V[1][1], av);
#pragma unroll
                        for (int d = 0; d < 3; d++)
                        {
#pragma unroll
                            for (int e = 0; e < 4; e++)
                            {
                                V[d][0][e] = V[d][1][e];
                                V[d][1][e] = V[d][2][e];
                            }
                            L[d][0] = L[d][1];
                            L[d][1] = L[d][2];
                            R[d][0] = R[d][1];
                            R[d][1] = R[d][2];
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st_old]);
            if (++st_old == S) st_old = 0;
        }
        if (++st_new == S)
        {
            st_new = 0;
            par_new ^= 1u;
        }
    }
}

// ---------------------------------------------------------------------------
// Restriction: coarse(i,j,k) = sum over the 27 fine neighbours of (2i,2j,2k)
// with weights (1,2,1)^3 / 64, evaluated separably in float.  One thread
// produces 4 consecutive coarse k.
// ---------------------------------------------------------------------------
__global__ void k_mg_restrict(int nxc, int nyc, int nzc, int perx, int pery, int perz,
    const float* __restrict__ fine, long long ldf, float* __restrict__ coarse,
    long long ldc, MaskView mask, const float* __restrict__ fine_w,
    const int* __restrict__ map_w, int f0)
{
    // block (zx, zy): zx threads along the coarse k-vectors, zy coarse rows;
    // grid.x tiles (k-vector, row), grid.y = coarse plane, grid.z = function
    const int nzv = nzc >> 2;
    const int tz  = (nzv + blockDim.x - 1) / blockDim.x;
    const int kv  = (blockIdx.x % tz) * blockDim.x + threadIdx.x;
    const int j   = (blockIdx.x / tz) * blockDim.y + threadIdx.y;
    const int i   = blockIdx.y;
    const int f   = blockIdx.z;
    if (kv >= nzv || j >= nyc) return;
    const int nx = 2 * nxc, ny = 2 * nyc, nz = 2 * nzc;
    const float* F = fine + (long long)f * ldf;
    const int k0 = kv * 4;
    float acc[4] = { 0.f, 0.f, 0.f, 0.f };
#pragma unroll
    for (int dx = -1; dx <= 1; dx++)
    {
        int x = 2 * i + dx;
        float wx = (dx == 0) ? 2.f : 1.f;
        const float* Fx = F;
        if (x < 0)
        {
            // fine plane -1: my own last plane (periodic, x not split), the
            // west neighbour's last plane (x split), or nothing (Dirichlet)
            // index in the neighbour's (whole) block; f is relative to f0
            const int fw = (fine_w && map_w) ? map_w[f] : f + f0;
            if (fine_w && fw >= 0)
            {
                Fx = fine_w + (long long)fw * ldf;
                x += nx;
            }
            else if (fine_w)
                wx = 0.f; // the west rank does not hold this orbital
            else if (perx)
                x += nx;
            else
                wx = 0.f;
        }
        float ay[4] = { 0.f, 0.f, 0.f, 0.f };
#pragma unroll
        for (int dy = -1; dy <= 1; dy++)
        {
            int y = 2 * j + dy;
            float wy = (dy == 0) ? 2.f : 1.f;
            if (y < 0)
            {
                if (pery)
                    y += ny;
                else
                    wy = 0.f;
            }
            const float* row = Fx + ((long long)(x < 0 ? 0 : x) * ny + (y < 0 ? 0 : y)) * nz;
            const float4 a = __ldg(reinterpret_cast<const float4*>(row + 2 * k0));
            const float4 b = __ldg(reinterpret_cast<const float4*>(row + 2 * k0 + 4));
            float lft;
            if (k0 > 0)
                lft = __ldg(row + 2 * k0 - 1);
            else
                lft = perz ? __ldg(row + nz - 1) : 0.f;
            // z reduction: fine 2k-1, 2k, 2k+1
            const float t0 = (lft + a.y) + 2.f * a.x;
            const float t1 = (a.y + a.w) + 2.f * a.z;
            const float t2 = (a.w + b.y) + 2.f * b.x;
            const float t3 = (b.y + b.w) + 2.f * b.z;
            ay[0] = fmaf(wy, t0, ay[0]);
            ay[1] = fmaf(wy, t1, ay[1]);
            ay[2] = fmaf(wy, t2, ay[2]);
            ay[3] = fmaf(wy, t3, ay[3]);
        }
#pragma unroll
        for (int e = 0; e < 4; e++)
            acc[e] = fmaf(wx, ay[e], acc[e]);
    }
    float* C = coarse + (long long)f * ldc + ((long long)i * nyc + j) * nzc + k0;
    float r[4] = { acc[0] * 0.015625f, acc[1] * 0.015625f, acc[2] * 0.015625f,
        acc[3] * 0.015625f };
    if (mask.off)
    {
        // rcoarse->app_mask(level + 1)  (Preconditioning.cc:192)
        const int iloc = i / mask.sub0;
        const int mo   = __ldg(mask.off + (long long)f * mask.subdivx + iloc);
        if (mo == -2)
            r[0] = r[1] = r[2] = r[3] = 0.f;
        else if (mo >= 0)
        {
            const float4 m = __ldg(reinterpret_cast<const float4*>(mask.pool
                + (long long)mo * mask.slab
                + ((long long)(i - iloc * mask.sub0) * nyc + j) * nzc + k0));
            r[0] = mask_apply(mask.op, r[0], m.x);
            r[1] = mask_apply(mask.op, r[1], m.y);
            r[2] = mask_apply(mask.op, r[2], m.z);
            r[3] = mask_apply(mask.op, r[3], m.w);
        }
    }
    *reinterpret_cast<float4*>(C) = make_float4(r[0], r[1], r[2], r[3]);
}


// the blocks a transfer kernel may touch on a decomposed box: index bit 2 = across x,
// bit 1 = across y, bit 0 = across z (the low side for the restriction, the high side for
// the prolongation); my own block where a direction is not split (periodic wrap)
struct Nb8
{
    const float* p[8];
};

// k_mg_restrict on any px x py x pz decomposition: fine plane / row / column -1 is read in
// place from the low neighbour's block (its last one); beyond a Dirichlet end it weighs 0
__global__ void k_mg_restrict3(int nxc, int nyc, int nzc, int endx, int endy, int endz, Nb8 nb,
    long long ldf, float* __restrict__ coarse, long long ldc, MaskView mask, int f0)
{
    const int nzv = nzc >> 2;
    const int tz  = (nzv + blockDim.x - 1) / blockDim.x;
    const int kv  = (blockIdx.x % tz) * blockDim.x + threadIdx.x;
    const int j   = (blockIdx.x / tz) * blockDim.y + threadIdx.y;
    const int i   = blockIdx.y;
    const int f   = blockIdx.z;
    if (kv >= nzv || j >= nyc) return;
    const int nx = 2 * nxc, ny = 2 * nyc, nz = 2 * nzc;
    const long long fo = (long long)(f + f0) * ldf;
    const int k0 = kv * 4;
    float acc[4] = { 0.f, 0.f, 0.f, 0.f };
#pragma unroll
    for (int dx = -1; dx <= 1; dx++)
    {
        int x = 2 * i + dx, bx = 0;
        float wx = (dx == 0) ? 2.f : 1.f;
        if (x < 0)
        {
            if (endx)
                wx = 0.f, x = 0;
            else
                x += nx, bx = 4;
        }
        float ay[4] = { 0.f, 0.f, 0.f, 0.f };
#pragma unroll
        for (int dy = -1; dy <= 1; dy++)
        {
            int y = 2 * j + dy, by = 0;
            float wy = (dy == 0) ? 2.f : 1.f;
            if (y < 0)
            {
                if (endy)
                    wy = 0.f, y = 0;
                else
                    y += ny, by = 2;
            }
            const long long ro = fo + ((long long)x * ny + y) * nz;
            const float* row   = nb.p[bx | by] + ro;
            const float4 a = __ldg(reinterpret_cast<const float4*>(row + 2 * k0));
            const float4 b = __ldg(reinterpret_cast<const float4*>(row + 2 * k0 + 4));
            float lft;
            if (k0 > 0)
                lft = __ldg(row + 2 * k0 - 1);
            else
                lft = endz ? 0.f : __ldg(nb.p[bx | by | 1] + ro + nz - 1);
            const float t0 = (lft + a.y) + 2.f * a.x;
            const float t1 = (a.y + a.w) + 2.f * a.z;
            const float t2 = (a.w + b.y) + 2.f * b.x;
            const float t3 = (b.y + b.w) + 2.f * b.z;
            ay[0] = fmaf(wy, t0, ay[0]);
            ay[1] = fmaf(wy, t1, ay[1]);
            ay[2] = fmaf(wy, t2, ay[2]);
            ay[3] = fmaf(wy, t3, ay[3]);
        }
#pragma unroll
        for (int e = 0; e < 4; e++)
            acc[e] = fmaf(wx, ay[e], acc[e]);
    }
    float* C = coarse + (long long)f * ldc + ((long long)i * nyc + j) * nzc + k0;
    float r[4] = { acc[0] * 0.015625f, acc[1] * 0.015625f, acc[2] * 0.015625f,
        acc[3] * 0.015625f };
    if (mask.off)
    {
        const int iloc = i / mask.sub0;
        const int mo   = __ldg(mask.off + (long long)f * mask.subdivx + iloc);
        if (mo == -2)
            r[0] = r[1] = r[2] = r[3] = 0.f;
        else if (mo >= 0)
        {
            const float4 m = __ldg(reinterpret_cast<const float4*>(mask.pool
                + (long long)mo * mask.slab
                + ((long long)(i - iloc * mask.sub0) * nyc + j) * nzc + k0));
            r[0] = mask_apply(mask.op, r[0], m.x);
            r[1] = mask_apply(mask.op, r[1], m.y);
            r[2] = mask_apply(mask.op, r[2], m.z);
            r[3] = mask_apply(mask.op, r[3], m.w);
        }
    }
    *reinterpret_cast<float4*>(C) = make_float4(r[0], r[1], r[2], r[3]);
}

// ---------------------------------------------------------------------------
// v -= P e: trilinear prolongation in the reference's float operand order
// (MGkernels.cc:60-198 collapse to a closed form per fine point, cf.
// k_extend3D in fd_ghosted.cu), then Preconditioning.cc:206.  One thread
// handles 4 consecutive fine z.  The result is stored with zero low layers
// where requested.
// ---------------------------------------------------------------------------
template <bool P3>
__global__ void k_mg_prolong_correct(int nx, int ny, int nz, int perx, int pery, int perz,
    int zlx, int zly, int zlz, const float* __restrict__ coarse, long long ldc,
    float* __restrict__ v, long long ldv, MaskView mask, const float* __restrict__ coarse_e,
    const int* __restrict__ map_e, int f0, Nb8 nb)
{
    // one thread: the 2 x 2 x 4 fine brick above coarse (cx0, cy0, cz0..cz0+1); the
    // four coarse rows it needs are loaded once and shared by the 16 fine points
    const int nzv = nz >> 2;
    const int nxc = nx >> 1, nyc = ny >> 1, nzc = nz >> 1;
    const int tz  = (nzv + blockDim.x - 1) / blockDim.x;
    const int zv  = (blockIdx.x % tz) * blockDim.x + threadIdx.x;
    const int cy0 = (blockIdx.x / tz) * blockDim.y + threadIdx.y;
    const int cx0 = blockIdx.y;
    const int f   = blockIdx.z;
    if (zv >= nzv || cy0 >= nyc) return;
    const int z0  = zv * 4;
    const int cz0 = z0 >> 1;
    int cx1 = cx0 + 1, cy1 = cy0 + 1, cz2 = cz0 + 2;
    float mx = 1.f, my = 1.f, mz = 1.f;
    const float* C   = coarse + (long long)f * ldc;
    const float* Cx1 = C; // block holding coarse plane cx1
    int bx = 0, by = 0, bz = 0; // P3: which neighbour holds plane cx1 / row cy1 / column cz2
    if constexpr (P3)
    {
        // perx / pery / perz here mean "not a Dirichlet end of the domain": coarse plane
        // nxc, row nyc, column nzc are the high neighbour's first ones, read in place
        if (cx1 == nxc)
        {
            cx1 = 0;
            if (perx) bx = 4; else mx = 0.f;
        }
        if (cy1 == nyc)
        {
            cy1 = 0;
            if (pery) by = 2; else my = 0.f;
        }
        if (cz2 == nzc)
        {
            cz2 = 0;
            if (perz) bz = 1; else mz = 0.f;
        }
    }
    else if (cx1 == nxc)
    {
        // coarse plane nxc: my own plane 0 (periodic, x not split), the east
        // neighbour's plane 0 (x split), or nothing (Dirichlet)
        cx1 = 0;
        const int fe = (coarse_e && map_e) ? map_e[f] : f + f0;
        if (coarse_e && fe >= 0)
            Cx1 = coarse_e + (long long)fe * ldc;
        else if (coarse_e || !perx)
            mx = 0.f; // the east rank does not hold it, or the domain ends
    }
    if (!P3 && cy1 == nyc)
    {
        cy1 = 0;
        if (!pery) my = 0.f;
    }
    if (!P3 && cz2 == nzc)
    {
        cz2 = 0;
        if (!perz) mz = 0.f;
    }
    // coarse rows (x0,y0), (x0,y1), (x1,y0), (x1,y1): values at cz0, cz0+1, cz0+2
    float c00[3], c01[3], c10[3], c11[3];
    auto row3 = [&](const float* Cb, int bits, int cx, int cy, float m, float(&o)[3]) {
        const long long ro = ((long long)cx * nyc + cy) * nzc;
        const float* r     = (P3 ? nb.p[bits] + (long long)(f + f0) * ldc : Cb) + ro;
        const float* rz    = (P3 ? nb.p[bits | bz] + (long long)(f + f0) * ldc : Cb) + ro;
        const float2 a = __ldg(reinterpret_cast<const float2*>(r + cz0));
        o[0] = a.x * m;
        o[1] = a.y * m;
        o[2] = __ldg(rz + cz2) * (m * mz);
    };
    row3(C, 0, cx0, cy0, 1.f, c00);
    row3(C, by, cx0, cy1, my, c01);
    row3(Cx1, bx, cx1, cy0, mx, c10);
    row3(Cx1, bx | by, cx1, cy1, mx * my, c11);

#pragma unroll
    for (int ox = 0; ox < 2; ox++)
#pragma unroll
        for (int oy = 0; oy < 2; oy++)
        {
            const int x = 2 * cx0 + ox, y = 2 * cy0 + oy;
            float w[4];
#pragma unroll
            for (int e = 0; e < 4; e++)
            {
                const int k   = e >> 1; // lower coarse index relative to cz0
                const bool oz = e & 1;
                // names follow k_extend3D: c[0]=lower corner, c[1]=+z, c[Y]=+y, c[X]=+x
                const float a0 = c00[k], a1 = c00[k + 1];
                float val;
                if (!ox && !oy)
                    val = oz ? 0.5f * (a0 + a1) : a0;
                else if (!ox && oy)
                {
                    const float y0v = c01[k], y1v = c01[k + 1];
                    val = oz ? 0.25f * (((y1v + a1) + y0v) + a0) : 0.5f * (y0v + a0);
                }
                else if (ox && !oy)
                {
                    const float x0v = c10[k], x1v = c10[k + 1];
                    val = oz ? 0.25f * (((x1v + x0v) + a1) + a0) : 0.5f * (x0v + a0);
                }
                else
                {
                    const float y0v = c01[k], y1v = c01[k + 1];
                    const float x0v = c10[k], x1v = c10[k + 1];
                    const float q0v = c11[k], q1v = c11[k + 1];
                    val = oz ? 0.125f
                                   * (((((((q1v + q0v) + x1v) + x0v) + y1v) + y0v) + a1) + a0)
                             : 0.25f * (((q0v + x0v) + y0v) + a0);
                }
                w[e] = val;
            }
            if (mask.off)
            {
                // gfv_work_[level]->app_mask(level) on P e  (Preconditioning.cc:204)
                const int iloc = x / mask.sub0;
                const int mo   = __ldg(mask.off + (long long)f * mask.subdivx + iloc);
                if (mo == -2)
                    w[0] = w[1] = w[2] = w[3] = 0.f;
                else if (mo >= 0)
                {
                    const float4 m = __ldg(reinterpret_cast<const float4*>(mask.pool
                        + (long long)mo * mask.slab
                        + ((long long)(x - iloc * mask.sub0) * ny + y) * nz + z0));
                    w[0] = mask_apply(mask.op, w[0], m.x);
                    w[1] = mask_apply(mask.op, w[1], m.y);
                    w[2] = mask_apply(mask.op, w[2], m.z);
                    w[3] = mask_apply(mask.op, w[3], m.w);
                }
            }
            float* pv = v + (long long)f * ldv + ((long long)x * ny + y) * nz + z0;
            float4 vv = *reinterpret_cast<float4*>(pv);
            vv.x = __fsub_rn(vv.x, w[0]);
            vv.y = __fsub_rn(vv.y, w[1]);
            vv.z = __fsub_rn(vv.z, w[2]);
            vv.w = __fsub_rn(vv.w, w[3]);
            if ((zlx && x == 0) || (zly && y == 0)) vv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (zlz && z0 == 0) vv.x = 0.f;
            *reinterpret_cast<float4*>(pv) = vv;
        }
}

// ORBDTYPE double -> float copy of the residual (OrbitalsPreconditioning.cc:103)
__global__ void k_mg_convert(long long npt, const double* __restrict__ in, long long ldi,
    float* __restrict__ out, long long ldo)
{
    const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (t >= npt) return;
    const int f = blockIdx.y;
    const double2 a = *reinterpret_cast<const double2*>(in + (long long)f * ldi + t);
    const double2 b = *reinterpret_cast<const double2*>(in + (long long)f * ldi + t + 2);
    *reinterpret_cast<float4*>(out + (long long)f * ldo + t)
        = make_float4((float)a.x, (float)a.y, (float)b.x, (float)b.y);
}

// v = (float)(c * (double)f) with zero low layers: the start vector of a level
// for boxes with a Dirichlet direction (periodic boxes fold this product into
// the first sweep's tile loads)
__global__ void k_mg_scale(int nx, int ny, int nz, int zlx, int zly, int zlz, float ch,
    float cl, const float* __restrict__ f, long long ldf, float* __restrict__ v,
    long long ldv, MaskView mask)
{
    const int nzv = nz >> 2;
    const int tz  = (nzv + blockDim.x - 1) / blockDim.x;
    const int zv  = (blockIdx.x % tz) * blockDim.x + threadIdx.x;
    const int y   = (blockIdx.x / tz) * blockDim.y + threadIdx.y;
    const int x   = blockIdx.y;
    const int fn  = blockIdx.z;
    if (zv >= nzv || y >= ny) return;
    const long long o = ((long long)x * ny + y) * nz + zv * 4;
    const float4 a = __ldg(reinterpret_cast<const float4*>(f + (long long)fn * ldf + o));
    float4 r = make_float4(mul_split(ch, cl, a.x), mul_split(ch, cl, a.y),
        mul_split(ch, cl, a.z), mul_split(ch, cl, a.w));
    if (mask.off)
    {
        // coarse levels: the first sweep from v = 0 gives omega * f, then
        // v.app_mask(level)  (Preconditioning.cc:175-176)
        const int iloc = x / mask.sub0;
        const int mo   = __ldg(mask.off + (long long)fn * mask.subdivx + iloc);
        if (mo == -2)
            r = make_float4(0.f, 0.f, 0.f, 0.f);
        else if (mo >= 0)
        {
            const float4 m = __ldg(reinterpret_cast<const float4*>(mask.pool
                + (long long)mo * mask.slab
                + ((long long)(x - iloc * mask.sub0) * ny + y) * nz + zv * 4));
            r.x = mask_apply(mask.op, r.x, m.x);
            r.y = mask_apply(mask.op, r.y, m.y);
            r.z = mask_apply(mask.op, r.z, m.z);
            r.w = mask_apply(mask.op, r.w, m.w);
        }
    }
    if ((zlx && x == 0) || (zly && y == 0)) r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (zlz && zv == 0) r.x = 0.f;
    *reinterpret_cast<float4*>(v + (long long)fn * ldv + o) = r;
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
static int round_up_i(int a, int b) { return (a + b - 1) / b * b; }

struct JacobiCfg
{
    int RY, YG, NB, S, XC;
};

static bool g_jacobi_p3 = false; // lay out the z-halo arrays of the P3 kernels
static bool jacobi_layout(const JacobiCfg& c, int G, int ny, int nz, JacobiParams& P,
    size_t& smem)
{
    const int TY = c.RY * c.YG;
    if (TY > ny || ny % TY || TY < G) return false;
    P.TY  = TY;
    P.YG  = c.YG;
    P.NB  = c.NB;
    P.S   = c.S;
    P.XC  = c.XC;
    P.nzv = nz / 4;
    if (c.YG * P.nzv > 512) return false;
    P.tpo = round_up_i(c.YG * P.nzv, 32);
    if (32 + c.NB * P.tpo > 544) return false;
    P.row_bytes   = nz * 4;
    P.off_mid     = round_up_i(G * P.row_bytes, 128);
    P.off_hi      = P.off_mid + round_up_i(TY * P.row_bytes, 128);
    P.tile_bytes  = P.off_hi + round_up_i(G * P.row_bytes, 128);
    if (g_jacobi_p3)
    {
        P.zlo_off    = P.tile_bytes;
        P.zarr_mid   = 128;
        P.zarr_hi    = 128 + round_up_i(TY * 32, 128);
        P.zarr_bytes = P.zarr_hi + 128;
        P.tile_bytes += P.zarr_bytes;
    }
    P.stage_bytes = c.NB * P.tile_bytes;
    smem          = (size_t)kMgBarBytes + (size_t)c.S * P.stage_bytes;
    return smem <= 227 * 1024 && c.S >= 2 * G + 2 && c.S <= 16;
}

static bool jacobi_choose(int G, int nx, int ny, int nz, int nfunc, JacobiCfg& best)
{
    if (const char* env = getenv("MGB_MG_CFG"))
    {
        JacobiCfg c;
        if (sscanf(env, "%d,%d,%d,%d,%d", &c.RY, &c.YG, &c.NB, &c.S, &c.XC) == 5)
        {
            if (c.XC <= 0 || c.XC > nx) c.XC = nx;
            JacobiParams tmp;
            size_t sm;
            if ((c.RY == 4 || c.RY == 2) && c.RY >= G && jacobi_layout(c, G, ny, nz, tmp, sm))
            {
                best = c;
                return true;
            }
        }
    }
    double best_cost = 1e30;
    bool found       = false;
    for (int ry = 4; ry >= 2; ry -= 2)
        for (int yg = 1; yg <= 64; yg++)
            for (int nb = 1; nb <= 8; nb++)
                for (int s = 14; s >= 2 * G + 2; s--)
                {
                    JacobiCfg c = { ry, yg, nb, s, nx };
                    if (nb > nfunc && nb > 1) continue;
                    JacobiParams tmp;
                    size_t sm;
                    if (!jacobi_layout(c, G, ny, nz, tmp, sm)) continue;
                    const int TY      = ry * yg;
                    const int threads = 32 + nb * tmp.tpo;
                    double cost = (double)(TY + 2 * G) / TY;          // L2 -> SM re-reads
                    cost *= (double)tmp.tpo / (yg * tmp.nzv);          // idle lanes
                    if (threads < 256) cost *= 1.0 + 0.3 * (256 - threads) / 256.;
                    if (ry == 2) cost *= 1.05;
                    if (threads > 416) cost *= 1.15; // 96-register class
                    // bytes the TMA ring keeps in flight per SM beyond the 2G+1
                    // resident planes; below ~64 KB the HBM latency shows
                    const double inflight = (double)(s - (2 * G + 1)) * tmp.stage_bytes;
                    if (inflight < 65536.) cost *= 1.0 + 0.4 * (65536. - inflight) / 65536.;
                    cost *= 1.0 + 0.002 * s; // no deeper than useful
                    if (cost < best_cost)
                    {
                        best_cost = cost;
                        best      = c;
                        found     = true;
                    }
                }
    if (!found) return false;
    // enough CTAs for ~4 waves: split x into chunks if needed
    const long long ctas
        = (long long)(ny / (best.RY * best.YG)) * ((nfunc + best.NB - 1) / best.NB);
    if (ctas < 4 * 148)
    {
        int chunks = (int)((4 * 148 + ctas - 1) / ctas);
        int xc     = (nx + chunks - 1) / chunks;
        const int xmin = 4 * G;
        if (xc < xmin) xc = nx < xmin ? nx : xmin;
        best.XC = xc;
    }
    return true;
}

static void split_double(double c, float& h, float& l)
{
    h = (float)c;
    l = (float)(c - (double)h);
}

bool mg_fused_level_ok(const mgb_grid& g, int lap_type)
{
    const int G = (lap_type == MGB_LAP_4) ? 2 : 1;
    if (lap_type != MGB_LAP_4M && lap_type != MGB_LAP_2 && lap_type != MGB_LAP_4) return false;
    const int nx = g.dim[0], ny = g.dim[1], nz = g.dim[2];
    if (nz % 4 || nz < 4 || nz > 512 || nz < 2 * G) return false;
    if (nx < 1 || ny < 2 * G) return false;
    JacobiCfg c;
    return jacobi_choose(G, nx, ny, nz, 1, c);
}

template <int LAP, bool SCALE, bool MASK>
static int launch_jacobi(const JacobiParams& P, const JacobiCfg& c, dim3 grid, int threads,
    size_t smem, cudaStream_t st)
{
    // launch-bound classes: 9 / 13 / 17 warps -> 168 / 128 / 96 registers
#define MGB_LAUNCH_JP(RYV, MT, P3V)                                            \
    {                                                                          \
        auto kern = k_mg_jacobi<LAP, RYV, SCALE, MT, MASK, P3V>;               \
        MGB_CUDA(cudaFuncSetAttribute(                                         \
            kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        kern<<<grid, threads, smem, st>>>(P);                                  \
    }
#define MGB_LAUNCH_J(RYV, MT)                                                  \
    if (P.nbmaps)                                                              \
        MGB_LAUNCH_JP(RYV, MT, true)                                           \
    else                                                                       \
        MGB_LAUNCH_JP(RYV, MT, false)
#define MGB_LAUNCH_JR(RYV)                                                     \
    if (threads <= 288)                                                        \
        MGB_LAUNCH_J(RYV, 288)                                                 \
    else if (threads <= 416)                                                   \
        MGB_LAUNCH_J(RYV, 416)                                                 \
    else                                                                       \
        MGB_LAUNCH_J(RYV, 544)
    if (c.RY == 4)
    {
        MGB_LAUNCH_JR(4)
    }
    else
    {
        MGB_LAUNCH_JR(2)
    }
#undef MGB_LAUNCH_JR
#undef MGB_LAUNCH_J
#undef MGB_LAUNCH_JP
    MGB_LAUNCHED("k_mg_jacobi");
    return MGB_OK;
}

int mg_jacobi(const MgJacobiArgs& a, cudaStream_t st)
{
    const mgb_grid& gr = *a.grid;
    const int nx = gr.dim[0], ny = gr.dim[1], nz = gr.dim[2];
    const int G = (a.lap_type == MGB_LAP_4) ? 2 : 1;
    JacobiCfg c;
    g_jacobi_p3 = a.nb3d != nullptr;
    if (!jacobi_choose(G, nx, ny, nz, a.nfunc, c))
    {
        g_jacobi_p3 = false;
        set_error("mg_jacobi: no tile configuration for %d x %d x %d", nx, ny, nz);
        return MGB_ENOTSUP;
    }
    JacobiParams P;
    memset(&P, 0, sizeof(P));
    size_t smem       = 0;
    const bool lay_ok = jacobi_layout(c, G, ny, nz, P, smem);
    g_jacobi_p3       = false;
    if (!lay_ok) return MGB_ENOTSUP;
    int rc;
    if (a.nb3d)
    {
        if ((rc = nb_table(reinterpret_cast<const void* const*>(a.nb3d), a.nbz, false, nz, ny, nx,
                 (long long)a.ld_in, a.nfunc, P.TY, G, st, &P.nbmaps)))
            return rc;
        P.zsplit = a.nbz != nullptr;
        for (int d = 0; d < 3; d++)
        {
            P.end_lo[d] = gr.bc[d] != 1 && gr.coord[d] == 0;
            P.end_hi[d] = gr.bc[d] != 1 && gr.coord[d] == gr.nproc[d] - 1;
        }
    }
    if ((rc = make_map(&P.in_mid, false, a.in, 4, nz, ny, nx, (long long)a.ld_in, a.nfunc, P.TY)))
        return rc;
    if ((rc = make_map(&P.in_halo, false, a.in, 4, nz, ny, nx, (long long)a.ld_in, a.nfunc, G)))
        return rc;
    P.split_x = gr.nproc[0] > 1 && !a.nb3d;
    if (P.split_x)
    {
        const bool per = gr.bc[0] == 1;
        P.have_w = (per || gr.coord[0] > 0) && a.peer_w;
        P.have_e = (per || gr.coord[0] < gr.nproc[0] - 1) && a.peer_e;
        if (P.have_w)
        {
            if ((rc = make_map(&P.pw_mid, false, a.peer_w, 4, nz, ny, nx, (long long)a.ld_in,
                     a.nfunc, P.TY)))
                return rc;
            if ((rc = make_map(
                     &P.pw_halo, false, a.peer_w, 4, nz, ny, nx, (long long)a.ld_in, a.nfunc, G)))
                return rc;
        }
        if (P.have_e)
        {
            if ((rc = make_map(&P.pe_mid, false, a.peer_e, 4, nz, ny, nx, (long long)a.ld_in,
                     a.nfunc, P.TY)))
                return rc;
            if ((rc = make_map(
                     &P.pe_halo, false, a.peer_e, 4, nz, ny, nx, (long long)a.ld_in, a.nfunc, G)))
                return rc;
        }
    }
    P.map_w  = a.map_w;
    P.map_e  = a.map_e;
    P.f      = a.f;
    P.vout   = a.out;
    P.vout64 = a.out64;
    P.wout   = a.w;
    P.ld_f   = (long long)a.ld_f;
    P.ld_out = (long long)a.ld_out;
    P.ld_w   = (long long)a.ld_w;
    P.nx     = nx;
    P.ny     = ny;
    P.nz     = nz;
    P.nfunc  = a.nfunc;
    for (int d = 0; d < 3; d++)
    {
        P.per[d]      = gr.bc[d] == 1;
        P.zero_low[d] = a.zero_low[d];
    }
    split_double(a.scale, P.sh, P.sl);
    split_double(-a.omega, P.oh, P.ol);
    const double inv12 = 1. / 12.;
    const double i2[3] = { 1. / (gr.h[0] * gr.h[0]), 1. / (gr.h[1] * gr.h[1]),
        1. / (gr.h[2] * gr.h[2]) };
    if (a.lap_type == MGB_LAP_4)
    {
        for (int d = 0; d < 3; d++)
        {
            P.c[2 * d]     = (float)(-16. * inv12 * i2[d]);
            P.c[2 * d + 1] = (float)(inv12 * i2[d]);
        }
    }
    else if (a.lap_type == MGB_LAP_2)
    {
        for (int d = 0; d < 3; d++)
            P.c[d] = (float)(-i2[d]);
    }
    else
    {
        const double c0 = 16. * inv12 * (i2[0] + i2[1] + i2[2]);
        P.c[0] = (float)(-10. * inv12 * i2[0] + 0.125 * c0);
        P.c[1] = (float)(-10. * inv12 * i2[1] + 0.125 * c0);
        P.c[2] = (float)(-10. * inv12 * i2[2] + 0.125 * c0);
        P.c[3] = (float)(-inv12 * (i2[0] + i2[1]));
        P.c[4] = (float)(-inv12 * (i2[2] + i2[1]));
        P.c[5] = (float)(-inv12 * (i2[0] + i2[2]));
    }
    dim3 grid((unsigned)(ny / P.TY), (unsigned)((a.nfunc + c.NB - 1) / c.NB),
        (unsigned)((nx + c.XC - 1) / c.XC));
    if (grid.y > 65535 || grid.z > 65535) return MGB_ENOTSUP;
    const int threads = 32 + c.NB * P.tpo;
    const bool scale  = a.scale != 1.0;
    P.mask            = a.mask;
    const bool masked = a.mask.off != nullptr;
#define MGB_J(LAP)                                                                     \
    (masked ? (scale ? launch_jacobi<LAP, true, true>(P, c, grid, threads, smem, st)   \
                     : launch_jacobi<LAP, false, true>(P, c, grid, threads, smem, st)) \
            : (scale ? launch_jacobi<LAP, true, false>(P, c, grid, threads, smem, st)  \
                     : launch_jacobi<LAP, false, false>(P, c, grid, threads, smem, st)))
    if (a.lap_type == MGB_LAP_4) return MGB_J(kLap4);
    if (a.lap_type == MGB_LAP_2) return MGB_J(kLap2);
    return MGB_J(kLapMehr);
#undef MGB_J
}

// launch shape for the (k-vector, row, plane, function) kernels
struct VecLaunch
{
    dim3 block, grid;
};
static VecLaunch vec_launch(int nx, int ny, int nzv, int nf)
{
    VecLaunch L;
    int bx = 1;
    while (bx < nzv && bx < 64) bx *= 2;
    const int by = 256 / bx;
    L.block      = dim3(bx, by, 1);
    L.grid       = dim3((unsigned)(((nzv + bx - 1) / bx) * ((ny + by - 1) / by)), (unsigned)nx,
        (unsigned)nf);
    return L;
}

static MaskView mask_from(const MaskView& m, int f0)
{
    MaskView v = m;
    if (v.off) v.off += (size_t)f0 * v.subdivx;
    return v;
}

// nb3d (27 blocks, index ((dx+1)*3 + (dy+1))*3 + (dz+1)) -> the 8 a transfer kernel touches
static Nb8 nb8_of(const float* const* nb3d, int side)
{
    Nb8 n;
    for (int b = 0; b < 8; b++)
    {
        const int dx = (b & 4) ? side : 0, dy = (b & 2) ? side : 0, dz = (b & 1) ? side : 0;
        n.p[b]       = nb3d[((dx + 1) * 3 + (dy + 1)) * 3 + (dz + 1)];
    }
    return n;
}

int mg_restrict(const mgb_grid& fine, const float* w, size_t ldf, float* coarse, size_t ldc,
    int nfunc, const MaskView& mask, const float* w_west, const int* map_w, cudaStream_t st,
    const float* const* nb3d)
{
    const int nxc = fine.dim[0] / 2, nyc = fine.dim[1] / 2, nzc = fine.dim[2] / 2;
    for (int f0 = 0; nb3d && f0 < nfunc; f0 += 65535)
    {
        const int nf = (nfunc - f0 < 65535) ? nfunc - f0 : 65535;
        const VecLaunch L = vec_launch(nxc, nyc, nzc / 4, nf);
        k_mg_restrict3<<<L.grid, L.block, 0, st>>>(nxc, nyc, nzc,
            fine.bc[0] != 1 && fine.coord[0] == 0, fine.bc[1] != 1 && fine.coord[1] == 0,
            fine.bc[2] != 1 && fine.coord[2] == 0, nb8_of(nb3d, -1), (long long)ldf,
            coarse + (size_t)f0 * ldc, (long long)ldc, mask_from(mask, f0), f0);
        MGB_LAUNCHED("k_mg_restrict3");
    }
    if (nb3d) return MGB_OK;
    for (int f0 = 0; f0 < nfunc; f0 += 65535)
    {
        const int nf = (nfunc - f0 < 65535) ? nfunc - f0 : 65535;
        const VecLaunch L = vec_launch(nxc, nyc, nzc / 4, nf);
        k_mg_restrict<<<L.grid, L.block, 0, st>>>(nxc, nyc, nzc, fine.bc[0] == 1, fine.bc[1] == 1,
            fine.bc[2] == 1, w + (size_t)f0 * ldf, (long long)ldf,
            coarse + (size_t)f0 * ldc, (long long)ldc, mask_from(mask, f0),
            w_west, map_w ? map_w + f0 : nullptr, f0);
        MGB_LAUNCHED("k_mg_restrict");
    }
    return MGB_OK;
}

int mg_prolong_correct(const mgb_grid& fine, const float* coarse, size_t ldc, float* v,
    size_t ldv, int nfunc, const int zero_low[3], const MaskView& mask, const float* coarse_east,
    const int* map_e, cudaStream_t st, const float* const* nb3d)
{
    const int nx = fine.dim[0], ny = fine.dim[1], nz = fine.dim[2];
    Nb8 none;
    memset(&none, 0, sizeof(none));
    for (int f0 = 0; f0 < nfunc; f0 += 65535)
    {
        const int nf = (nfunc - f0 < 65535) ? nfunc - f0 : 65535;
        const VecLaunch L = vec_launch(nx / 2, ny / 2, nz / 4, nf);
        if (nb3d)
        {
            // "periodic" = the high neighbour exists (not a Dirichlet end of the domain)
            int cont[3];
            for (int d = 0; d < 3; d++)
                cont[d] = !(fine.bc[d] != 1 && fine.coord[d] == fine.nproc[d] - 1);
            k_mg_prolong_correct<true><<<L.grid, L.block, 0, st>>>(nx, ny, nz, cont[0], cont[1],
                cont[2], zero_low[0], zero_low[1], zero_low[2], coarse + (size_t)f0 * ldc,
                (long long)ldc, v + (size_t)f0 * ldv, (long long)ldv, mask_from(mask, f0),
                nullptr, nullptr, f0, nb8_of(nb3d, +1));
        }
        else
            k_mg_prolong_correct<false><<<L.grid, L.block, 0, st>>>(nx, ny, nz, fine.bc[0] == 1,
                fine.bc[1] == 1, fine.bc[2] == 1, zero_low[0], zero_low[1], zero_low[2],
                coarse + (size_t)f0 * ldc, (long long)ldc, v + (size_t)f0 * ldv, (long long)ldv,
                mask_from(mask, f0), coarse_east, map_e ? map_e + f0 : nullptr, f0, none);
        MGB_LAUNCHED("k_mg_prolong_correct");
    }
    return MGB_OK;
}

int mg_convert(size_t npt, const double* in, size_t ldi, float* out, size_t ldo, int nfunc,
    cudaStream_t st)
{
    for (int f0 = 0; f0 < nfunc; f0 += 65535)
    {
        const int nf = (nfunc - f0 < 65535) ? nfunc - f0 : 65535;
        dim3 grid((unsigned)((npt / 4 + 255) / 256), (unsigned)nf);
        k_mg_convert<<<grid, 256, 0, st>>>((long long)npt, in + (size_t)f0 * ldi,
            (long long)ldi, out + (size_t)f0 * ldo, (long long)ldo);
        MGB_LAUNCHED("k_mg_convert");
    }
    return MGB_OK;
}

int mg_scale(const mgb_grid& gr, double c, const float* f, size_t ldf, float* v, size_t ldv,
    int nfunc, const int zero_low[3], const MaskView& mask, cudaStream_t st)
{
    const int nx = gr.dim[0], ny = gr.dim[1], nz = gr.dim[2];
    float ch, cl;
    split_double(c, ch, cl);
    for (int f0 = 0; f0 < nfunc; f0 += 65535)
    {
        const int nf = (nfunc - f0 < 65535) ? nfunc - f0 : 65535;
        const VecLaunch L = vec_launch(nx, ny, nz / 4, nf);
        k_mg_scale<<<L.grid, L.block, 0, st>>>(nx, ny, nz, zero_low[0], zero_low[1], zero_low[2],
            ch, cl, f + (size_t)f0 * ldf, (long long)ldf, v + (size_t)f0 * ldv,
            (long long)ldv, mask_from(mask, f0));
        MGB_LAUNCHED("k_mg_scale");
    }
    return MGB_OK;
}

} // namespace mgb
