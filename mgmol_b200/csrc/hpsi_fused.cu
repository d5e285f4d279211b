// Fused local Hamiltonian, TMA-pipelined x-streaming kernel (path 1 of
// mgb_hpsi).  One pass over phi: read each orbital value once from HBM, write
// H phi once.  Algorithmic traffic 2*sizeof(ORBDTYPE) bytes per grid-point *
// orbital (+ 8/N for the shared potential), HBM-bound.
//
//   hphi = A phi + B (V .* phi)     Mehrstellen, 19-pt A + 7-pt B   (G = 1)
//   hphi = A phi + V .* phi         classical 4th order, 13-pt A    (G = 2)
// (Hamiltonian<T>::applyLocal, src/Hamiltonian.cc:85-159; stencils
//  src/pb/FDkernels.cc:447-584, src/pb/FDoper.cc:321-399.)
//
// Decomposition of the work
//   grid.x : tile of TY = RY*YG rows in y (full rows in z, nz <= 256)
//   grid.y : block of NB orbitals that share one staged copy of the potential
//   grid.z : chunk of XC planes in x
// A CTA streams the y-z planes of its chunk along x through an S-stage ring of
// shared-memory buffers.  Warp 0 is the TMA producer: per plane it issues three
// boxes per field (G halo rows below, TY rows, G halo rows above; the halo
// boxes carry the periodic wrap in y or fall outside the tensor and are
// zero-filled for Dirichlet), completing on the stage's "full" mbarrier.  The
// x halo is just another plane: wrapped coordinate, zero plane, or a plane of
// the neighbour rank's halo buffer.  Consumer threads each own VEC contiguous z
// points (16 bytes) of RY rows; z neighbours come from shared memory with the
// wrap folded into the index.  The x direction never touches shared memory:
// each plane contributes a(plane) to its own output and b(plane) to the two
// neighbouring outputs, so only two (four for G=2) running sums per point live
// in registers.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "hpsi.h"
#include "tma_ptx.cuh"

namespace mgb
{

// explicit shared-space accesses with 32-bit addresses
__device__ __forceinline__ void lds_vec(uint32_t a, double (&o)[2])
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "r"(a));
}
__device__ __forceinline__ void lds_vec(uint32_t a, float (&o)[4])
{
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3])
                 : "r"(a));
}
__device__ __forceinline__ void lds_pair(uint32_t a, double& x, double& y)
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
__device__ __forceinline__ void lds_pair(uint32_t a, float& x, float& y)
{
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(a));
}
__device__ __forceinline__ double lds_one_f64(uint32_t a)
{
    double x;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(a));
    return x;
}
__device__ __forceinline__ float lds_one_f32(uint32_t a)
{
    float x;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a));
    return x;
}
template <typename T>
__device__ __forceinline__ T lds_one(uint32_t a)
{
    if constexpr (sizeof(T) == 8)
        return lds_one_f64(a);
    else
        return lds_one_f32(a);
}

// ---------------------------------------------------------------------------
// Kernel parameters
// ---------------------------------------------------------------------------
// coefficient slots
enum
{
    kC0 = 0, kCx, kCy, kCz, kCxy, kCyz, kCxz, // Mehrstellen A (FDkernels.cc:457-463)
    kD0, kD1x, kD2x, kD1y, kD2y, kD1z, kD2z,  // 4th order (FDoper.cc:331-343)
    kNumCoef
};

struct alignas(64) FusedParams
{
    CUtensorMap psi_mid, psi_halo;   // phi  (z, y, x, orbital)
    CUtensorMap v_mid, v_halo;       // vtot (z, y, x)
    CUtensorMap xpsi_mid, xpsi_halo; // x-halo buffer of phi  (z, y, 2g, orbital)
    CUtensorMap xv_mid, xv_halo;     // x-halo buffer of vtot (z, y, 2g)
    CUtensorMap pw_mid, pw_halo;     // west neighbour's phi block (peer memory)
    CUtensorMap pe_mid, pe_halo;     // east neighbour's phi block (peer memory)
    int peer;                        // x-halo planes of phi come from pw / pe
    const int* map_w;                // color slot on that rank (-1: absent), or null
    const int* map_e;
    void* out;
    long long ldh;
    int nx, ny, nz, nfunc;
    int TY, YG, NB, S, XC;
    int nzv, tpo;
    int row_bytes, off_mid, off_hi, tile_bytes, stage_bytes;
    int split_x, first_x, last_x;
    // ---- PEER == 2: any px x py x pz decomposition, every halo read in place --------
    // nbmaps[((dx+1)*3 + (dy+1))*3 + (dz+1)][kind]: tensor maps over the block of the
    // rank at coord + (dx,dy,dz) (myself where a direction is not split: the periodic
    // wrap); kind 0: box {nz, TY}, 1: {nz, G}, 2: {ZH, TY}, 3: {ZH, G} (ZH = 16 bytes
    // of z: the z-halo columns).  Device memory, 128-byte aligned.
    const CUtensorMap* nbmaps;
    CUtensorMap vg_map; // Mehrstellen: padded ghosted V in the compute type, box {nzp / nvb, TY+2G}
    int zsplit;         // z halo columns come from boxes (z is split), not from the wrapped index
    int first_y, last_y, first_z, last_z; // non-periodic AND at the domain's end
    int per_y, per_z;
    int zarr_mid, zarr_hi, zarr_bytes, zlo_off; // z-halo entries inside a tile
    int zstride, zhi_off; // 32, 16: pushed buffers (one 32-byte entry per row); 16, zarr_bytes: in place (two arrays)
    int vrow_bytes, nvb;                        // padded V tile: row pitch, boxes per plane
    int pol;      // L2 policy of the phi loads (see the producer)
    int orb_fast; // blockIdx.x walks the orbital blocks (CTAs of a wave share V tiles in L2)
    double cd[kNumCoef];
    float cf[kNumCoef];
};

constexpr int kBarBytes = 1024;

template <typename T>
__device__ __forceinline__ T coef(const FusedParams& P, int i)
{
    if constexpr (sizeof(T) == 8)
        return P.cd[i];
    else
        return P.cf[i];
}

// product rounded on its own (never contracted into a following add): the
// reference rounds V*psi before B is applied (GridFuncVector.cc:131), and a
// contraction that depends on the lane's position inside its vector would
// break bit-exact translation covariance
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }

template <typename T, int VEC>
__device__ __forceinline__ void store_vec_streaming(T* p, const T (&v)[VEC])
{
    if constexpr (VEC == 2)
        __stcs(reinterpret_cast<double2*>(p), make_double2(v[0], v[1]));
    else
        __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
}

// where does plane p (local x index, may be outside [0,nx)) come from?
// 0: all zeros, 1: main tensor at x = xc, 2: x-halo buffer at plane xc
template <bool PERIODIC, int G>
__device__ __forceinline__ int plane_source(const FusedParams& P, int p, int& xc)
{
    if (p >= 0 && p < P.nx)
    {
        if (!PERIODIC && p == 0 && P.first_x) return 0; // zeroed first layer
        xc = p;
        return 1;
    }
    if (P.split_x)
    {
        if (!PERIODIC && ((p < 0 && P.first_x) || (p >= P.nx && P.last_x)))
            return 0;
        xc = (p < 0) ? p + G : G + (p - P.nx);
        return 2;
    }
    if (PERIODIC)
    {
        xc = (p < 0) ? p + P.nx : p - P.nx;
        return 1;
    }
    return 0;
}

// PEER: the x-halo planes of phi are read from the neighbours' blocks (a separate
// instantiation, so that the single-rank kernel carries none of that code)
template <typename T, int RY, bool PERIODIC, bool LAP4, int MAXT, int PEER>
__global__ void __launch_bounds__(MAXT, 1)
    k_hpsi_tma(const __grid_constant__ FusedParams P)
{
    constexpr int G   = LAP4 ? 2 : 1;
    constexpr int VEC = 16 / (int)sizeof(T);
    using CT          = T; // compute type

    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full  = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 8;

    const int tid     = threadIdx.x;
    const int y0      = (P.orb_fast ? blockIdx.y : blockIdx.x) * P.TY;
    const int orb0    = (P.orb_fast ? blockIdx.x : blockIdx.y) * P.NB;
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
        // ------------------------- TMA producer --------------------------
        // (PEER == 2: up to 3 NB lanes issue, one per orbital and box)
        if (PEER != 2 && tid != 0) return;
        if (PEER == 2 && tid >= 3 * P.NB) return;
        unsigned char* stages     = smem + kBarBytes;
        // phi is streamed once, but the rows at a tile's edge are also the halo
        // rows of the neighbouring y tile: pol 1 keeps phi at normal priority, pol 2
        // additionally pins the (small) halo boxes until the neighbour's load
        const uint64_t pol_keep   = policy_evict_last();
        const uint64_t pol_stream = P.pol == 0 ? policy_evict_first() : policy_evict_normal();
        const uint64_t pol_halo   = P.pol == 2 ? pol_keep : pol_stream;
        int norb = P.nfunc - orb0;
        if (norb > P.NB) norb = P.NB;
        int ylo = y0 - G, yhi = y0 + P.TY;
        if (PERIODIC)
        {
            if (ylo < 0) ylo += P.ny;
            if (yhi >= P.ny) yhi -= P.ny;
        }
        const uint32_t bytes_full_tile = (uint32_t)((P.TY + 2 * G) * P.row_bytes);
        const uint32_t bytes_v_tile
            = LAP4 ? (uint32_t)(P.TY * P.row_bytes) : bytes_full_tile;
        const uint32_t tx = bytes_v_tile + (uint32_t)norb * bytes_full_tile;
        int stage = 0;
        uint32_t par = 0;
        for (int it = 0; it < nplanes; it++)
        {
            const int p = xb - G + it;
            mbar_wait(&empty[stage], par ^ 1u);
            int xc        = 0;
            const int src = plane_source<PERIODIC, G>(P, p, xc);
            if (src == 0)
            {
                if (tid == 0) mbar_arrive(&full[stage]);
            }
            else if constexpr (PEER == 2)
            {
                // which rank's block holds plane p, and at which x coordinate
                const int dx  = (p < 0) ? 0 : (p >= P.nx ? 2 : 1);
                const int xcp = (p < 0) ? P.nx + p : (p >= P.nx ? p - P.nx : p);
                const uint32_t rows   = (uint32_t)(P.TY + 2 * G);
                // a cross stencil (4th order) never looks at the z neighbours of halo rows
                const uint32_t psi_tx = rows * (uint32_t)P.row_bytes
                                        + (P.zsplit ? (LAP4 ? (uint32_t)P.TY : rows) * 32u : 0u);
                const uint32_t v_tx   = LAP4 ? (uint32_t)(P.TY * P.row_bytes) : rows * (uint32_t)P.vrow_bytes;
                unsigned char* sb = stages + (size_t)stage * P.stage_bytes;
                if (tid == 0)
                {
                    mbar_arrive_expect_tx(&full[stage], v_tx + (uint32_t)norb * psi_tx);
                    if constexpr (LAP4)
                        tma_load_3d(sb + P.off_mid, &P.v_mid, &full[stage], 0, y0, p, pol_keep);
                    else
                        // padded ghosted V (a row = nvb chunks of <= 256 elements, so the box
                        // lands as whole rows): plane p + G, rows y0 .. y0+TY+2G-1 (ghosted)
                        tma_load_4d(sb, &P.vg_map, &full[stage], 0, 0, y0, p + G, pol_keep);
                }
                // one lane per (orbital, box): the rows below the tile, the tile's rows, the
                // rows above -- the y neighbour's last / first rows (the wrap when y is not
                // split), or out of range (zero fill) at a Dirichlet end -- each with the
                // 32-byte z-halo entries of its rows
                if (tid < 3 * norb)
                {
                    const int o = tid / 3, b = tid - 3 * o;
                    unsigned char* tb = sb + (size_t)(1 + o) * P.tile_bytes;
                    int dyi = 1, yc = y0, to = P.off_mid, zo = P.zarr_mid;
                    if (b == 0)
                    {
                        dyi = (y0 - G < 0 && !P.first_y) ? 0 : 1;
                        yc  = (y0 - G < 0) ? (P.first_y ? -G : P.ny - G) : y0 - G;
                        to = 0, zo = 0;
                    }
                    else if (b == 2)
                    {
                        dyi = (y0 + P.TY >= P.ny && !P.last_y) ? 2 : 1;
                        yc  = (y0 + P.TY >= P.ny) ? (P.last_y ? P.ny : 0) : y0 + P.TY;
                        to = P.off_hi, zo = P.zarr_hi;
                    }
                    const CUtensorMap* m = P.nbmaps + dx * 36 + dyi * 12;
                    const int kind       = (b == 1) ? 0 : 1;
                    tma_load_4d(tb + to, m + 4 + kind, &full[stage], 0, yc, xcp, orb0 + o,
                        b == 1 ? pol_stream : pol_halo);
                    if (P.zsplit && (!LAP4 || b == 1))
                    {
                        if (P.zstride == 32)
                            tma_load_4d(tb + P.zlo_off + zo, m + 2 + kind, &full[stage], 0, yc, xcp,
                                orb0 + o, pol_halo);
                        else
                        {
                            // in place: the z neighbours' last / first 16 bytes of these rows
                            // (beyond a Dirichlet end: out of range, zero fill)
                            constexpr int ZH = VEC;
                            tma_load_4d(tb + P.zlo_off + zo, m + 2 + kind, &full[stage],
                                P.first_z ? -ZH : P.nz - ZH, yc, xcp, orb0 + o, pol_halo);
                            tma_load_4d(tb + P.zlo_off + P.zarr_bytes + zo, m + 8 + 2 + kind,
                                &full[stage], P.last_z ? P.nz : 0, yc, xcp, orb0 + o, pol_halo);
                        }
                    }
                }
            }
            else
            {
                mbar_arrive_expect_tx(&full[stage], tx);
                unsigned char* sb = stages + (size_t)stage * P.stage_bytes;
                const CUtensorMap* vm = (src == 1) ? &P.v_mid : &P.xv_mid;
                const CUtensorMap* vh = (src == 1) ? &P.v_halo : &P.xv_halo;
                const CUtensorMap* pm = (src == 1) ? &P.psi_mid : &P.xpsi_mid;
                const CUtensorMap* ph = (src == 1) ? &P.psi_halo : &P.xpsi_halo;
                int xcp = xc;
                const int* cmap = nullptr;
                if (PEER == 1 && src == 2)
                {
                    cmap = (p < 0) ? P.map_w : P.map_e;
                    // read the neighbour's boundary plane straight from its
                    // block over NVLink: last planes of the west rank, first
                    // planes of the east rank
                    pm  = (p < 0) ? &P.pw_mid : &P.pe_mid;
                    ph  = (p < 0) ? &P.pw_halo : &P.pe_halo;
                    xcp = (p < 0) ? P.nx + p : p - P.nx;
                }
                // potential tile, shared by the NB orbitals of this CTA
                tma_load_3d(sb + P.off_mid, vm, &full[stage], 0, y0, xc, pol_keep);
                if (!LAP4)
                {
                    tma_load_3d(sb, vh, &full[stage], 0, ylo, xc, pol_keep);
                    tma_load_3d(sb + P.off_hi, vh, &full[stage], 0, yhi, xc, pol_keep);
                }
                if (PEER != 1 || !cmap)
                {
                    // the issue loop of the producer thread: kept minimal
                    for (int o = 0; o < norb; o++)
                    {
                        unsigned char* tb = sb + (size_t)(1 + o) * P.tile_bytes;
                        tma_load_4d(tb + P.off_mid, pm, &full[stage], 0, y0, xcp, orb0 + o,
                            pol_stream);
                        tma_load_4d(tb, ph, &full[stage], 0, ylo, xcp, orb0 + o, pol_halo);
                        tma_load_4d(tb + P.off_hi, ph, &full[stage], 0, yhi, xcp, orb0 + o,
                            pol_halo);
                    }
                }
                else
                {
                    // the orbital may sit in another color slot on the
                    // neighbour (gid-addressed packets, src/pb/GridFuncVector.cc:
                    // 1225-1246,1374-1393); absent there: an out-of-range
                    // function index makes TMA deliver zeros
                    for (int o = 0; o < norb; o++)
                    {
                        unsigned char* tb = sb + (size_t)(1 + o) * P.tile_bytes;
                        int fo = cmap[orb0 + o];
                        if (fo < 0) fo = P.nfunc;
                        tma_load_4d(tb + P.off_mid, pm, &full[stage], 0, y0, xcp, fo, pol_stream);
                        tma_load_4d(tb, ph, &full[stage], 0, ylo, xcp, fo, pol_halo);
                        tma_load_4d(tb + P.off_hi, ph, &full[stage], 0, yhi, xcp, fo, pol_halo);
                    }
                }
            }
            if (++stage == S)
            {
                stage = 0;
                par ^= 1u;
            }
        }
        return;
    }

    // ------------------------------ consumers ------------------------------
    const int ct   = tid - 32;
    const int lane = tid & 31;
    const int grp  = ct / P.tpo;
    const int lt   = ct - grp * P.tpo;
    const int orb  = orb0 + grp;
    const bool active = (lt < P.YG * P.nzv) && (orb < P.nfunc);
    const int zv   = active ? lt % P.nzv : 0;
    const int yg   = active ? lt / P.nzv : 0;
    const int z0   = zv * VEC;
    const int rr0  = yg * RY;
    const int rb   = P.row_bytes;

    // Row offsets inside a tile.  Rows rr0-G .. rr0+RY+G-1 are walked in
    // order; the first G may sit in the low-halo box (only when rr0 == 0) and
    // the last G in the high-halo box (only when rr0+RY == TY).
    const int mid0 = P.off_mid + (rr0 - G) * rb;
    const uint32_t startA = (rr0 == 0) ? 0u : (uint32_t)mid0;
    const uint32_t startB = (uint32_t)(mid0 + G * rb);
    const uint32_t startC = (rr0 + RY == P.TY) ? (uint32_t)P.off_hi
                                               : (uint32_t)(mid0 + (RY + G) * rb);
    // z neighbours (byte offsets inside the row) and Dirichlet masks
    int zl = LAP4 ? z0 - 2 : z0 - 1;
    int zr = z0 + VEC;
    CT ml = (CT)1, mr = (CT)1, m0 = (CT)1;
    if (zl < 0)
    {
        zl += P.nz;
        if (!PERIODIC) ml = (CT)0;
    }
    if (zr >= P.nz)
    {
        zr -= P.nz;
        if (!PERIODIC) mr = (CT)0;
    }
    if (!PERIODIC && z0 == 0) m0 = (CT)0; // zeroed first z layer
    const uint32_t zoff  = (uint32_t)(z0 * (int)sizeof(T));
    const uint32_t zloff = (uint32_t)(zl * (int)sizeof(T));
    const uint32_t zroff = (uint32_t)(zr * (int)sizeof(T));
    // tile-local row index of the zeroed first y layer (Dirichlet), as a
    // position i in this thread's row walk; -100 when it is not in the walk
    int izero = (!PERIODIC) ? (0 - (y0 + rr0 - G)) : -100;
    // PEER == 2 with z split: the G columns left of z = 0 / right of z = nz-1 sit in the
    // tile's z-halo arrays (16 bytes per row: the neighbour's last / first columns)
    bool isL = false, isR = false;
    uint32_t zsA = 0, zsB = 0, zsC = 0, zLb = 0, zRb = 0;
    if constexpr (PEER == 2)
    {
        if (P.zsplit)
        {
            isL = z0 == 0;
            isR = z0 + VEC == P.nz;
            ml = mr = (CT)1; // a Dirichlet end delivers zero columns (out-of-range box)
        }
        if (!PERIODIC)
        {
            if (!P.first_z) m0 = (CT)1;
            if (!P.first_y) izero = -100;
        }
        zsA = (rr0 == 0) ? 0u : (uint32_t)(P.zarr_mid + (rr0 - G) * P.zstride);
        zsB = (uint32_t)(P.zarr_mid + rr0 * P.zstride);
        zsC = (rr0 + RY == P.TY) ? (uint32_t)P.zarr_hi
                                 : (uint32_t)(P.zarr_mid + (rr0 + RY) * P.zstride);
        // the last G (1: Mehrstellen, 2: 4th order) elements of the low array's row, the
        // first of the high array's
        zLb = (uint32_t)(P.zlo_off + 16 - G * (int)sizeof(T));
        zRb = (uint32_t)(P.zlo_off + P.zhi_off);
    }

    const uint32_t stage0 = smem_u32(smem + kBarBytes);
    const uint32_t grp_off = (uint32_t)((1 + grp) * P.tile_bytes);

    T* outp = reinterpret_cast<T*>(P.out) + (long long)orb * P.ldh
              + (long long)(y0 + rr0) * P.nz + z0;
    const long long plane_elems = (long long)P.ny * P.nz;

    int stage    = 0;
    uint32_t par = 0;

    if constexpr (!LAP4)
    {
        // =================== Mehrstellen, G = 1 ===========================
        const CT c0 = coef<CT>(P, kC0), cx = coef<CT>(P, kCx), cy = coef<CT>(P, kCy),
                 cz = coef<CT>(P, kCz), cxy = coef<CT>(P, kCxy),
                 cyz = coef<CT>(P, kCyz), cxz = coef<CT>(P, kCxz);
        const CT half = (CT)0.5, c12 = (CT)(1. / 12.);
        CT acc[RY][VEC], bprev[RY][VEC];
#pragma unroll
        for (int r = 0; r < RY; r++)
#pragma unroll
            for (int e = 0; e < VEC; e++)
            {
                acc[r][e]   = (CT)0;
                bprev[r][e] = (CT)0;
            }

        for (int it = 0; it < nplanes; it++)
        {
            const int p   = xb - G + it;
            int xc        = 0;
            const int src = plane_source<PERIODIC, G>(P, p, xc);
            mbar_wait(&full[stage], par);
            if (active)
            {
                T* optr = outp + (long long)(p - 1) * plane_elems;
                const bool do_store = (it >= 2);
                if (src != 0)
                {
                    const uint32_t vb = stage0 + (uint32_t)stage * (uint32_t)P.stage_bytes;
                    const uint32_t pb = vb + grp_off;
                    CT Pm[VEC], szm[VEC], wm[VEC];
                    CT Pc[VEC], szc[VEC], wc[VEC], wzc[VEC];
                    CT Pp[VEC], szp[VEC], wp[VEC], wzp[VEC];
                    // psi, z-sum of psi, w = V*psi, z-sum of w of one row
                    auto load_row = [&](uint32_t ro, uint32_t zo, int i, CT(&Pq)[VEC],
                                        CT(&szq)[VEC], CT(&wq)[VEC], CT(&wzq)[VEC]) {
                        CT Vq[VEC];
                        lds_vec(pb + ro + zoff, Pq);
                        uint32_t aL = pb + ro + zloff, aR = pb + ro + zroff;
                        uint32_t aV = vb + ro + zoff, aVL = vb + ro + zloff, aVR = vb + ro + zroff;
                        if constexpr (PEER == 2)
                        {
                            if (isL) aL = pb + zLb + zo;
                            if (isR) aR = pb + zRb + zo;
                            // padded ghosted V tile: row rr0 + i, element z + VEC
                            aV  = vb + (uint32_t)((rr0 + i) * P.vrow_bytes) + zoff + 16u;
                            aVL = aV - (uint32_t)sizeof(T);
                            aVR = aV + 16u;
                        }
                        lds_vec(aV, Vq);
                        CT PL = lds_one<CT>(aL);
                        CT PR = lds_one<CT>(aR);
                        const CT VL = lds_one<CT>(aVL);
                        const CT VR = lds_one<CT>(aVR);
                        if (!PERIODIC)
                        {
                            PL *= ml;
                            PR *= mr;
                            Pq[0] *= m0;
                            if (i == izero)
                            {
#pragma unroll
                                for (int e = 0; e < VEC; e++)
                                    Pq[e] = (CT)0;
                                PL = (CT)0;
                                PR = (CT)0;
                            }
                        }
                        const CT wL = mul_rn(VL, PL), wR = mul_rn(VR, PR);
#pragma unroll
                        for (int e = 0; e < VEC; e++)
                            wq[e] = mul_rn(Vq[e], Pq[e]);
#pragma unroll
                        for (int e = 0; e < VEC; e++)
                        {
                            const CT left  = (e == 0) ? PL : Pq[e - 1];
                            const CT right = (e == VEC - 1) ? PR : Pq[e + 1];
                            szq[e]         = left + right;
                            const CT wl    = (e == 0) ? wL : wq[e - 1];
                            const CT wr    = (e == VEC - 1) ? wR : wq[e + 1];
                            wzq[e]         = wl + wr;
                        }
                    };
                    CT dummy[VEC];
                    load_row(startA, zsA, 0, Pm, szm, wm, dummy);
                    load_row(startB, zsB, 1, Pc, szc, wc, wzc);
                    uint32_t ro = startB, zo = zsB;
#pragma unroll
                    for (int r = 0; r < RY; r++)
                    {
                        ro = (r == RY - 1) ? startC : ro + (uint32_t)rb;
                        zo = (r == RY - 1) ? zsC : zo + (uint32_t)P.zstride;
                        load_row(ro, zo, r + 2, Pp, szp, wp, wzp);
                        CT o[VEC];
#pragma unroll
                        for (int e = 0; e < VEC; e++)
                        {
                            const CT py  = Pm[e] + Pp[e];
                            const CT szy = szm[e] + szp[e];
                            const CT ws  = wzc[e] + (wm[e] + wp[e]);
                            CT a         = c0 * Pc[e];
                            a            = fma(cz, szc[e], a);
                            a            = fma(cy, py, a);
                            a            = fma(cyz, szy, a);
                            a            = fma(half, wc[e], a);
                            a            = fma(c12, ws, a);
                            CT b         = cx * Pc[e];
                            b            = fma(cxz, szc[e], b);
                            b            = fma(cxy, py, b);
                            b            = fma(c12, wc[e], b);
                            o[e]         = acc[r][e] + b;
                            acc[r][e]    = a + bprev[r][e];
                            bprev[r][e]  = b;
                        }
                        if (do_store) store_vec_streaming<T, VEC>(optr, o);
                        optr += P.nz;
#pragma unroll
                        for (int e = 0; e < VEC; e++)
                        {
                            Pm[e]  = Pc[e];
                            szm[e] = szc[e];
                            wm[e]  = wc[e];
                            Pc[e]  = Pp[e];
                            szc[e] = szp[e];
                            wc[e]  = wp[e];
                            wzc[e] = wzp[e];
                        }
                    }
                }
                else
                {
#pragma unroll
                    for (int r = 0; r < RY; r++)
                    {
                        CT o[VEC];
#pragma unroll
                        for (int e = 0; e < VEC; e++)
                        {
                            o[e]        = acc[r][e];
                            acc[r][e]   = bprev[r][e];
                            bprev[r][e] = (CT)0;
                        }
                        if (do_store) store_vec_streaming<T, VEC>(optr, o);
                        optr += P.nz;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == S)
            {
                stage = 0;
                par ^= 1u;
            }
        }
    }
    else
    {
        // =================== classical 4th order, G = 2 ====================
        const CT d0 = coef<CT>(P, kD0), d1x = coef<CT>(P, kD1x),
                 d2x = coef<CT>(P, kD2x), d1y = coef<CT>(P, kD1y),
                 d2y = coef<CT>(P, kD2y), d1z = coef<CT>(P, kD1z),
                 d2z = coef<CT>(P, kD2z);
        // running sums for outputs p-2 (s2), p-1 (s1), p (s0), p+1 (sp)
        CT s2[RY][VEC], s1[RY][VEC], s0[RY][VEC], sp[RY][VEC];
#pragma unroll
        for (int r = 0; r < RY; r++)
#pragma unroll
            for (int e = 0; e < VEC; e++)
                s2[r][e] = s1[r][e] = s0[r][e] = sp[r][e] = (CT)0;

        for (int it = 0; it < nplanes; it++)
        {
            const int p   = xb - G + it;
            int xc        = 0;
            const int src = plane_source<PERIODIC, G>(P, p, xc);
            mbar_wait(&full[stage], par);
            if (active)
            {
                T* optr = outp + (long long)(p - 2) * plane_elems;
                const bool do_store = (it >= 4);
                if (src != 0)
                {
                    const uint32_t vb = stage0 + (uint32_t)stage * (uint32_t)P.stage_bytes;
                    const uint32_t pb = vb + grp_off;
                    // sliding window of centre vectors: rows r-2 .. r+2
                    CT W[5][VEC];
                    auto load_center = [&](uint32_t ro, int i, CT(&Pq)[VEC]) {
                        lds_vec(pb + ro + zoff, Pq);
                        if (!PERIODIC)
                        {
                            Pq[0] *= m0;
                            if (i == izero)
                            {
#pragma unroll
                                for (int e = 0; e < VEC; e++)
                                    Pq[e] = (CT)0;
                            }
                        }
                    };
                    load_center(startA, 0, W[0]);
                    load_center(startA + (uint32_t)rb, 1, W[1]);
                    load_center(startB, 2, W[2]);
                    uint32_t ro_c = startB;            // centre row of output r
                    uint32_t ro_n = startB;            // newest loaded row
                    if (RY > 1)
                    {
                        ro_n = startB + (uint32_t)rb;
                        load_center(ro_n, 3, W[3]);
                    }
                    else
                    {
                        ro_n = startC;
                        load_center(ro_n, 3, W[3]);
                    }
#pragma unroll
                    for (int r = 0; r < RY; r++)
                    {
                        // row index r+4 of the walk: mid rows up to RY+G-1,
                        // then the high-halo rows
                        const int inew = r + 4;
                        if (inew < RY + G)
                            ro_n += (uint32_t)rb;
                        else if (inew == RY + G)
                            ro_n = startC;
                        else
                            ro_n += (uint32_t)rb;
                        load_center(ro_n, inew, W[4]);
                        // z neighbours of the centre row
                        CT L2, L1, R1, R2; // z-2, z-1, z+VEC, z+VEC+1
                        uint32_t aL = pb + ro_c + zloff, aR = pb + ro_c + zroff;
                        if constexpr (PEER == 2)
                        {
                            if (isL) aL = pb + zLb + zsB + (uint32_t)(r * P.zstride);
                            if (isR) aR = pb + zRb + zsB + (uint32_t)(r * P.zstride);
                        }
                        lds_pair(aL, L2, L1);
                        lds_pair(aR, R1, R2);
                        if (!PERIODIC)
                        {
                            L2 *= ml;
                            L1 *= ml;
                            R1 *= mr;
                            R2 *= mr;
                            // a left pair starting at z = 0 holds the zeroed
                            // first layer in its first element
                            if (z0 == 2 && (PEER != 2 || P.first_z)) L2 = (CT)0;
                            if (r + 2 == izero) L2 = L1 = R1 = R2 = (CT)0;
                        }
                        CT Vq[VEC];
                        lds_vec(vb + ro_c + zoff, Vq);
                        CT o[VEC];
#pragma unroll
                        for (int e = 0; e < VEC; e++)
                        {
                            const CT c   = W[2][e];
                            const CT zm1 = (e == 0) ? L1 : W[2][e > 0 ? e - 1 : 0];
                            const CT zm2 = (e == 0)   ? L2
                                           : (e == 1) ? L1
                                                      : W[2][e > 1 ? e - 2 : 0];
                            const CT zp1
                                = (e == VEC - 1) ? R1 : W[2][e < VEC - 1 ? e + 1 : 0];
                            const CT zp2 = (e == VEC - 1)   ? R2
                                           : (e == VEC - 2) ? R1
                                                            : W[2][e < VEC - 2 ? e + 2 : 0];
                            CT inp   = (d0 + Vq[e]) * c;
                            inp      = fma(d1y, W[1][e] + W[3][e], inp);
                            inp      = fma(d2y, W[0][e] + W[4][e], inp);
                            inp      = fma(d1z, zm1 + zp1, inp);
                            inp      = fma(d2z, zm2 + zp2, inp);
                            o[e]     = fma(d2x, c, s2[r][e]);
                            s2[r][e] = fma(d1x, c, s1[r][e]);
                            s1[r][e] = s0[r][e] + inp;
                            s0[r][e] = fma(d1x, c, sp[r][e]);
                            sp[r][e] = d2x * c;
                        }
                        if (do_store) store_vec_streaming<T, VEC>(optr, o);
                        optr += P.nz;
                        ro_c += (uint32_t)rb;
#pragma unroll
                        for (int e = 0; e < VEC; e++)
                        {
                            W[0][e] = W[1][e];
                            W[1][e] = W[2][e];
                            W[2][e] = W[3][e];
                            W[3][e] = W[4][e];
                        }
                    }
                }
                else
                {
#pragma unroll
                    for (int r = 0; r < RY; r++)
                    {
                        CT o[VEC];
#pragma unroll
                        for (int e = 0; e < VEC; e++)
                        {
                            o[e]     = s2[r][e];
                            s2[r][e] = s1[r][e];
                            s1[r][e] = s0[r][e];
                            s0[r][e] = sp[r][e];
                            sp[r][e] = (CT)0;
                        }
                        if (do_store) store_vec_streaming<T, VEC>(optr, o);
                        optr += P.nz;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == S)
            {
                stage = 0;
                par ^= 1u;
            }
        }
    }
}

// float copy of the potential for ORBDTYPE float (the reference multiplies
// float psi by the double potential and rounds to float,
// GridFuncVector.cc:131; the relative difference is one float ulp)
__global__ void k_f64_to_f32(size_t n, const double* __restrict__ in,
    float* __restrict__ out)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        out[i] = (float)in[i];
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn)
    {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q)
                == cudaSuccess
            && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// tensor (z, y, x [, orbital]) with a box of `rows` full z-rows in one plane
int make_map(CUtensorMap* m, bool f64, const void* base, int rank, int nz,
    int ny, int nxp, long long ld_elems, int nfunc, int rows)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc)
    {
        set_error("cuTensorMapEncodeTiled entry point not found");
        return MGB_ECUDA;
    }
    const cuuint64_t es = f64 ? 8 : 4;
    cuuint64_t dims[4]  = { (cuuint64_t)nz, (cuuint64_t)ny, (cuuint64_t)nxp,
        (cuuint64_t)(nfunc > 0 ? nfunc : 1) };
    cuuint64_t strides[3] = { (cuuint64_t)nz * es, (cuuint64_t)nz * ny * es,
        (cuuint64_t)ld_elems * es };
    cuuint32_t box[4]  = { (cuuint32_t)nz, (cuuint32_t)rows, 1, 1 };
    cuuint32_t estr[4] = { 1, 1, 1, 1 };
    CUresult r = enc(m, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
        (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
    {
        set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %d %d %d rows %d",
            (int)r, rank, nz, ny, nxp, rows);
        return MGB_ECUDA;
    }
    return MGB_OK;
}

static char g_last_kernel[160] = "";
const char* hpsi_last_kernel() { return g_last_kernel; }

// tensor (z, y, x, orbital) with a box {bz, rows} (bz elements of z)
static int make_map_box(CUtensorMap* m, bool f64, const void* base, int nz, int ny, int nx,
    long long ld_elems, int nfunc, int bz, int rows)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc)
    {
        set_error("cuTensorMapEncodeTiled entry point not found");
        return MGB_ECUDA;
    }
    const cuuint64_t es = f64 ? 8 : 4;
    cuuint64_t dims[4]  = { (cuuint64_t)nz, (cuuint64_t)ny, (cuuint64_t)nx,
        (cuuint64_t)(nfunc > 0 ? nfunc : 1) };
    cuuint64_t strides[3] = { (cuuint64_t)nz * es, (cuuint64_t)nz * ny * es,
        (cuuint64_t)ld_elems * es };
    cuuint32_t box[4]  = { (cuuint32_t)bz, (cuuint32_t)rows, 1, 1 };
    cuuint32_t estr[4] = { 1, 1, 1, 1 };
    CUresult r = enc(m, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
        const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
    {
        set_error("cuTensorMapEncodeTiled failed (%d): box %d x %d of %d %d %d", (int)r, bz, rows,
            nz, ny, nx);
        return MGB_ECUDA;
    }
    return MGB_OK;
}

// padded ghosted potential in the compute type: (nzp = nvb chunks, ny+2G rows, nx+2G planes),
// box = whole rows {chunk, nvb, rows, 1}
static int make_map_vpad(CUtensorMap* m, bool f64, const void* base, int nzp, int nvb, int nyg,
    int nxg, int rows)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) return MGB_ECUDA;
    const cuuint64_t es = f64 ? 8 : 4;
    const int chunk     = nzp / nvb;
    cuuint64_t dims[4]  = { (cuuint64_t)chunk, (cuuint64_t)nvb, (cuuint64_t)nyg, (cuuint64_t)nxg };
    cuuint64_t strides[3] = { (cuuint64_t)chunk * es, (cuuint64_t)nzp * es,
        (cuuint64_t)nzp * nyg * es };
    cuuint32_t box[4]  = { (cuuint32_t)chunk, (cuuint32_t)nvb, (cuuint32_t)rows, 1 };
    cuuint32_t estr[4] = { 1, 1, 1, 1 };
    CUresult r = enc(m, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
        const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
    {
        set_error("cuTensorMapEncodeTiled failed (%d) for the padded potential", (int)r);
        return MGB_ECUDA;
    }
    return MGB_OK;
}

// ghosted double V (width G) -> padded compute-type V: row pitch nzp = nz + 2 ZH (ZH = 16
// bytes), element z' = z + ZH; columns no stencil reads are zero
template <typename T>
__global__ void k_vpad(int nxg, int nyg, int nz, int G, int nzp, const double* __restrict__ vg,
    T* __restrict__ out)
{
    constexpr int ZH    = 16 / (int)sizeof(T);
    const long long tot = (long long)nxg * nyg * nzp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot;
         i += (long long)gridDim.x * blockDim.x)
    {
        const int zp       = (int)(i % nzp);
        const long long rw = i / nzp;
        const int zg       = zp - ZH + G; // ghosted z index
        out[i] = (zg >= 0 && zg < nz + 2 * G) ? (T)vg[rw * (nz + 2 * G) + zg] : (T)0;
    }
}

// table of the neighbours' tensor maps of the PEER == 2 kernels, cached per block
struct NbTable
{
    const void* nb[27];
    const void* nbz[9];
    int nz, ny, nx, nfunc, TY, G, f64;
    long long ld;
    CUtensorMap* dev;
    unsigned long long stamp;
};
static NbTable g_nbtab[48];
static unsigned long long g_nbstamp = 0;

int nb_table(const void* const* nb, const void* const* nbz, bool f64, int nz, int ny, int nx,
    long long ld, int nfunc, int TY, int G, cudaStream_t st, const CUtensorMap** out)
{
    NbTable* slot = nullptr;
    // no buffers of pushed columns: the z maps lie over the z neighbours' blocks themselves
    const void* inpl[9] = { (const void*)1, nullptr };
    const bool inplace_z = nbz == nullptr;
    if (!nbz) nbz = inpl;
    for (auto& t : g_nbtab)
    {
        if (t.dev && t.nz == nz && t.ny == ny && t.nx == nx && t.nfunc == nfunc && t.TY == TY
            && t.G == G && t.f64 == (int)f64 && t.ld == ld && memcmp(t.nb, nb, sizeof(t.nb)) == 0
            && memcmp(t.nbz, nbz, sizeof(t.nbz)) == 0)
        {
            t.stamp = ++g_nbstamp;
            *out    = t.dev;
            return MGB_OK;
        }
        if (!slot || t.stamp < slot->stamp) slot = &t;
    }
    static CUtensorMap host[27 * 4];
    memset(host, 0, sizeof(host));
    const int ZH = f64 ? 2 : 4;
    for (int r = 0; r < 27; r++)
    {
        const int dx = r / 9, dy = (r / 3) % 3, dz = r % 3;
        for (int kind = 0; kind < 4; kind++)
        {
            const bool zbox = kind >= 2, mid = (kind & 1) == 0;
            if (zbox && inplace_z)
            {
                // z-halo columns read in place: 16 bytes of the rows of the rank across z
                if (dz == 1 || !nb[r]) continue;
                if (int rc = make_map_box(&host[r * 4 + kind], f64, nb[r], nz, ny, nx, ld, nfunc, ZH,
                        mid ? TY : G))
                    return rc;
                continue;
            }
            if (zbox != (dz == 0)) continue; // pushed z-halo columns: kept in the dz = -1 entries
            if (mid && dy != 1) continue;    // the tile's own rows never come from a y neighbour
            if (!zbox)
            {
                if (!nb[r]) continue;
                if (int rc = make_map_box(&host[r * 4 + kind], f64, nb[r], nz, ny, nx, ld, nfunc, nz,
                        mid ? TY : G))
                    return rc;
                continue;
            }
            // the dense buffer of pushed columns of the rank at (dx, dy, 0): 32 bytes per
            // row, the columns below z = 0 then the columns above z = nz-1
            const void* zb = nbz[dx * 3 + dy];
            if (!zb) continue;
            if (int rc = make_map_box(&host[r * 4 + kind], f64, zb, 2 * ZH, ny, nx,
                    (long long)2 * ZH * ny * nx, nfunc, 2 * ZH, mid ? TY : G))
                return rc;
        }
    }
    if (!slot->dev) MGB_CUDA(cudaMalloc(&slot->dev, sizeof(host)));
    // the table may still be read by a kernel in flight on another configuration
    MGB_CUDA(cudaStreamSynchronize(st));
    MGB_CUDA(cudaMemcpy(slot->dev, host, sizeof(host), cudaMemcpyHostToDevice));
    memcpy(slot->nb, nb, sizeof(slot->nb));
    memcpy(slot->nbz, nbz, sizeof(slot->nbz));
    slot->nz = nz, slot->ny = ny, slot->nx = nx, slot->nfunc = nfunc, slot->TY = TY, slot->G = G;
    slot->f64 = f64, slot->ld = ld;
    slot->stamp = ++g_nbstamp;
    *out        = slot->dev;
    return MGB_OK;
}

struct FusedCfg
{
    int RY, YG, NB, S, XC;
};

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

static bool g_layout_peer3 = false; // lay out the z-halo arrays of the PEER == 2 kernels
static bool cfg_layout(const FusedCfg& c, int G, int ny, int nz, int es,
    FusedParams& P, size_t& smem)
{
    const int TY = c.RY * c.YG;
    if (TY > ny || ny % TY) return false;
    if (TY < G) return false;
    const int vec = 16 / es;
    P.TY          = TY;
    P.YG          = c.YG;
    P.NB          = c.NB;
    P.S           = c.S;
    P.XC          = c.XC;
    P.nzv         = nz / vec;
    P.tpo         = round_up(c.YG * P.nzv, 32);
    if (32 + c.NB * P.tpo > 544) return false;
    P.row_bytes   = nz * es;
    P.off_mid     = round_up(G * P.row_bytes, 128);
    P.off_hi      = P.off_mid + round_up(TY * P.row_bytes, 128);
    P.tile_bytes  = P.off_hi + round_up(G * P.row_bytes, 128);
    if (g_layout_peer3)
    {
        // 32 bytes per tile row (the 16 bytes of columns below z = 0, the 16 bytes above
        // z = nz-1), each box at a 128-byte boundary: [G rows][TY rows][G rows]
        P.zlo_off    = P.tile_bytes;
        P.zarr_mid   = 128;
        P.zarr_hi    = 128 + round_up(TY * 32, 128);
        P.zarr_bytes = P.zarr_hi + 128;
        P.tile_bytes += 2 * P.zarr_bytes;
        P.vrow_bytes = P.row_bytes + 32;
    }
    P.stage_bytes = (c.NB + 1) * P.tile_bytes;
    smem          = (size_t)kBarBytes + (size_t)c.S * P.stage_bytes;
    return smem <= 227 * 1024 && c.S >= 2 && c.S <= 8;
}

static bool choose_cfg(bool lap4, int es, int nx, int ny, int nz, int nfunc,
    FusedCfg& best)
{
    const int G = lap4 ? 2 : 1;
    if (const char* env = getenv("MGB_HPSI_CFG"))
    {
        FusedCfg c;
        if (sscanf(env, "%d,%d,%d,%d,%d", &c.RY, &c.YG, &c.NB, &c.S, &c.XC) == 5)
        {
            if (c.XC <= 0 || c.XC > nx) c.XC = nx;
            FusedParams tmp;
            size_t sm;
            const bool ry_ok = lap4 ? (c.RY == 4 || c.RY == 2) : (c.RY == 8 || c.RY == 4);
            if (ry_ok && cfg_layout(c, G, ny, nz, es, tmp, sm))
            {
                best = c;
                return true;
            }
        }
    }
    // Boxes whose potential does not stay in L2 (256^3: 134 MB): the x range is cut
    // into chunks so that a chunk of V (evict_last) serves every orbital from L2 --
    // CUDA schedules blockIdx.z slowest, i.e. all orbitals x y tiles of one chunk
    // before the next -- at the price of 2G re-read planes per chunk.  One orbital
    // per CTA, 8- / 16-row tiles, the deepest ring that fits.  Measured on B200 at
    // 256^3 (tools/cfg_sweep.py, tools/cfg_try.py; profiles/r02_hpsi_256_cfg.md).
    const bool big_v = (size_t)nx * ny * nz * es > ((size_t)48 << 20);
    // ... and the same tile family for any box with 2 KB rows (nz = 256 doubles), whose
    // tiles the search below would make 8 rows x 3 orbitals with a 2-deep ring
    if (big_v || nz * es >= 2048)
    {
        // 8 consumer warps: as many orbitals per CTA as 256 threads cover
        const int tpo = round_up(2 * (nz / (16 / es)), 32);
        int nb        = 256 / tpo;
        if (nb < 1) nb = 1;
        if (nb > 4) nb = 4;
        if (nb > nfunc) nb = nfunc;
        for (int s = lap4 ? 4 : 3; s >= 2; s--)
        {
            FusedCfg c = { lap4 ? 4 : 8, 2, nb, s, big_v ? (lap4 ? 64 : 128) : nx };
            if (c.XC > nx) c.XC = nx;
            FusedParams tmp;
            size_t sm;
            if (cfg_layout(c, G, ny, nz, es, tmp, sm))
            {
                best = c;
                return true;
            }
        }
    }
    double best_cost = 1e30;
    bool found       = false;
    const int rys_m[2] = { 8, 4 }, rys_4[2] = { 4, 2 };
    for (int ri = 0; ri < 2; ri++)
        for (int yg = 1; yg <= 32; yg++)
            for (int nb = 1; nb <= 4; nb++)
                for (int s = 2; s <= 4; s++)
                {
                    FusedCfg c = { lap4 ? rys_4[ri] : rys_m[ri], yg, nb, s, nx };
                    if (nb > nfunc && nb > 1) continue;
                    FusedParams tmp;
                    size_t sm;
                    if (!cfg_layout(c, G, ny, nz, es, tmp, sm)) continue;
                    const int TY      = c.RY * c.YG;
                    const int threads = 32 + nb * tmp.tpo;
                    // L2->SM bytes per point relative to one read of phi
                    double cost = (double)(TY + 2 * G) / TY * (1. + 1. / nb);
                    if (threads < 256) cost *= 1.3;
                    if (threads < 128) cost *= 2.0;
                    if (threads > 416) cost *= 1.5; // 96-register class spills
                    const bool big_ry = lap4 ? c.RY == 4 : c.RY == 8;
                    if (big_ry && threads > 288) cost *= 1.05; // mild spills at 128
                    cost *= (s >= 3) ? 1.0 : 1.05;
                    if (cost < best_cost)
                    {
                        best_cost = cost;
                        best      = c;
                        found     = true;
                    }
                }
    if (!found) return false;
    // enough CTAs to fill the machine: split x into chunks if needed
    const long long ctas = (long long)(ny / (best.RY * best.YG))
                           * ((nfunc + best.NB - 1) / best.NB);
    if (ctas < 2 * 148)
    {
        int chunks = (int)((2 * 148 + ctas - 1) / ctas);
        int xc     = (nx + chunks - 1) / chunks;
        if (xc < 8) xc = nx < 8 ? nx : 8;
        best.XC = xc;
    }
    return true;
}

template <typename T, int RY, bool PERIODIC, bool LAP4>
static int launch_ry(const FusedParams& P, dim3 grid, int threads, size_t smem,
    cudaStream_t st)
{
    // launch-bound classes: 9 / 13 / 17 warps -> 168 / 128 / 96 registers
#define MGB_LAUNCH_MAXT_P(MT, PEERV)                                           \
    {                                                                          \
        auto kern = k_hpsi_tma<T, RY, PERIODIC, LAP4, MT, PEERV>;              \
        MGB_CUDA(cudaFuncSetAttribute(                                         \
            kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        kern<<<grid, threads, smem, st>>>(P);                                  \
        MGB_LAUNCHED("k_hpsi_tma");                                            \
    }
#define MGB_LAUNCH_MAXT(MT)                                                    \
    if (P.peer == 2)                                                           \
        MGB_LAUNCH_MAXT_P(MT, 2)                                               \
    else if (P.peer)                                                           \
        MGB_LAUNCH_MAXT_P(MT, 1)                                               \
    else                                                                       \
        MGB_LAUNCH_MAXT_P(MT, 0)
    if (threads <= 288)
        MGB_LAUNCH_MAXT(288)
    else if (threads <= 416)
        MGB_LAUNCH_MAXT(416)
    else
        MGB_LAUNCH_MAXT(544)
#undef MGB_LAUNCH_MAXT
#undef MGB_LAUNCH_MAXT_P
    return MGB_OK;
}

template <typename T, bool PERIODIC, bool LAP4>
static int launch_fused(const FusedParams& P, const FusedCfg& c, dim3 grid,
    int threads, size_t smem, cudaStream_t st)
{
    if constexpr (LAP4)
    {
        if (c.RY == 4) return launch_ry<T, 4, PERIODIC, LAP4>(P, grid, threads, smem, st);
        return launch_ry<T, 2, PERIODIC, LAP4>(P, grid, threads, smem, st);
    }
    else
    {
        if (c.RY == 8) return launch_ry<T, 8, PERIODIC, LAP4>(P, grid, threads, smem, st);
        return launch_ry<T, 4, PERIODIC, LAP4>(P, grid, threads, smem, st);
    }
}

int hpsi_tma(const HpsiArgs& a, cudaStream_t st)
{
    const mgb_grid* gr = a.grid;
    const bool f64     = a.dtype == MGB_F64;
    const int es       = f64 ? 8 : 4;
    const int vec      = 16 / es;
    const bool lap4    = a.lap_type == MGB_LAP_4;
    const int G        = a.g;
    const int nx = gr->dim[0], ny = gr->dim[1], nz = gr->dim[2];
    // ---- eligibility --------------------------------------------------------
    const bool periodic  = gr->bc[0] == 1 && gr->bc[1] == 1 && gr->bc[2] == 1;
    const bool dirichlet = gr->bc[0] == 0 && gr->bc[1] == 0 && gr->bc[2] == 0;
    if (!(periodic || dirichlet)) return MGB_ENOTSUP;
    const bool peer3 = a.nb3d != nullptr;
    if (!peer3 && (gr->nproc[1] != 1 || gr->nproc[2] != 1)) return MGB_ENOTSUP;
    if (peer3 && ((nz * es) % 32 || (!lap4 && !a.vghost))) return MGB_ENOTSUP;
    if (nz > 256 || nz % vec || nz < vec * 2) return MGB_ENOTSUP;
    if (a.ld % vec || a.ldh % vec) return MGB_ENOTSUP;
    if (((uintptr_t)a.phi | (uintptr_t)a.hphi | (uintptr_t)a.vtot) & 15)
        return MGB_ENOTSUP;
    if (nx < G || ny < 2 * G) return MGB_ENOTSUP;
    const bool split_x = gr->nproc[0] > 1 || peer3;
    const bool peer = !peer3 && split_x && a.peer_w && a.peer_e;
    if (!peer3 && split_x && ((!a.xhalo_phi && !peer) || !a.xhalo_v)) return MGB_ENOTSUP;
    if (!peer3 && split_x && (((uintptr_t)a.xhalo_phi | (uintptr_t)a.xhalo_v) & 15))
        return MGB_ENOTSUP;
    g_layout_peer3 = peer3;
    if (peer && (((uintptr_t)a.peer_w | (uintptr_t)a.peer_e) & 15)) return MGB_ENOTSUP;
    FusedCfg c;
    if (!choose_cfg(lap4, es, nx, ny, nz, a.nfunc, c))
    {
        g_layout_peer3 = false;
        return MGB_ENOTSUP;
    }

    FusedParams P;
    memset(&P, 0, sizeof(P));
    size_t smem = 0;
    const bool lay_ok = cfg_layout(c, G, ny, nz, es, P, smem);
    g_layout_peer3    = false;
    if (!lay_ok) return MGB_ENOTSUP;

    // potential in the compute type
    const void* vsrc  = a.vtot;
    const void* xvsrc = a.xhalo_v;
    if (!f64)
    {
        const size_t npt = (size_t)nx * ny * nz;
        const size_t nh  = (split_x && !peer3) ? (size_t)2 * G * ny * nz : 0;
        float* vf        = (float*)scratch(0, (npt + nh) * sizeof(float));
        if (!vf) return MGB_ECUDA;
        k_f64_to_f32<<<296, 256, 0, st>>>(npt, a.vtot, vf);
        MGB_LAUNCHED("k_f64_to_f32");
        if (nh)
        {
            k_f64_to_f32<<<64, 256, 0, st>>>(nh, a.xhalo_v, vf + npt);
            MGB_LAUNCHED("k_f64_to_f32");
        }
        vsrc  = vf;
        xvsrc = vf + npt;
    }

    const int TY = P.TY;
    int rc;
    if ((rc = make_map(&P.psi_mid, f64, a.phi, 4, nz, ny, nx, (long long)a.ld, a.nfunc, TY)))
        return rc;
    if ((rc = make_map(&P.psi_halo, f64, a.phi, 4, nz, ny, nx, (long long)a.ld, a.nfunc, G)))
        return rc;
    if ((rc = make_map(&P.v_mid, f64, vsrc, 3, nz, ny, nx, 0, 0, TY))) return rc;
    if ((rc = make_map(&P.v_halo, f64, vsrc, 3, nz, ny, nx, 0, 0, G))) return rc;
    if (peer3)
    {
        P.peer    = 2;
        P.first_y = gr->bc[1] != 1 && gr->coord[1] == 0;
        P.last_y  = gr->bc[1] != 1 && gr->coord[1] == gr->nproc[1] - 1;
        P.first_z = gr->bc[2] != 1 && gr->coord[2] == 0;
        P.last_z  = gr->bc[2] != 1 && gr->coord[2] == gr->nproc[2] - 1;
        P.zsplit  = gr->nproc[2] > 1 || getenv("MGB_HPSI_FORCE_ZBOXES") != nullptr;
        P.zstride = a.nbz ? 32 : 16;
        P.zhi_off = a.nbz ? 16 : P.zarr_bytes;
        if ((rc = nb_table(a.nb3d, a.nbz, f64, nz, ny, nx, (long long)a.ld, a.nfunc, TY, G, st,
                 &P.nbmaps)))
            return rc;
        if (!lap4)
        {
            // the potential with its halo: the caller's ghosted copy, padded to whole
            // 16-byte columns, in the compute type
            const int ZH  = 16 / es;
            const int nzp = nz + 2 * ZH;
            P.nvb         = nzp > 256 ? 2 : 1;
            const size_t n = (size_t)(nx + 2 * G) * (ny + 2 * G) * nzp;
            void* vp       = scratch(7, n * es);
            if (!vp) return MGB_ECUDA;
            if (f64)
                k_vpad<double><<<592, 256, 0, st>>>(nx + 2 * G, ny + 2 * G, nz, G, nzp, a.vghost,
                    (double*)vp);
            else
                k_vpad<float><<<592, 256, 0, st>>>(nx + 2 * G, ny + 2 * G, nz, G, nzp, a.vghost,
                    (float*)vp);
            MGB_LAUNCHED("k_vpad");
            if ((rc = make_map_vpad(&P.vg_map, f64, vp, nzp, P.nvb, ny + 2 * G, nx + 2 * G,
                     TY + 2 * G)))
                return rc;
        }
    }
    else if (split_x)
    {
        const long long hs = (long long)2 * G * ny * nz;
        if (peer)
        {
            P.peer  = 1;
            P.map_w = a.map_w;
            P.map_e = a.map_e;
            if ((rc = make_map(&P.pw_mid, f64, a.peer_w, 4, nz, ny, nx, (long long)a.ld, a.nfunc, TY)))
                return rc;
            if ((rc = make_map(&P.pw_halo, f64, a.peer_w, 4, nz, ny, nx, (long long)a.ld, a.nfunc, G)))
                return rc;
            if ((rc = make_map(&P.pe_mid, f64, a.peer_e, 4, nz, ny, nx, (long long)a.ld, a.nfunc, TY)))
                return rc;
            if ((rc = make_map(&P.pe_halo, f64, a.peer_e, 4, nz, ny, nx, (long long)a.ld, a.nfunc, G)))
                return rc;
            P.xpsi_mid  = P.psi_mid;
            P.xpsi_halo = P.psi_halo;
        }
        else
        {
            if ((rc = make_map(
                     &P.xpsi_mid, f64, a.xhalo_phi, 4, nz, ny, 2 * G, hs, a.nfunc, TY)))
                return rc;
            if ((rc = make_map(
                     &P.xpsi_halo, f64, a.xhalo_phi, 4, nz, ny, 2 * G, hs, a.nfunc, G)))
                return rc;
        }
        if ((rc = make_map(&P.xv_mid, f64, xvsrc, 3, nz, ny, 2 * G, 0, 0, TY)))
            return rc;
        if ((rc = make_map(&P.xv_halo, f64, xvsrc, 3, nz, ny, 2 * G, 0, 0, G)))
            return rc;
    }
    else
    {
        P.xpsi_mid  = P.psi_mid;
        P.xpsi_halo = P.psi_halo;
        P.xv_mid    = P.v_mid;
        P.xv_halo   = P.v_halo;
    }
    P.out     = a.hphi;
    P.ldh     = (long long)a.ldh;
    P.nx      = nx;
    P.ny      = ny;
    P.nz      = nz;
    P.nfunc   = a.nfunc;
    P.split_x = split_x;
    P.first_x = gr->coord[0] == 0;
    P.last_x  = gr->coord[0] == gr->nproc[0] - 1;
    P.per_y   = gr->bc[1] == 1;
    P.per_z   = gr->bc[2] == 1;
    const double inv12 = 1. / 12.;
    const double i2[3] = { 1. / (gr->h[0] * gr->h[0]), 1. / (gr->h[1] * gr->h[1]),
        1. / (gr->h[2] * gr->h[2]) };
    double* cd = P.cd;
    cd[kC0]  = 16. * inv12 * (i2[0] + i2[1] + i2[2]);
    cd[kCx]  = -10. * inv12 * i2[0] + 0.125 * cd[kC0];
    cd[kCy]  = -10. * inv12 * i2[1] + 0.125 * cd[kC0];
    cd[kCz]  = -10. * inv12 * i2[2] + 0.125 * cd[kC0];
    cd[kCxy] = -inv12 * (i2[0] + i2[1]);
    cd[kCyz] = -inv12 * (i2[2] + i2[1]);
    cd[kCxz] = -inv12 * (i2[0] + i2[2]);
    cd[kD1x] = -16. * inv12 * i2[0];
    cd[kD2x] = inv12 * i2[0];
    cd[kD1y] = -16. * inv12 * i2[1];
    cd[kD2y] = inv12 * i2[1];
    cd[kD1z] = -16. * inv12 * i2[2];
    cd[kD2z] = inv12 * i2[2];
    cd[kD0]  = -2. * (cd[kD1x] + cd[kD2x] + cd[kD1y] + cd[kD2y] + cd[kD1z] + cd[kD2z]);
    for (int i = 0; i < kNumCoef; i++)
        P.cf[i] = (float)cd[i];

    dim3 grid((unsigned)(ny / TY), (unsigned)((a.nfunc + c.NB - 1) / c.NB),
        (unsigned)((nx + c.XC - 1) / c.XC));
    if (grid.y > 65535 || grid.z > 65535) return MGB_ENOTSUP;
    const int threads = 32 + c.NB * P.tpo;
    snprintf(g_last_kernel, sizeof(g_last_kernel),
        "k_hpsi_tma<%s,RY=%d,PERIODIC=%d,LAP4=%d,MAXT=%d,PEER=%d> YG=%d NB=%d S=%d XC=%d",
        f64 ? "double" : "float", c.RY, periodic ? 1 : 0, lap4 ? 1 : 0,
        threads <= 288 ? 288 : (threads <= 416 ? 416 : 544), P.peer, c.YG, c.NB, c.S, c.XC);
    // A potential that does not stay in L2 (256^3 doubles = 134 MB) is fetched
    // from HBM once per wave when the CTAs of a wave work on the same y tile
    // of different orbitals; a resident one (128^3) keeps the y-fast order.
    {
        const size_t vbytes = (size_t)nx * ny * nz * es;
        int of              = 0; // orbital-fast order loses the neighbour tiles' halo rows from L2 (measured)
        (void)vbytes;
        if (const char* env = getenv("MGB_HPSI_ORDER")) of = atoi(env);
        P.pol = 1; // measured: +4 % at 128^3 over evict_first (halo rows of the neighbour tile hit L2)
        if (const char* env = getenv("MGB_HPSI_POL")) P.pol = atoi(env);
        if (of)
        {
            P.orb_fast = 1;
            const unsigned t = grid.x;
            grid.x           = grid.y;
            grid.y           = t;
            if (grid.y > 65535) return MGB_ENOTSUP;
        }
    }

    if (f64)
    {
        if (periodic)
            return lap4 ? launch_fused<double, true, true>(P, c, grid, threads, smem, st)
                        : launch_fused<double, true, false>(P, c, grid, threads, smem, st);
        return lap4 ? launch_fused<double, false, true>(P, c, grid, threads, smem, st)
                    : launch_fused<double, false, false>(P, c, grid, threads, smem, st);
    }
    if (periodic)
        return lap4 ? launch_fused<float, true, true>(P, c, grid, threads, smem, st)
                    : launch_fused<float, true, false>(P, c, grid, threads, smem, st);
    return lap4 ? launch_fused<float, false, true>(P, c, grid, threads, smem, st)
                : launch_fused<float, false, false>(P, c, grid, threads, smem, st);
}

} // namespace mgb
