// Generic fused local-Hamiltonian kernel (path 2 of mgb_hpsi).
//
// One thread per output point, no shared memory: reads phi / vtot straight from
// the no-ghost arrays with the boundary condition folded into the index
// (periodic wrap, Dirichlet zero + zeroed first low layer, or the x-halo
// buffer of a neighbouring rank).  It accepts every local box shape, which the
// TMA-pipelined kernel in hpsi_fused.cu does not, and -- compiled with
// -fmad=false and written in the reference's operand order with the
// reference's intermediate casts -- it is BIT-IDENTICAL to
// Hamiltonian::applyLocal (src/Hamiltonian.cc:85-159) on the CPU.
#include "common.cuh"
#include "hpsi.h"

namespace mgb
{

template <typename T>
struct GenericView
{
    const T* phi;       // + f*ld
    const T* xhalo;     // [2g][ny][nz] of this function, or nullptr
    const double* v;    // [nx][ny][nz]
    const double* xhv;  // [2g][ny][nz] or nullptr
    int nx, ny, nz, g;
    bool periodic;      // all three directions periodic, else all Dirichlet
    bool first_x, last_x, split_x;

    // value of the ghosted, boundary-traded field at local (ix,iy,iz)
    __device__ __forceinline__ bool locate(int& ix, int& iy, int& iz, bool& halo) const
    {
        halo = false;
        if (periodic)
        {
            iy = (iy < 0) ? iy + ny : (iy >= ny ? iy - ny : iy);
            iz = (iz < 0) ? iz + nz : (iz >= nz ? iz - nz : iz);
            if (ix < 0 || ix >= nx)
            {
                if (split_x)
                {
                    halo = true;
                    ix   = (ix < 0) ? ix + g : g + (ix - nx);
                }
                else
                    ix = (ix < 0) ? ix + nx : ix - nx;
            }
            return true;
        }
        // Dirichlet-0: ghosts and the first low layer are zero
        // (src/pb/GridFunc.cc:2188-2336)
        if (iy <= 0 || iy >= ny || iz <= 0 || iz >= nz) return false;
        if (ix < 0)
        {
            if (first_x) return false;
            halo = true;
            ix   = ix + g;
            return true;
        }
        if (ix >= nx)
        {
            if (last_x) return false;
            halo = true;
            ix   = g + (ix - nx);
            return true;
        }
        if (ix == 0 && first_x) return false;
        return true;
    }
    __device__ __forceinline__ T psi(int ix, int iy, int iz) const
    {
        bool halo;
        if (!locate(ix, iy, iz, halo)) return (T)0;
        const long long o = ((long long)ix * ny + iy) * nz + iz;
        return halo ? xhalo[o] : phi[o];
    }
    __device__ __forceinline__ double pot(int ix, int iy, int iz) const
    {
        bool halo;
        if (!locate(ix, iy, iz, halo)) return 0.;
        const long long o = ((long long)ix * ny + iy) * nz + iz;
        return halo ? xhv[o] : v[o];
    }
    // GridFuncVector::pointwiseProduct element, (T)(v1*v2)
    __device__ __forceinline__ T vpsi(int ix, int iy, int iz) const
    {
        return (T)(psi(ix, iy, iz) * pot(ix, iy, iz));
    }
};

struct GenericCoef
{
    // Mehrstellen (FDkernels.cc:457-463)
    double c0, cx, cy, cz, cxy, cyz, cxz;
    // 4th order (FDoper.cc:331-343)
    double d0, d1x, d2x, d1y, d2y, d1z, d2z;
};

template <typename T>
__global__ void k_hpsi_generic(int lap4, GenericCoef c, GenericView<T> view0,
    long long ld, long long ldh, long long halo_stride, T* __restrict__ hphi,
    int tiles_z)
{
    const int tz = blockIdx.x % tiles_z;
    const int ty = blockIdx.x / tiles_z;
    const int iz = tz * blockDim.x + threadIdx.x;
    const int iy = ty * blockDim.y + threadIdx.y;
    const int ix = blockIdx.y;
    const int f  = blockIdx.z;
    GenericView<T> w = view0;
    if (iz >= w.nz || iy >= w.ny) return;
    w.phi += (long long)f * ld;
    if (w.xhalo) w.xhalo += (long long)f * halo_stride;
    const long long o = ((long long)ix * w.ny + iy) * w.nz + iz;
    T out;
    if (lap4)
    {
        // FDoper<T>::del2_4th_withPot, src/pb/FDoper.cc:379-387
        out = (T)((c.d0 + w.v[o]) * (double)w.psi(ix, iy, iz)
                  + c.d1x
                        * ((double)w.psi(ix - 1, iy, iz)
                            + (double)w.psi(ix + 1, iy, iz))
                  + c.d1y
                        * ((double)w.psi(ix, iy - 1, iz)
                            + (double)w.psi(ix, iy + 1, iz))
                  + c.d1z
                        * ((double)w.psi(ix, iy, iz - 1)
                            + (double)w.psi(ix, iy, iz + 1))
                  + c.d2x
                        * ((double)w.psi(ix - 2, iy, iz)
                            + (double)w.psi(ix + 2, iy, iz))
                  + c.d2y
                        * ((double)w.psi(ix, iy - 2, iz)
                            + (double)w.psi(ix, iy + 2, iz))
                  + c.d2z
                        * ((double)w.psi(ix, iy, iz - 2)
                            + (double)w.psi(ix, iy, iz + 2)));
    }
    else
    {
        // work1 = B (V psi): FDkernelRHS_4th_Mehr1 (FDkernels.cc:571-575) on
        // the pointwise product (GridFuncVector.cc:131)
        const T bw = (T)(0.5 * (double)w.vpsi(ix, iy, iz)
                         + (1. / 12.)
                               * (double)(w.vpsi(ix - 1, iy, iz)
                                          + w.vpsi(ix + 1, iy, iz)
                                          + w.vpsi(ix, iy - 1, iz)
                                          + w.vpsi(ix, iy + 1, iz)
                                          + w.vpsi(ix, iy, iz - 1)
                                          + w.vpsi(ix, iy, iz + 1)));
        // gfvw1 = -Lap phi: FDkernelDel2_4th_Mehr (FDkernels.cc:499-512)
        const T lap = (T)(
            c.c0 * (double)w.psi(ix, iy, iz)
            + c.cz * (double)(w.psi(ix, iy, iz - 1) + w.psi(ix, iy, iz + 1))
            + c.cy * (double)(w.psi(ix, iy - 1, iz) + w.psi(ix, iy + 1, iz))
            + c.cx * (double)(w.psi(ix - 1, iy, iz) + w.psi(ix + 1, iy, iz))
            + c.cxz
                  * (double)(w.psi(ix - 1, iy, iz - 1) + w.psi(ix - 1, iy, iz + 1)
                             + w.psi(ix + 1, iy, iz - 1)
                             + w.psi(ix + 1, iy, iz + 1))
            + c.cyz
                  * (double)(w.psi(ix, iy - 1, iz - 1) + w.psi(ix, iy - 1, iz + 1)
                             + w.psi(ix, iy + 1, iz - 1)
                             + w.psi(ix, iy + 1, iz + 1))
            + c.cxy
                  * (double)(w.psi(ix - 1, iy - 1, iz) + w.psi(ix + 1, iy - 1, iz)
                             + w.psi(ix - 1, iy + 1, iz)
                             + w.psi(ix + 1, iy + 1, iz)));
        // gfv_work1.axpy(1., gfvw1): y += (T)(1.0 * (double)x)
        out = bw;
        out += (T)(1. * (double)lap);
    }
    hphi[(long long)f * ldh + o] = out;
}

template <typename T>
int hpsi_generic_t(const HpsiArgs& a, cudaStream_t st)
{
    const mgb_grid* gr = a.grid;
    GenericView<T> w;
    w.phi      = (const T*)a.phi;
    w.xhalo    = (const T*)a.xhalo_phi;
    w.v        = a.vtot;
    w.xhv      = a.xhalo_v;
    w.nx       = gr->dim[0];
    w.ny       = gr->dim[1];
    w.nz       = gr->dim[2];
    w.g        = a.g;
    w.periodic = (gr->bc[0] == 1);
    w.split_x  = gr->nproc[0] > 1;
    w.first_x  = gr->coord[0] == 0;
    w.last_x   = gr->coord[0] == gr->nproc[0] - 1;
    GenericCoef c;
    const double inv12 = 1. / 12.;
    const double i2[3] = { 1. / (gr->h[0] * gr->h[0]), 1. / (gr->h[1] * gr->h[1]),
        1. / (gr->h[2] * gr->h[2]) };
    c.c0  = 16. * inv12 * (i2[0] + i2[1] + i2[2]);
    c.cx  = -10. * inv12 * i2[0] + 0.125 * c.c0;
    c.cy  = -10. * inv12 * i2[1] + 0.125 * c.c0;
    c.cz  = -10. * inv12 * i2[2] + 0.125 * c.c0;
    c.cxy = -inv12 * (i2[0] + i2[1]);
    c.cyz = -inv12 * (i2[2] + i2[1]);
    c.cxz = -inv12 * (i2[0] + i2[2]);
    const double cc0 = inv12 * i2[0], cc1 = inv12 * i2[1], cc2 = inv12 * i2[2];
    c.d1x = -16. * cc0;
    c.d2x = 1. * cc0;
    c.d1y = -16. * cc1;
    c.d2y = 1. * cc1;
    c.d1z = -16. * cc2;
    c.d2z = 1. * cc2;
    c.d0  = -2. * (c.d1x + c.d2x + c.d1y + c.d2y + c.d1z + c.d2z);

    RowLaunch L = row_launch(w.nx, w.ny, w.nz, 1);
    const long long halo_stride = (long long)2 * a.g * w.ny * w.nz;
    // grid.z <= 65535 functions per launch
    for (int f0 = 0; f0 < a.nfunc; f0 += 65535)
    {
        const int nf = (a.nfunc - f0 < 65535) ? a.nfunc - f0 : 65535;
        dim3 grid(L.grid.x, L.grid.y, (unsigned)nf);
        GenericView<T> wf = w;
        wf.phi += (long long)f0 * a.ld;
        if (wf.xhalo) wf.xhalo += (long long)f0 * halo_stride;
        k_hpsi_generic<T><<<grid, L.block, 0, st>>>(a.lap_type == MGB_LAP_4, c,
            wf, (long long)a.ld, (long long)a.ldh, halo_stride,
            (T*)a.hphi + (long long)f0 * a.ldh, L.tiles_z);
        MGB_LAUNCHED("k_hpsi_generic");
    }
    return MGB_OK;
}

// B u = u/2 + (1/12) sum of the 6 face neighbours on a no-ghost block with the
// boundary condition folded into the index: Lap<T>::rhs = FDoper::rhs_4th_Mehr1
// / FDkernelRHS_4th_Mehr1 (src/pb/FDkernels.cc:522-584) after setDataWithGhosts
// + trade_boundaries, as MGmol::computeResidualUsingHPhi applies it to every
// orbital (src/MGmol.cc:1252-1260).  Bit-identical (this unit is -fmad=false).
// MEHR2: B2 u = 2/3 u + 1/36 (faces) + 1/72 (edges), Laph4MP's operator
// (FDoper::rhs_4th_Mehr2(GridFunc&, T*), src/pb/FDoper.cc:573-637).
template <typename T, bool MEHR2>
__global__ void k_rhs_generic(GenericView<T> view0, long long ld, long long ldo,
    long long halo_stride, T* __restrict__ out, int tiles_z)
{
    const int tz = blockIdx.x % tiles_z;
    const int ty = blockIdx.x / tiles_z;
    const int iz = tz * blockDim.x + threadIdx.x;
    const int iy = ty * blockDim.y + threadIdx.y;
    const int ix = blockIdx.y;
    const int f  = blockIdx.z;
    GenericView<T> w = view0;
    if (iz >= w.nz || iy >= w.ny) return;
    w.phi += (long long)f * ld;
    if (w.xhalo) w.xhalo += (long long)f * halo_stride;
    const long long o = ((long long)ix * w.ny + iy) * w.nz + iz;
    if (MEHR2)
    {
        out[(long long)f * ldo + o]
            = (T)((2. / 3.) * (double)w.psi(ix, iy, iz)
                  + (1. / 36.)
                        * (double)(w.psi(ix - 1, iy, iz) + w.psi(ix + 1, iy, iz)
                                   + w.psi(ix, iy - 1, iz) + w.psi(ix, iy + 1, iz)
                                   + w.psi(ix, iy, iz - 1) + w.psi(ix, iy, iz + 1))
                  + (1. / 72.)
                        * (double)(w.psi(ix - 1, iy - 1, iz) + w.psi(ix + 1, iy - 1, iz)
                                   + w.psi(ix - 1, iy + 1, iz) + w.psi(ix + 1, iy + 1, iz)
                                   + w.psi(ix, iy - 1, iz - 1) + w.psi(ix, iy - 1, iz + 1)
                                   + w.psi(ix, iy + 1, iz - 1) + w.psi(ix, iy + 1, iz + 1)
                                   + w.psi(ix - 1, iy, iz - 1) + w.psi(ix - 1, iy, iz + 1)
                                   + w.psi(ix + 1, iy, iz - 1) + w.psi(ix + 1, iy, iz + 1)));
        return;
    }
    out[(long long)f * ldo + o]
        = (T)(0.5 * (double)w.psi(ix, iy, iz)
              + (1. / 12.)
                    * (double)(w.psi(ix - 1, iy, iz) + w.psi(ix + 1, iy, iz)
                               + w.psi(ix, iy - 1, iz) + w.psi(ix, iy + 1, iz)
                               + w.psi(ix, iy, iz - 1) + w.psi(ix, iy, iz + 1)));
}

// B u on whole rows, 16 bytes of z per thread (the common case: box owned by one rank, z
// extent a multiple of the vector, 16-byte aligned blocks).  Same expression, operand order and
// roundings as k_rhs_generic -- bit-identical -- but the five row loads are coalesced vector
// loads (the y neighbours hit L1, the x neighbours L2: CUDA runs the y tiles of a plane, then the
// next plane of the same function) and there is no per-tap index logic: k_rhs_generic spends its
// time on that and moves 1.3 TB/s at 256^3 x 512.
template <typename T>
struct RowVec;
template <>
struct RowVec<double>
{
    typedef double2 type;
    static constexpr int V = 2;
};
template <>
struct RowVec<float>
{
    typedef float4 type;
    static constexpr int V = 4;
};

template <typename T, bool PERIODIC>
__global__ void k_rhs_rows(const T* __restrict__ phi, long long ld, T* __restrict__ out,
    long long ldo, int nx, int ny, int nz, int tiles_z)
{
    constexpr int V = RowVec<T>::V;
    typedef typename RowVec<T>::type VT;
    const int zc = (blockIdx.x % tiles_z) * blockDim.x + threadIdx.x;
    const int iy = (blockIdx.x / tiles_z) * blockDim.y + threadIdx.y;
    const int ix = blockIdx.y;
    const int z0 = zc * V;
    if (z0 >= nz || iy >= ny) return;
    const T* p = phi + (long long)blockIdx.z * ld;
    // row (jx, jy) of the boundary-traded field, or null where it is zero (Dirichlet-0: ghosts
    // and the first low layer, src/pb/GridFunc.cc:2188-2336)
    auto row = [&](int jx, int jy) -> const T* {
        if (PERIODIC)
        {
            jx = jx < 0 ? jx + nx : (jx >= nx ? jx - nx : jx);
            jy = jy < 0 ? jy + ny : (jy >= ny ? jy - ny : jy);
        }
        else if (jx <= 0 || jx >= nx || jy <= 0 || jy >= ny)
            return nullptr;
        return p + ((long long)jx * ny + jy) * nz;
    };
    T c[V], xm[V], xp[V], ym[V], yp[V];
    auto loadv = [&](const T* r, T (&d)[V]) {
        if (r)
        {
            const VT v = *reinterpret_cast<const VT*>(r + z0);
            const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
            for (int i = 0; i < V; i++) d[i] = e[i];
        }
        else
        {
#pragma unroll
            for (int i = 0; i < V; i++) d[i] = (T)0;
        }
    };
    const T* rc = row(ix, iy);
    loadv(rc, c);
    loadv(row(ix - 1, iy), xm);
    loadv(row(ix + 1, iy), xp);
    loadv(row(ix, iy - 1), ym);
    loadv(row(ix, iy + 1), yp);
    T zlo, zhi; // the z neighbours outside this thread's vector
    if (PERIODIC)
    {
        zlo = rc[z0 == 0 ? nz - 1 : z0 - 1];
        zhi = rc[z0 + V == nz ? 0 : z0 + V];
    }
    else
    {
        zlo = (rc && z0 - 1 > 0) ? rc[z0 - 1] : (T)0;
        zhi = (rc && z0 + V < nz) ? rc[z0 + V] : (T)0;
        if (z0 == 0) c[0] = xm[0] = xp[0] = ym[0] = yp[0] = (T)0; // first low layer in z
    }
    T o[V];
#pragma unroll
    for (int i = 0; i < V; i++)
    {
        const T zm = i == 0 ? zlo : c[i - 1];
        const T zp = i == V - 1 ? zhi : c[i + 1];
        o[i] = (T)(0.5 * (double)c[i] + (1. / 12.) * (double)(xm[i] + xp[i] + ym[i] + yp[i] + zm + zp));
    }
    *reinterpret_cast<VT*>(out + (long long)blockIdx.z * ldo + ((long long)ix * ny + iy) * nz + z0)
        = *reinterpret_cast<const VT*>(o);
}

template <typename T>
static int rhs_generic_t(bool mehr2, const mgb_grid* gr, const T* phi, size_t ld, const T* xhalo,
    T* out, size_t ldo, int nfunc, cudaStream_t st)
{
    GenericView<T> w;
    w.phi      = phi;
    w.xhalo    = xhalo;
    w.v        = nullptr;
    w.xhv      = nullptr;
    w.nx       = gr->dim[0];
    w.ny       = gr->dim[1];
    w.nz       = gr->dim[2];
    w.g        = 1;
    w.periodic = (gr->bc[0] == 1);
    w.split_x  = gr->nproc[0] > 1;
    w.first_x  = gr->coord[0] == 0;
    w.last_x   = gr->coord[0] == gr->nproc[0] - 1;
    constexpr int V = RowVec<T>::V;
    if (!mehr2 && !w.split_x && !xhalo && w.nz % V == 0
        && (((uintptr_t)phi | (uintptr_t)out) & 15) == 0 && (ld * sizeof(T)) % 16 == 0
        && (ldo * sizeof(T)) % 16 == 0)
    {
        const int nzv = w.nz / V;
        int bx = 32;
        while (bx < nzv && bx < 128) bx *= 2;
        const int by      = 256 / bx;
        const int tiles_z = (nzv + bx - 1) / bx;
        dim3 block(bx, by);
        for (int f0 = 0; f0 < nfunc; f0 += 65535)
        {
            const int nf = (nfunc - f0 < 65535) ? nfunc - f0 : 65535;
            dim3 grid((unsigned)(tiles_z * ((w.ny + by - 1) / by)), (unsigned)w.nx, (unsigned)nf);
            if (w.periodic)
                k_rhs_rows<T, true><<<grid, block, 0, st>>>(phi + (long long)f0 * ld, (long long)ld,
                    out + (long long)f0 * ldo, (long long)ldo, w.nx, w.ny, w.nz, tiles_z);
            else
                k_rhs_rows<T, false><<<grid, block, 0, st>>>(phi + (long long)f0 * ld, (long long)ld,
                    out + (long long)f0 * ldo, (long long)ldo, w.nx, w.ny, w.nz, tiles_z);
            MGB_LAUNCHED("k_rhs_rows");
        }
        return MGB_OK;
    }
    RowLaunch L = row_launch(w.nx, w.ny, w.nz, 1);
    const long long halo_stride = (long long)2 * w.ny * w.nz;
    for (int f0 = 0; f0 < nfunc; f0 += 65535)
    {
        const int nf = (nfunc - f0 < 65535) ? nfunc - f0 : 65535;
        dim3 grid(L.grid.x, L.grid.y, (unsigned)nf);
        GenericView<T> wf = w;
        wf.phi += (long long)f0 * ld;
        if (wf.xhalo) wf.xhalo += (long long)f0 * halo_stride;
        if (mehr2)
            k_rhs_generic<T, true><<<grid, L.block, 0, st>>>(wf, (long long)ld, (long long)ldo,
                halo_stride, out + (long long)f0 * ldo, L.tiles_z);
        else
            k_rhs_generic<T, false><<<grid, L.block, 0, st>>>(wf, (long long)ld, (long long)ldo,
                halo_stride, out + (long long)f0 * ldo, L.tiles_z);
        MGB_LAUNCHED("k_rhs_generic");
    }
    return MGB_OK;
}

int rhs_generic(int dtype, bool mehr2, const mgb_grid* gr, const void* phi, size_t ld,
    const void* xhalo, void* out, size_t ldo, int nfunc, cudaStream_t st)
{
    return dtype == MGB_F64 ? rhs_generic_t<double>(mehr2, gr, (const double*)phi, ld,
                                  (const double*)xhalo, (double*)out, ldo, nfunc, st)
                            : rhs_generic_t<float>(mehr2, gr, (const float*)phi, ld,
                                  (const float*)xhalo, (float*)out, ldo, nfunc, st);
}

int hpsi_generic(const HpsiArgs& a, cudaStream_t st)
{
    return a.dtype == MGB_F64 ? hpsi_generic_t<double>(a, st)
                              : hpsi_generic_t<float>(a, st);
}

} // namespace mgb
