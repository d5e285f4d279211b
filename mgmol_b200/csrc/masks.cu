// Localization masks of LocGridOrbitals: the data Map2Masks /
// GridMask hold per gid, per multigrid level and per x-slab
// (src/Map2Masks.cc:25-61, src/GridMask.h:41-52), flattened per COLOR -- the
// caller resolves overlapping_gids_[iloc][color] -> mask when it fills the set
// (gid == -1 slots are "zero", src/Map2Masks.cc:53-58) -- and the kernels that
// apply them to ghosted (GridFuncVector::app_mask, src/pb/GridFuncVector.cc:
// 2428-2438) and no-ghost blocks (LocGridOrbitals::applyMask, src/
// LocGridOrbitals.cc:427-452).  The fused multigrid kernels (mg_fused.cu) read
// the same tables.
#include <cstring>
#include <vector>

#include "masks.h"

struct mgb_mask_level
{
    int dim[3];
    int sub0;
    long long slab;
    std::vector<int> off;       // host table
    std::vector<double> values; // host pool (slab-sized chunks)
    int nslabs;
    int* off_dev;
    float* pool_f;
    double* pool_d;
};

struct mgb_masks
{
    mgb_grid grid;
    int nlevels, subdivx, ncolors, op;
    bool committed;
    std::vector<mgb_mask_level> lev;
};

namespace mgb
{

// the double copy of the values is only needed to mask ORBDTYPE double blocks
// (the V-cycle is float): uploaded on first use
static int ensure_pool64(mgb_mask_level& L)
{
    if (L.pool_d) return MGB_OK;
    const size_t n = L.values.size();
    MGB_CUDA(cudaMalloc(&L.pool_d, sizeof(double) * n));
    MGB_CUDA(cudaMemcpy(L.pool_d, L.values.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    return MGB_OK;
}

MaskView mask_view(const mgb_masks* m, int level, int* rc)
{
    MaskView v = no_mask();
    *rc        = MGB_OK;
    if (!m) return v;
    if (!m->committed || level < 0 || level >= m->nlevels)
    {
        set_error(!m->committed ? "mask set used before mgb_masks_commit"
                                : "mask set has no level %d",
            level);
        *rc = MGB_EINVAL;
        return v;
    }
    const mgb_mask_level& L = m->lev[level];
    v.off     = L.off_dev;
    v.pool    = L.pool_f;
    v.pool64  = L.pool_d;
    v.sub0    = L.sub0;
    v.subdivx = m->subdivx;
    v.op      = m->op;
    v.slab    = L.slab;
    return v;
}

int masks_ncolors(const mgb_masks* m) { return m->ncolors; }
int masks_nlevels(const mgb_masks* m) { return m->nlevels; }
bool masks_match(const mgb_masks* m, const mgb_grid* g)
{
    return m->grid.dim[0] == g->dim[0] && m->grid.dim[1] == g->dim[1]
           && m->grid.dim[2] == g->dim[2];
}

// ghosted block: whole interior x planes INCLUDING their y/z ghosts are
// zeroed for a "zero" slab (memset of incx * sub0 values, src/GridMask.h:
// 169-180, src/Map2Masks.cc:55-58); values touch interior points only.
template <typename T>
__global__ void k_app_mask_ghosted(Box b, MaskView mv, T* __restrict__ u)
{
    const int f  = blockIdx.z;
    const int ix = blockIdx.y;
    const int o  = mv.off[f * mv.subdivx + ix / mv.sub0];
    if (o == -1) return;
    const long long rowlen = b.incy; // nz + 2g
    const long long nrows  = b.ny + 2 * b.g;
    T* plane = u + (long long)f * b.sizeg + (long long)(ix + b.g) * b.incx;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nrows * rowlen;
         t += (long long)gridDim.x * blockDim.x)
    {
        if (o == -2)
        {
            plane[t] = (T)0;
            continue;
        }
        const int iy = (int)(t / rowlen) - b.g, iz = (int)(t % rowlen) - b.g;
        if (iy < 0 || iy >= b.ny || iz < 0 || iz >= b.nz) continue;
        const long long mi
            = (long long)o * mv.slab + ((long long)(ix % mv.sub0) * b.ny + iy) * b.nz + iz;
        if (sizeof(T) == 8 && mv.pool64)
            plane[t] = mask_apply(mv.op, plane[t], mv.pool64[mi]);
        else
            // a float pool holds exactly representable values: widening it
            // is the same comparison / product as with lmasktype float
            plane[t] = mask_apply(mv.op, plane[t], mv.pool[mi]);
    }
}

template <typename T>
__global__ void k_app_mask_noghost(long long plane, int nx, MaskView mv, T* __restrict__ u,
    long long ld)
{
    const int f  = blockIdx.z;
    const int ix = blockIdx.y;
    const int o  = mv.off[f * mv.subdivx + ix / mv.sub0];
    if (o == -1) return;
    T* p = u + (long long)f * ld + (long long)ix * plane;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < plane;
         t += (long long)gridDim.x * blockDim.x)
    {
        if (o == -2)
        {
            p[t] = (T)0;
            continue;
        }
        const long long mi = (long long)o * mv.slab + (long long)(ix % mv.sub0) * plane + t;
        if (sizeof(T) == 8 && mv.pool64)
            p[t] = mask_apply(mv.op, p[t], mv.pool64[mi]);
        else
            p[t] = mask_apply(mv.op, p[t], mv.pool[mi]);
    }
}

static void free_level(mgb_mask_level& L)
{
    if (L.off_dev) cudaFree(L.off_dev);
    if (L.pool_f) cudaFree(L.pool_f);
    if (L.pool_d) cudaFree(L.pool_d);
    L.off_dev = nullptr;
    L.pool_f  = nullptr;
    L.pool_d  = nullptr;
}

} // namespace mgb

using namespace mgb;

extern "C"
{

int mgb_masks_create(mgb_masks** out, const mgb_grid* grid, int mg_levels, int subdivx,
    int ncolors, int op)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(grid)) return rc;
    MGB_REQUIRE(out, "mgb_masks_create: null out pointer");
    MGB_REQUIRE(mg_levels >= 0 && mg_levels < 10, "mgb_masks_create: nclevels %d", mg_levels);
    MGB_REQUIRE(subdivx > 0 && subdivx < 100, "mgb_masks_create: subdivx %d", subdivx);
    MGB_REQUIRE(ncolors > 0, "mgb_masks_create: ncolors %d", ncolors);
    MGB_REQUIRE(op == MGB_MASK_MULT || op == MGB_MASK_MAX, "mgb_masks_create: op %d", op);
    // subdim0_[l] = (dim(0) >> l) / subdivx must tile the box on every level
    // (src/GridMask.cc:60-70)
    MGB_REQUIRE((grid->dim[0] >> mg_levels) % subdivx == 0 && (grid->dim[0] >> mg_levels) > 0,
        "mgb_masks_create: dim[0]=%d not divisible by subdivx * 2^levels", grid->dim[0]);
    mgb_masks* m = new mgb_masks();
    m->grid      = *grid;
    m->nlevels   = mg_levels + 1;
    m->subdivx   = subdivx;
    m->ncolors   = ncolors;
    m->op        = op;
    m->committed = false;
    m->lev.resize(m->nlevels);
    for (int l = 0; l < m->nlevels; l++)
    {
        mgb_mask_level& L = m->lev[l];
        for (int d = 0; d < 3; d++)
            L.dim[d] = grid->dim[d] >> l;
        L.sub0   = L.dim[0] / subdivx;
        L.slab   = (long long)L.sub0 * L.dim[1] * L.dim[2];
        L.off.assign((size_t)ncolors * subdivx, -1); // masks default to 1
        L.nslabs  = 0;
        L.off_dev = nullptr;
        L.pool_f  = nullptr;
        L.pool_d  = nullptr;
    }
    *out = m;
    return MGB_OK;
}

int mgb_masks_set(mgb_masks* m, int level, int iloc, int color, int state, int dtype,
    const void* values_host)
{
    MGB_REQUIRE(m, "mgb_masks_set: null handle");
    MGB_REQUIRE(level >= 0 && level < m->nlevels, "mgb_masks_set: level %d", level);
    MGB_REQUIRE(iloc >= 0 && iloc < m->subdivx, "mgb_masks_set: iloc %d", iloc);
    MGB_REQUIRE(color >= 0 && color < m->ncolors, "mgb_masks_set: color %d", color);
    MGB_REQUIRE(state >= -1 && state <= 2, "mgb_masks_set: state %d", state);
    mgb_mask_level& L = m->lev[level];
    int& o            = L.off[(size_t)color * m->subdivx + iloc];
    m->committed      = false;
    if (state <= 0)
        o = -2;
    else if (state == 1)
        o = -1;
    else
    {
        MGB_REQUIRE(values_host, "mgb_masks_set: state 2 needs values");
        MGB_REQUIRE(dtype == MGB_F32 || dtype == MGB_F64, "mgb_masks_set: bad dtype");
        if (o < 0)
        {
            o = L.nslabs++;
            L.values.resize((size_t)L.nslabs * L.slab);
        }
        double* dst = L.values.data() + (size_t)o * L.slab;
        if (dtype == MGB_F64)
            memcpy(dst, values_host, sizeof(double) * L.slab);
        else
            for (long long i = 0; i < L.slab; i++)
                dst[i] = (double)((const float*)values_host)[i];
    }
    return MGB_OK;
}

int mgb_masks_commit(mgb_masks* m)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(m, "mgb_masks_commit: null handle");
    for (int l = 0; l < m->nlevels; l++)
    {
        mgb_mask_level& L = m->lev[l];
        free_level(L);
        MGB_CUDA(cudaMalloc(&L.off_dev, sizeof(int) * L.off.size()));
        MGB_CUDA(cudaMemcpy(
            L.off_dev, L.off.data(), sizeof(int) * L.off.size(), cudaMemcpyHostToDevice));
        // slabs that were replaced by a zero / one state stay in the pool
        // unused; sets are rebuilt when localization regions move
        const size_t n = (size_t)(L.nslabs > 0 ? L.nslabs : 1) * L.slab;
        L.values.resize(n);
        std::vector<float> f(n);
        for (size_t i = 0; i < n; i++)
            f[i] = (float)L.values[i];
        MGB_CUDA(cudaMalloc(&L.pool_f, sizeof(float) * n));
        MGB_CUDA(cudaMemcpy(L.pool_f, f.data(), sizeof(float) * n, cudaMemcpyHostToDevice));
    }
    m->committed = true;
    return MGB_OK;
}

int mgb_masks_destroy(mgb_masks* m)
{
    if (!m) return MGB_OK;
    for (auto& L : m->lev)
        free_level(L);
    delete m;
    return MGB_OK;
}

int mgb_gfv_app_mask(int dtype, const mgb_masks* m, int level, int ghosts, void* ghosted,
    int nfunc, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(ghosted, "mgb_gfv_app_mask: null pointer");
    if (!m || nfunc == 0) return MGB_OK; // map2masks_ == nullptr: no mask applied
    MGB_REQUIRE(nfunc > 0 && nfunc <= m->ncolors, "mgb_gfv_app_mask: nfunc %d > ncolors %d",
        nfunc, m->ncolors);
    MGB_REQUIRE(ghosts >= 0 && ghosts < 10, "mgb_gfv_app_mask: ghosts %d", ghosts);
    int rc;
    if (dtype == MGB_F64 && m->committed && level >= 0 && level < m->nlevels)
        if ((rc = ensure_pool64(const_cast<mgb_masks*>(m)->lev[level]))) return rc;
    const MaskView mv = mask_view(m, level, &rc);
    if (rc) return rc;
    const mgb_mask_level& L = m->lev[level];
    const Box b(L.dim[0], L.dim[1], L.dim[2], ghosts);
    const long long per = (long long)(b.ny + 2 * b.g) * b.incy;
    unsigned gx         = (unsigned)((per + 255) / 256);
    if (gx > 64) gx = 64;
    for (int f0 = 0; f0 < nfunc; f0 += 65535)
    {
        const int nf = (nfunc - f0 < 65535) ? nfunc - f0 : 65535;
        MaskView v   = mv;
        v.off += (size_t)f0 * m->subdivx;
        dim3 grid(gx, (unsigned)b.nx, (unsigned)nf);
        if (dtype == MGB_F64)
            k_app_mask_ghosted<double><<<grid, 256, 0, as_stream(stream)>>>(
                b, v, (double*)ghosted + (size_t)f0 * b.sizeg);
        else if (dtype == MGB_F32)
            k_app_mask_ghosted<float><<<grid, 256, 0, as_stream(stream)>>>(
                b, v, (float*)ghosted + (size_t)f0 * b.sizeg);
        else
        {
            set_error("mgb_gfv_app_mask: bad dtype");
            return MGB_EINVAL;
        }
        MGB_LAUNCHED("k_app_mask_ghosted");
    }
    return MGB_OK;
}

int mgb_app_mask(int dtype, const mgb_masks* m, int level, void* noghost, size_t ld,
    int nfunc, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(noghost, "mgb_app_mask: null pointer");
    if (!m || nfunc == 0) return MGB_OK;
    MGB_REQUIRE(nfunc > 0 && nfunc <= m->ncolors, "mgb_app_mask: nfunc %d > ncolors %d", nfunc,
        m->ncolors);
    int rc;
    if (dtype == MGB_F64 && m->committed && level >= 0 && level < m->nlevels)
        if ((rc = ensure_pool64(const_cast<mgb_masks*>(m)->lev[level]))) return rc;
    const MaskView mv = mask_view(m, level, &rc);
    if (rc) return rc;
    const mgb_mask_level& L = m->lev[level];
    const long long plane   = (long long)L.dim[1] * L.dim[2];
    MGB_REQUIRE(ld >= (size_t)(plane * L.dim[0]), "mgb_app_mask: ld too small");
    unsigned gx = (unsigned)((plane + 255) / 256);
    if (gx > 64) gx = 64;
    for (int f0 = 0; f0 < nfunc; f0 += 65535)
    {
        const int nf = (nfunc - f0 < 65535) ? nfunc - f0 : 65535;
        MaskView v   = mv;
        v.off += (size_t)f0 * m->subdivx;
        dim3 grid(gx, (unsigned)L.dim[0], (unsigned)nf);
        if (dtype == MGB_F64)
            k_app_mask_noghost<double><<<grid, 256, 0, as_stream(stream)>>>(
                plane, L.dim[0], v, (double*)noghost + (size_t)f0 * ld, (long long)ld);
        else if (dtype == MGB_F32)
            k_app_mask_noghost<float><<<grid, 256, 0, as_stream(stream)>>>(
                plane, L.dim[0], v, (float*)noghost + (size_t)f0 * ld, (long long)ld);
        else
        {
            set_error("mgb_app_mask: bad dtype");
            return MGB_EINVAL;
        }
        MGB_LAUNCHED("k_app_mask_noghost");
    }
    return MGB_OK;
}

} // extern "C"
