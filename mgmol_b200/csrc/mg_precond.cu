// Block multigrid preconditioner in float: Preconditioning<float>
// (src/Preconditioning.cc:15-216) and OrbitalsPreconditioning<T>::precond_mg
// (src/OrbitalsPreconditioning.cc:87-117), all orbitals of a block at once.
//
// The V-cycle is a fixed sequence of kernels on one stream (no host
// synchronisation); every level keeps its float work blocks resident in HBM
// for the lifetime of the handle, as the reference's gfv_work_/gfv_rcoarse_/
// gfv_newv_ vectors do.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hpsi.h"
#include "masks.h"
#include "mg_fused.h"

struct mgb_precond
{
    int lap_type;
    int max_levels;
    int nfunc_max;
    int g;
    std::vector<mgb_grid> grid;  // per level
    std::vector<double> jf;      // jacobiFactor per level
    std::vector<float*> work;    // gfv_work_[level]
    std::vector<float*> rcoarse; // gfv_rcoarse_[level] (lives on level+1 grid)
    std::vector<float*> newv;    // gfv_newv_[level]
    float* v0;                   // OrbitalsPreconditioning::gfv_work_
    float* f0;                   // OrbitalsPreconditioning::gfv_work2_
    bool literal_ready;          // the ghosted blocks above are allocated
    // fused path (mg_fused.cu): no-ghost float blocks per level
    bool fused_ok;               // every level can run the fused kernels
    int mode;                    // 0 automatic, 1 literal, 2 fused
    int last_mode;               // what the last precond_mg call used (1 / 2)
    bool fused_ready;
    std::vector<float*> fa, fb;  // ping-pong iterates
    std::vector<float*> fw;      // residual of the last pre-smoothing sweep
    std::vector<float*> ff;      // right-hand side (level 0: converted residual)
    const mgb_masks* masks;      // GridFuncVector::map2masks_ (null: none)
    mgb_comm* comm;              // multi-rank boxes: halo exchange / peer reads
    bool peers_ready;            // the fused work blocks are registered
};

namespace mgb
{

static size_t sizeg_of(const mgb_grid& g)
{
    return (size_t)box_of(&g, g.ghosts).sizeg;
}

// pb::Lap constants: src/pb/Laph4M.h:29-31,80; Laph2.h:32-34,95;
// Laph4.h:35-38,106; Laph6.h:40-42,101; Laph8.h:40-42,100
static int lap_constants(int lap_type, const double h[3], double out[3])
{
    const double s = 1. / (h[0] * h[0]) + 1. / (h[1] * h[1]) + 1. / (h[2] * h[2]);
    double diag, jf;
    switch (lap_type)
    {
        case MGB_LAP_4M:
        case MGB_LAP_4MP:
            diag = (4. / 3.) * s;
            jf   = 1. / diag;
            break;
        case MGB_LAP_2:
            diag = 2. * s;
            jf   = 6. * (1. / diag) / 7.;
            break;
        case MGB_LAP_4:
            diag = 2.5 * s;
            jf   = (1. / diag) / 1.5;
            break;
        case MGB_LAP_6:
            diag = (49. / 18.) * s;
            jf   = (1. / diag) / 1.5;
            break;
        case MGB_LAP_8:
            diag = (1435. / 504.) * s;
            jf   = (1. / diag) / 1.5;
            break;
        default:
            set_error("lap_type %d: LapFactory::createLap() option invalid", lap_type);
            return MGB_EINVAL;
    }
    out[0] = diag;
    out[1] = 1. / diag;
    out[2] = jf;
    return MGB_OK;
}

static bool multi_rank(const mgb_grid& g)
{
    return g.nproc[0] * g.nproc[1] * g.nproc[2] > 1;
}

// GridFuncVector::trade_boundaries: local wraps on one rank, the Y -> Z -> X
// exchange over the communicator otherwise
static int trade(mgb_precond* p, const mgb_grid& gr, float* u, int nfunc, cudaStream_t st)
{
    if (multi_rank(gr))
        return mgb_halo_exchange_ghosted(p->comm, MGB_F32, &gr, u, nfunc, (void*)st);
    return mgb_gfv_trade_boundaries(MGB_F32, &gr, u, nfunc, (void*)st);
}

// GridFuncVector::jacobi with the updated_boundaries_ bookkeeping of the
// reference: applyLap trades v's boundaries unless they are flagged current.
static int jacobi(mgb_precond* p, int lap_type, int level, float* v, bool& v_upd,
    const float* f, int nfunc, cudaStream_t st)
{
    const mgb_grid& gr = p->grid[level];
    if (!v_upd)
        if (int rc = trade(p, gr, v, nfunc, st)) return rc;
    if (int rc = jacobi_literal(lap_type, &gr, v, f, p->work[level], nfunc,
            p->jf[level], st))
        return rc;
    v_upd = false;
    // gfv_v.app_mask(level)  (Preconditioning.cc:176,212)
    return mgb_gfv_app_mask(MGB_F32, p->masks, level, gr.ghosts, v, nfunc, (void*)st);
}

// Preconditioning<float>::mg, src/Preconditioning.cc:155-216
static int vcycle(mgb_precond* p, float* v, bool& v_upd, const float* f,
    int lap_type, int level, int nfunc, cudaStream_t st)
{
    const mgb_grid& gr = p->grid[level];
    const int ncycl    = (level == p->max_levels) ? 4 : 2;
    for (int it = 0; it < ncycl; it++) // :173-177
        if (int rc = jacobi(p, lap_type, level, v, v_upd, f, nfunc, st)) return rc;
    if (level == p->max_levels) return MGB_OK; // :179

    float* w = p->work[level];
    // :184 LOCALIZATION of the residual
    if (int rc = mgb_gfv_app_mask(MGB_F32, p->masks, level, gr.ghosts, w, nfunc, (void*)st))
        return rc;
    // :189 restrict3D trades w first (GridFuncVector.cc:1624-1631)
    if (int rc = trade(p, gr, w, nfunc, st)) return rc;
    float* rc_ = p->rcoarse[level];
    if (int rc = mgb_gfv_restrict3D(MGB_F32, &gr, w, rc_, nfunc, (void*)st)) return rc;
    // :192
    if (int rc = mgb_gfv_app_mask(
            MGB_F32, p->masks, level + 1, gr.ghosts, rc_, nfunc, (void*)st))
        return rc;

    float* nv            = p->newv[level];
    const mgb_grid& cgr  = p->grid[level + 1];
    MGB_CUDA(cudaMemsetAsync(nv, 0, sizeof(float) * sizeg_of(cgr) * nfunc, st)); // :198
    bool nv_upd = true; // resetData() sets the flag
    if (int rc = vcycle(p, nv, nv_upd, rc_, MGB_LAP_2, level + 1, nfunc, st)) // :199
        return rc;

    // :201 extend3D trades the coarse block first (GridFuncVector.cc:1633-1641)
    if (!nv_upd)
        if (int rc = trade(p, cgr, nv, nfunc, st)) return rc;
    if (int rc = mgb_gfv_extend3D(MGB_F32, &gr, nv, w, nfunc, (void*)st)) return rc;
    // :204
    if (int rc = mgb_gfv_app_mask(MGB_F32, p->masks, level, gr.ghosts, w, nfunc, (void*)st))
        return rc;

    // :206  v -= w
    if (int rc = mgb_axpy(MGB_F32, sizeg_of(gr) * nfunc, -1., w, v, (void*)st))
        return rc;
    v_upd = false;

    for (int it = 0; it < 2; it++) // :209-213
        if (int rc = jacobi(p, lap_type, level, v, v_upd, f, nfunc, st)) return rc;

    // :215 (the reference tests bc_[0], bc_[2], bc_[2])
    if (gr.bc[0] != 1 || gr.bc[2] != 1 || gr.bc[2] != 1)
    {
        if (!v_upd)
            if (int rc = trade(p, gr, v, nfunc, st)) return rc;
        v_upd = true;
    }
    return MGB_OK;
}

} // namespace mgb

using namespace mgb;

extern "C"
{

int mgb_lap_constants(int lap_type, const double h[3], double out[3])
{
    MGB_REQUIRE(h && out, "mgb_lap_constants: null pointer");
    return lap_constants(lap_type, h, out);
}

double mgb_gamma(double inv_diag, int mg_levels, double vmax, double small_eig)
{
    // OrbitalsPreconditioning<T>::setGamma, src/OrbitalsPreconditioning.cc:
    // 120-145
    double gamma = inv_diag;
    for (int ln = 0; ln <= mg_levels; ln++)
        gamma *= 4.;
    return 1.0 / (2.0 / gamma + fabs(vmax - small_eig));
}

int mgb_precond_create(mgb_precond** out, int lap_type, int mg_levels,
    const mgb_grid* grid, int nfunc_max)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(grid)) return rc;
    MGB_REQUIRE(out, "mgb_precond_create: null out pointer");
    MGB_REQUIRE(mg_levels >= 0 && mg_levels < 12, "bad mg_levels %d", mg_levels);
    MGB_REQUIRE(nfunc_max > 0, "nfunc_max <= 0");
    // GridFuncVector::applyLap has no case 10 and aborts
    // (src/pb/GridFuncVector.cc:2392-2396)
    MGB_REQUIRE(lap_type == MGB_LAP_4M || lap_type == MGB_LAP_2
                    || lap_type == MGB_LAP_4 || lap_type == MGB_LAP_6
                    || lap_type == MGB_LAP_8,
        "mgb_precond_create: lap_type %d has no batched applyLap", lap_type);
    const int gmin = (lap_type == MGB_LAP_4)   ? 2
                     : (lap_type == MGB_LAP_6) ? 3
                     : (lap_type == MGB_LAP_8) ? 4
                                               : 1;
    MGB_REQUIRE(grid->ghosts >= gmin,
        "mgb_precond_create: lap_type %d needs >= %d ghosts", lap_type, gmin);
    // local dims must be divisible by 2^levels (src/mgmol_run.cc:108-120)
    for (int d = 0; d < 3; d++)
        MGB_REQUIRE(grid->dim[d] % (1 << mg_levels) == 0,
            "mgb_precond_create: dim[%d]=%d not divisible by 2^%d", d,
            grid->dim[d], mg_levels);

    mgb_precond* p = new mgb_precond();
    p->lap_type    = lap_type;
    p->max_levels  = mg_levels;
    p->nfunc_max   = nfunc_max;
    p->g           = grid->ghosts;
    p->v0 = p->f0 = nullptr;
    p->literal_ready = p->fused_ready = false;
    p->mode = p->last_mode = 0;
    p->masks = nullptr;
    p->comm  = nullptr;
    p->peers_ready = false;
    if (const char* env = getenv("MGB_MG_MODE")) p->mode = atoi(env);
    mgb_grid g = *grid;
    int rc     = MGB_OK;
    // Mixed periodic/Dirichlet boxes stay on the literal path: there the
    // reference's zeroing skips the first x/y layers "to avoid setting values
    // twice" (src/pb/GridFunc.cc:2222-2236), so its result depends on stale
    // ghost values that only the ghosted layout reproduces.
    p->fused_ok = (grid->bc[0] == grid->bc[1] && grid->bc[1] == grid->bc[2]);
    // the fused kernels read x neighbours in place; y / z splits go through
    // the ghosted exchange of the literal sequence
    // decomposed boxes: the neighbours' blocks must have my shape (halos are read in place)
    for (int d = 0; d < 3; d++)
        if (grid->nproc[d] > 1 && grid->dim[d] * grid->nproc[d] != grid->gdim[d])
            p->fused_ok = false;
    // y / z splits take the in-place kernels of any decomposition, which need whole
    // 16-byte z-halo columns on every level (nz of the coarsest level a multiple of 4 is
    // checked per level below) and do not translate color slots
    if (getenv("MGB_MG_NO_3D") && (grid->nproc[1] != 1 || grid->nproc[2] != 1)) p->fused_ok = false;
    for (int l = 0; l <= mg_levels && rc == MGB_OK; l++)
    {
        if (l > 0)
            for (int d = 0; d < 3; d++)
            {
                // Grid::coarse_grid (src/pb/Grid.cc:214-231): half the points,
                // same ghost width, same cell -> doubled spacing
                g.dim[d] /= 2;
                g.gdim[d] /= 2;
                g.h[d] *= 2.;
            }
        p->grid.push_back(g);
        double c[3];
        // Preconditioning.cc:26-27 (level 0: lap_type), :122-123 (coarse: 1)
        rc = lap_constants(l == 0 ? lap_type : MGB_LAP_2, g.h, c);
        p->jf.push_back(c[2]);
        if (!mg_fused_level_ok(g, l == 0 ? lap_type : MGB_LAP_2)) p->fused_ok = false;
    }
    if (rc != MGB_OK)
    {
        delete p;
        return rc;
    }
    *out = p;
    return MGB_OK;
}

} // extern "C"

namespace mgb
{
// zero-filled like a freshly allocated GridFuncVector
// (src/pb/GridFuncVector.cc:22-43): with mixed boundary conditions the
// reference's result depends on ghost values that are never written again
static int dev_alloc(float** q, size_t bytes)
{
    if (cudaMalloc(q, bytes) != cudaSuccess)
    {
        *q = nullptr;
        set_error("mgb_precond: device allocation of %zu bytes failed", bytes);
        (void)cudaGetLastError();
        return MGB_ECUDA;
    }
    MGB_CUDA(cudaMemset(*q, 0, bytes));
    return MGB_OK;
}

// ghosted work blocks of the literal path (gfv_work_[l], gfv_rcoarse_[l],
// gfv_newv_[l], and OrbitalsPreconditioning's two level-0 blocks)
static int ensure_literal(mgb_precond* p)
{
    if (p->literal_ready) return MGB_OK;
    for (int l = 0; l <= p->max_levels; l++)
    {
        const size_t bytes = sizeof(float) * sizeg_of(p->grid[l]) * p->nfunc_max;
        float *w = nullptr, *r = nullptr, *n = nullptr;
        if (int rc = dev_alloc(&w, bytes)) return rc;
        p->work.push_back(w);
        if (l > 0)
        {
            if (int rc = dev_alloc(&r, bytes)) return rc;
            p->rcoarse.push_back(r);
            if (int rc = dev_alloc(&n, bytes)) return rc;
            p->newv.push_back(n);
        }
    }
    const size_t bytes0 = sizeof(float) * sizeg_of(p->grid[0]) * p->nfunc_max;
    if (int rc = dev_alloc(&p->v0, bytes0)) return rc;
    if (int rc = dev_alloc(&p->f0, bytes0)) return rc;
    p->literal_ready = true;
    return MGB_OK;
}

static size_t npt_of(const mgb_grid& g)
{
    return (size_t)g.dim[0] * g.dim[1] * g.dim[2];
}

static int ensure_fused(mgb_precond* p)
{
    if (p->fused_ready) return MGB_OK;
    for (int l = 0; l <= p->max_levels; l++)
    {
        const size_t bytes = sizeof(float) * npt_of(p->grid[l]) * p->nfunc_max;
        float *a = nullptr, *b = nullptr, *w = nullptr, *f = nullptr;
        if (int rc = dev_alloc(&a, bytes)) return rc;
        p->fa.push_back(a);
        if (int rc = dev_alloc(&b, bytes)) return rc;
        p->fb.push_back(b);
        if (l < p->max_levels)
            if (int rc = dev_alloc(&w, bytes)) return rc;
        p->fw.push_back(w);
        p->ff.push_back(f); // level 0: allocated on the first double-precision call
        // (x-split domains: always, so that it is registered with the others)
        if (l > 0 || multi_rank(p->grid[0]))
            if (int rc = dev_alloc(&p->ff[l], bytes)) return rc;
    }
    p->fused_ready = true;
    return MGB_OK;
}

// x-split domains: publish every fused work block so that the neighbours'
// kernels can read boundary planes in place (collective, once per handle)
static int ensure_peers(mgb_precond* p, cudaStream_t st)
{
    if (p->peers_ready || !multi_rank(p->grid[0])) return MGB_OK;
    for (int l = 0; l <= p->max_levels; l++)
        for (float* q : { p->fa[l], p->fb[l], p->fw[l], p->ff[l] })
            if (q)
                if (int rc = mgb_peer_register(p->comm, q, (void*)st)) return rc;
    p->peers_ready = true;
    return MGB_OK;
}

struct XPeers
{
    const float* w;
    const float* e;
};
// the west / east neighbours' copies of a registered work block (null where
// the domain ends, or on a single rank)
static int x_peers(mgb_precond* p, const mgb_grid& gr, const float* q, XPeers& out)
{
    out.w = out.e = nullptr;
    if (gr.nproc[0] == 1) return MGB_OK;
    const bool per = gr.bc[0] == 1;
    if (per || gr.coord[0] > 0)
    {
        out.w = (const float*)peer_view(
            p->comm, q, comm_rank_of(&gr, gr.coord[0] - 1, gr.coord[1], gr.coord[2]));
        if (!out.w) goto fail;
    }
    if (per || gr.coord[0] < gr.nproc[0] - 1)
    {
        out.e = (const float*)peer_view(
            p->comm, q, comm_rank_of(&gr, gr.coord[0] + 1, gr.coord[1], gr.coord[2]));
        if (!out.e) goto fail;
    }
    return MGB_OK;
fail:
    set_error("mgb_precond: a neighbour's work block cannot be mapped (CUDA IPC)");
    return MGB_ENOTSUP;
}

// y / z splits: the copies of a registered work block on the 27 Cartesian neighbours (my
// own block across a direction that is not split)
static bool use_3d(const mgb_grid& gr)
{
    // MGB_MG_FORCE_3D: test hook -- a single rank runs the any-decomposition kernels with
    // itself as every neighbour
    return gr.nproc[1] > 1 || gr.nproc[2] > 1 || getenv("MGB_MG_FORCE_3D") != nullptr;
}
static int nb_peers(mgb_precond* p, const mgb_grid& gr, const float* q, const float* out[27])
{
    if (!multi_rank(gr))
    {
        for (int i = 0; i < 27; i++)
            out[i] = q;
        return MGB_OK;
    }
    for (int dx = -1; dx <= 1; dx++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dz = -1; dz <= 1; dz++)
            {
                const int i = ((dx + 1) * 3 + (dy + 1)) * 3 + (dz + 1);
                out[i]      = (const float*)peer_view(p->comm, q,
                    comm_rank_of(&gr, gr.coord[0] + dx, gr.coord[1] + dy, gr.coord[2] + dz));
                if (!out[i])
                {
                    set_error("mgb_precond: a neighbour's work block cannot be mapped (CUDA IPC)");
                    return MGB_ENOTSUP;
                }
            }
    return MGB_OK;
}

// One level of Preconditioning<float>::mg (src/Preconditioning.cc:155-216) on
// no-ghost blocks with the fused kernels.  The start vector of the level is
// s * f: gamma * res at level 0 (OrbitalsPreconditioning.cc:104) and, on the
// coarse levels, the result omega * f of the first sweep from v = 0; `npre` is
// the number of pre-smoothing sweeps still to do.  The last sweep of level 0
// writes the caller's block (fout / fout64, leading dimension ldo).
static int cycle_fused(mgb_precond* p, int l, const float* f, size_t ldf, double s,
    int npre, float* fout, double* fout64, size_t ldo, int nfunc, float** result,
    cudaStream_t st)
{
    const mgb_grid& gr = p->grid[l];
    const size_t ld    = npt_of(gr);
    const int lap      = (l == 0) ? p->lap_type : MGB_LAP_2;
    const bool coarsest = (l == p->max_levels);
    const bool periodic = gr.bc[0] == 1 && gr.bc[1] == 1 && gr.bc[2] == 1;
    int zl[3], nozero[3] = { 0, 0, 0 };
    for (int d = 0; d < 3; d++)
        zl[d] = gr.bc[d] != 1 && gr.coord[d] == 0; // the rank owning the low face
    const bool split = multi_rank(gr);
    const bool d3    = use_3d(gr);
    const int *map_w = nullptr, *map_e = nullptr;
    int map_n = 0;
    if (split) comm_color_maps(p->comm, &map_w, &map_e, &map_n);
    MGB_REQUIRE(!(d3 && map_w),
        "mgb_precond_mg: color-slot translation (LocGridOrbitals) is served on x slabs only");
    MGB_REQUIRE(!map_w || map_n >= nfunc, "mgb_precond_mg: color maps cover %d colors, need %d",
        map_n, nfunc);
    // the last sweep of level 0 is followed by a trade only on the path through
    // Preconditioning.cc:215, whose test reads bc_[0], bc_[2], bc_[2]
    const bool final_trade = !coarsest && (gr.bc[0] != 1 || gr.bc[2] != 1);
    int mrc = MGB_OK;
    const MaskView mk  = mask_view(p->masks, l, &mrc);
    if (mrc) return mrc;
    const MaskView mkc = coarsest ? no_mask() : mask_view(p->masks, l + 1, &mrc);
    if (mrc) return mrc;
    float* cur = nullptr;
    bool pending_scale = true;
    // The start vector can be formed while the first sweep loads its tiles
    // unless low layers must be stored as zeros (Dirichlet) or, on a coarse
    // level, the start vector omega * f is itself masked (it is the result of
    // the reference's first sweep, followed by app_mask): then it is stored.
    if (!periodic || (l > 0 && mk.off))
    {
        if (int rc = mg_scale(gr, s, f, ldf, p->fa[l], ld, nfunc, zl, l > 0 ? mk : no_mask(), st))
            return rc;
        cur           = p->fa[l];
        pending_scale = false;
    }
    auto sweep = [&](bool emit_w, bool final_out) -> int {
        MgJacobiArgs a;
        a.lap_type = lap;
        a.grid     = &gr;
        a.in       = pending_scale ? f : cur;
        a.ld_in    = pending_scale ? ldf : ld;
        a.scale    = pending_scale ? s : 1.0;
        a.f        = f;
        a.ld_f     = ldf;
        float* nxt = (cur == p->fa[l]) ? p->fb[l] : p->fa[l];
        a.out      = final_out ? fout : nxt;
        a.out64    = final_out ? fout64 : nullptr;
        a.ld_out   = final_out ? ldo : ld;
        a.w        = emit_w ? p->fw[l] : nullptr;
        a.ld_w     = ld;
        a.nfunc    = nfunc;
        a.omega    = p->jf[l];
        const int* z = (final_out && !final_trade) ? nozero : zl;
        for (int d = 0; d < 3; d++)
            a.zero_low[d] = z[d];
        a.mask = mk;
        XPeers xp;
        const float* nb[27];
        a.nb3d = nullptr;
        const void* nbz[9];
        a.nbz = nullptr;
        if (d3)
        {
            if (int rc = nb_peers(p, gr, a.in, nb)) return rc;
            a.nb3d = nb;
            xp.w = xp.e = nullptr;
            if (gr.nproc[2] > 1 || getenv("MGB_HPSI_FORCE_ZBOXES"))
            {
                // z split: nobody is still reading its buffer of pushed columns, then the
                // source's z-edge columns go to the z neighbours
                if (split)
                    if (int rc = comm_barrier_neighbors(p->comm, &gr, st)) return rc;
                if (int rc = comm_zhalo(p->comm, &gr, 4, a.in, a.ld_in, nfunc, st, nbz)) return rc;
                a.nbz = nbz;
            }
        }
        else if (int rc = x_peers(p, gr, a.in, xp))
            return rc;
        a.peer_w = xp.w;
        a.peer_e = xp.e;
        a.map_w  = map_w;
        a.map_e  = map_e;
        // every rank has finished writing (and reading) the blocks involved
        if (split)
            if (int rc = comm_barrier_neighbors(p->comm, &gr, st)) return rc;
        if (int rc = mg_jacobi(a, st)) return rc;
        cur           = final_out ? nullptr : nxt;
        pending_scale = false;
        return MGB_OK;
    };
    for (int it = 0; it < npre; it++) // :173-177
    {
        const bool last = (it == npre - 1);
        if (int rc = sweep(last && !coarsest, last && coarsest && l == 0)) return rc;
    }
    if (coarsest) // :179
    {
        *result = cur;
        return MGB_OK;
    }
    // :189-192 restriction of the residual of the last pre-smoothing sweep
    {
        XPeers xp;
        const float* nb[27];
        if (d3)
        {
            if (int rc = nb_peers(p, gr, p->fw[l], nb)) return rc;
            xp.w = xp.e = nullptr;
        }
        else if (int rc = x_peers(p, gr, p->fw[l], xp))
            return rc;
        if (split)
            if (int rc = comm_barrier_neighbors(p->comm, &gr, st)) return rc;
        if (int rc = mg_restrict(gr, p->fw[l], ld, p->ff[l + 1], npt_of(p->grid[l + 1]), nfunc, mkc,
                xp.w, map_w, st, d3 ? nb : nullptr))
            return rc;
    }
    // :198-199 coarse correction from a zero start: its first sweep gives
    // omega * f, folded into the start vector
    const int ncycl_c = (l + 1 == p->max_levels) ? 4 : 2;
    float* e          = nullptr;
    if (int rc = cycle_fused(p, l + 1, p->ff[l + 1], npt_of(p->grid[l + 1]), p->jf[l + 1],
            ncycl_c - 1, nullptr, nullptr, 0, nfunc, &e, st))
        return rc;
    // :201-206 v -= P e
    {
        XPeers xp;
        const float* nb[27];
        if (d3)
        {
            if (int rc = nb_peers(p, p->grid[l + 1], e, nb)) return rc;
            xp.w = xp.e = nullptr;
        }
        else if (int rc = x_peers(p, p->grid[l + 1], e, xp))
            return rc;
        if (split)
            if (int rc = comm_barrier_neighbors(p->comm, &gr, st)) return rc;
        if (int rc = mg_prolong_correct(gr, e, npt_of(p->grid[l + 1]), cur, ld, nfunc, zl, mk, xp.e,
                map_e, st, d3 ? nb : nullptr))
            return rc;
    }
    for (int it = 0; it < 2; it++) // :209-213
        if (int rc = sweep(false, it == 1 && l == 0)) return rc;
    *result = cur;
    return MGB_OK;
}
} // namespace mgb

extern "C"
{

int mgb_precond_destroy(mgb_precond* p)
{
    if (!p) return MGB_OK;
    if (p->peers_ready && p->comm)
    {
        // collective: nobody reads my work blocks any more, every rank closes its
        // mappings of its neighbours' blocks, and only then are they freed
        comm_barrier(p->comm, nullptr);
        cudaDeviceSynchronize();
        for (auto* vec : { &p->fa, &p->fb, &p->fw, &p->ff })
            for (float* q : *vec)
                if (q) mgb_peer_unregister(p->comm, q);
        comm_barrier(p->comm, nullptr);
        cudaDeviceSynchronize();
    }
    for (float* q : p->work)
        if (q) cudaFree(q);
    for (float* q : p->rcoarse)
        if (q) cudaFree(q);
    for (float* q : p->newv)
        if (q) cudaFree(q);
    if (p->v0) cudaFree(p->v0);
    if (p->f0) cudaFree(p->f0);
    for (auto* vec : { &p->fa, &p->fb, &p->fw, &p->ff })
        for (float* q : *vec)
            if (q) cudaFree(q);
    delete p;
    return MGB_OK;
}

int mgb_precond_set_masks(mgb_precond* p, const mgb_masks* m)
{
    MGB_REQUIRE(p, "mgb_precond_set_masks: null handle");
    if (m)
    {
        MGB_REQUIRE(masks_match(m, &p->grid[0]),
            "mgb_precond_set_masks: mask set built for another grid");
        MGB_REQUIRE(masks_nlevels(m) >= p->max_levels + 1,
            "mgb_precond_set_masks: mask set has %d levels, the V-cycle needs %d",
            masks_nlevels(m), p->max_levels + 1);
        MGB_REQUIRE(masks_ncolors(m) >= p->nfunc_max,
            "mgb_precond_set_masks: mask set has %d colors, the block up to %d",
            masks_ncolors(m), p->nfunc_max);
    }
    p->masks = m;
    return MGB_OK;
}

int mgb_precond_set_comm(mgb_precond* p, mgb_comm* comm)
{
    MGB_REQUIRE(p, "mgb_precond_set_comm: null handle");
    p->comm = comm;
    return MGB_OK;
}

int mgb_precond_set_mode(mgb_precond* p, int mode)
{
    MGB_REQUIRE(p, "mgb_precond_set_mode: null handle");
    MGB_REQUIRE(mode >= 0 && mode <= 2, "mgb_precond_set_mode: mode %d", mode);
    MGB_REQUIRE(mode != 2 || p->fused_ok,
        "mgb_precond_set_mode: the fused kernels cannot run this grid");
    p->mode = mode;
    return MGB_OK;
}

int mgb_precond_last_mode(mgb_precond* p) { return p ? p->last_mode : 0; }

int mgb_precond_vcycle(mgb_precond* p, float* v, const float* f, int nfunc, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(p && v && f, "mgb_precond_vcycle: null pointer");
    MGB_REQUIRE(nfunc >= 0 && nfunc <= p->nfunc_max, "nfunc %d > nfunc_max %d",
        nfunc, p->nfunc_max);
    if (nfunc == 0) return MGB_OK;
    if (int rc = ensure_literal(p)) return rc;
    bool v_upd = false;
    return vcycle(p, v, v_upd, f, p->lap_type, 0, nfunc, as_stream(stream));
}

int mgb_precond_mg(mgb_precond* p, int dtype, void* res, size_t ld, int nfunc,
    double gamma, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(p && res, "mgb_precond_mg: null pointer");
    MGB_REQUIRE(dtype == MGB_F32 || dtype == MGB_F64, "bad dtype");
    MGB_REQUIRE(nfunc >= 0 && nfunc <= p->nfunc_max, "nfunc %d > nfunc_max %d",
        nfunc, p->nfunc_max);
    MGB_REQUIRE(gamma > 0., "gamma must be > 0 (assert in precond_mg)");
    if (nfunc == 0) return MGB_OK;
    const mgb_grid& gr = p->grid[0];
    cudaStream_t st    = as_stream(stream);
    int rc;
    const size_t es = dtype == MGB_F64 ? 8 : 4;
    MGB_REQUIRE(!multi_rank(gr) || p->comm,
        "mgb_precond_mg: the box is one of %d x %d x %d ranks but no communicator was "
        "attached (mgb_precond_set_comm)",
        gr.nproc[0], gr.nproc[1], gr.nproc[2]);
    bool can_fuse = p->fused_ok && ld % 4 == 0 && ((uintptr_t)res & 15) == 0
                    && ld * es % 16 == 0;
    if (can_fuse && multi_rank(gr))
    {
        // x-split: the neighbours' work blocks must be mappable (CUDA IPC);
        // otherwise the literal sequence with the packed exchange serves
        if ((rc = ensure_fused(p))) return rc;
        rc = ensure_peers(p, st);
        if (rc == MGB_ENOTSUP)
            can_fuse = false;
        else if (rc)
            return rc;
    }
    MGB_REQUIRE(p->mode != 2 || can_fuse,
        "mgb_precond_mg: fused mode forced but this grid/block is not eligible");
    if (p->mode == 2 || (p->mode == 0 && can_fuse))
    {
        if ((rc = ensure_fused(p))) return rc;
        // nobody is still reading my work blocks from the previous call
        if (multi_rank(gr))
            if ((rc = comm_barrier_neighbors(p->comm, &gr, st))) return rc;
        const float* f = (const float*)res;
        size_t ldf     = ld;
        if (dtype == MGB_F64)
        {
            // orbitals.setDataWithGhosts(gfv_work2_): ORBDTYPE -> float  (:103)
            if (!p->ff[0])
                if ((rc = dev_alloc(&p->ff[0], sizeof(float) * npt_of(gr) * p->nfunc_max)))
                    return rc;
            if ((rc = mg_convert(npt_of(gr), (const double*)res, ld, p->ff[0], npt_of(gr),
                     nfunc, st)))
                return rc;
            f   = p->ff[0];
            ldf = npt_of(gr);
        }
        else if (multi_rank(gr))
        {
            // decomposed box: the first sweep reads the neighbours' planes of f in place,
            // so f must live in a registered block (same kernels, hence the same
            // bits, as on a single rank)
            MGB_CUDA(cudaMemcpy2DAsync(p->ff[0], npt_of(gr) * sizeof(float), res,
                ld * sizeof(float), npt_of(gr) * sizeof(float), (size_t)nfunc,
                cudaMemcpyDeviceToDevice, st));
            f   = p->ff[0];
            ldf = npt_of(gr);
        }
        float* unused  = nullptr;
        const int ncyc = (p->max_levels == 0) ? 4 : 2;
        p->last_mode   = 2;
        return cycle_fused(p, 0, f, ldf, gamma, ncyc, dtype == MGB_F32 ? (float*)res : nullptr,
            dtype == MGB_F64 ? (double*)res : nullptr, ld, nfunc, &unused, st);
    }
    p->last_mode = 1;
    if ((rc = ensure_literal(p))) return rc;
    const size_t n = sizeg_of(gr) * nfunc;
    // gfv_work_->resetData()                           (OrbitalsPrecond.cc:99)
    MGB_CUDA(cudaMemsetAsync(p->v0, 0, sizeof(float) * n, st));
    // orbitals.setDataWithGhosts(gfv_work2_)  ORBDTYPE -> float        (:103)
    if ((rc = mgb_gfv_set_with_ghosts(dtype, MGB_F32, &gr, res, ld, p->f0, nfunc, stream)))
        return rc;
    // gfv_work_->axpy(gamma_, *gfv_work2_)                             (:104)
    if ((rc = mgb_axpy(MGB_F32, n, gamma, p->f0, p->v0, stream))) return rc;
    // precond_->mg(*gfv_work_, *gfv_work2_, lap_type_, 0)              (:107)
    bool v_upd = false;
    if ((rc = vcycle(p, p->v0, v_upd, p->f0, p->lap_type, 0, nfunc, st))) return rc;
    // orbitals.setPsi(*gfv_work_)  float -> ORBDTYPE                   (:109)
    return mgb_gfv_get_values(MGB_F32, dtype, &gr, p->v0, res, ld, nfunc, stream);
}

} // extern "C"
