// Block multigrid preconditioner in float: Preconditioning<float>
// (src/Preconditioning.cc:15-216) and OrbitalsPreconditioning<T>::precond_mg
// (src/OrbitalsPreconditioning.cc:87-117), all orbitals of a block at once.
//
// The V-cycle is a fixed sequence of kernels on one stream (no host
// synchronisation); every level keeps its float work blocks resident in HBM
// for the lifetime of the handle, as the reference's gfv_work_/gfv_rcoarse_/
// gfv_newv_ vectors do.
#include <cmath>
#include <cstring>
#include <vector>

#include "hpsi.h"

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

// GridFuncVector::jacobi with the updated_boundaries_ bookkeeping of the
// reference: applyLap trades v's boundaries unless they are flagged current.
static int jacobi(mgb_precond* p, int lap_type, int level, float* v, bool& v_upd,
    const float* f, int nfunc, cudaStream_t st)
{
    const mgb_grid& gr = p->grid[level];
    if (!v_upd)
        if (int rc = mgb_gfv_trade_boundaries(MGB_F32, &gr, v, nfunc, (void*)st))
            return rc;
    if (int rc = jacobi_literal(lap_type, &gr, v, f, p->work[level], nfunc,
            p->jf[level], st))
        return rc;
    v_upd = false;
    return MGB_OK;
}

// Preconditioning<float>::mg, src/Preconditioning.cc:155-216 (no masks)
static int vcycle(mgb_precond* p, float* v, bool& v_upd, const float* f,
    int lap_type, int level, int nfunc, cudaStream_t st)
{
    const mgb_grid& gr = p->grid[level];
    const int ncycl    = (level == p->max_levels) ? 4 : 2;
    for (int it = 0; it < ncycl; it++) // :173-177
        if (int rc = jacobi(p, lap_type, level, v, v_upd, f, nfunc, st)) return rc;
    if (level == p->max_levels) return MGB_OK; // :179

    float* w = p->work[level];
    // :189 restrict3D trades w first (GridFuncVector.cc:1624-1631)
    if (int rc = mgb_gfv_trade_boundaries(MGB_F32, &gr, w, nfunc, (void*)st)) return rc;
    float* rc_ = p->rcoarse[level];
    if (int rc = mgb_gfv_restrict3D(MGB_F32, &gr, w, rc_, nfunc, (void*)st)) return rc;

    float* nv            = p->newv[level];
    const mgb_grid& cgr  = p->grid[level + 1];
    MGB_CUDA(cudaMemsetAsync(nv, 0, sizeof(float) * sizeg_of(cgr) * nfunc, st)); // :198
    bool nv_upd = true; // resetData() sets the flag
    if (int rc = vcycle(p, nv, nv_upd, rc_, MGB_LAP_2, level + 1, nfunc, st)) // :199
        return rc;

    // :201 extend3D trades the coarse block first (GridFuncVector.cc:1633-1641)
    if (!nv_upd)
        if (int rc = mgb_gfv_trade_boundaries(MGB_F32, &cgr, nv, nfunc, (void*)st))
            return rc;
    if (int rc = mgb_gfv_extend3D(MGB_F32, &gr, nv, w, nfunc, (void*)st)) return rc;

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
            if (int rc = mgb_gfv_trade_boundaries(MGB_F32, &gr, v, nfunc, (void*)st))
                return rc;
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
    MGB_REQUIRE(grid->nproc[0] == 1 && grid->nproc[1] == 1 && grid->nproc[2] == 1,
        "mgb_precond_create: multi-rank V-cycle is driven through "
        "mgb_halo_exchange_ghosted by the host wrapper");

    mgb_precond* p = new mgb_precond();
    p->lap_type    = lap_type;
    p->max_levels  = mg_levels;
    p->nfunc_max   = nfunc_max;
    p->g           = grid->ghosts;
    p->v0 = p->f0 = nullptr;
    mgb_grid g    = *grid;
    int rc        = MGB_OK;
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
        const size_t bytes = sizeof(float) * sizeg_of(g) * nfunc_max;
        float* w           = nullptr;
        if (rc == MGB_OK && cudaMalloc(&w, bytes) != cudaSuccess) rc = MGB_ECUDA;
        p->work.push_back(w);
        if (l > 0)
        {
            float *r = nullptr, *n = nullptr;
            if (rc == MGB_OK && cudaMalloc(&r, bytes) != cudaSuccess) rc = MGB_ECUDA;
            if (rc == MGB_OK && cudaMalloc(&n, bytes) != cudaSuccess) rc = MGB_ECUDA;
            p->rcoarse.push_back(r);
            p->newv.push_back(n);
        }
    }
    const size_t bytes0 = sizeof(float) * sizeg_of(p->grid[0]) * nfunc_max;
    if (rc == MGB_OK && cudaMalloc(&p->v0, bytes0) != cudaSuccess) rc = MGB_ECUDA;
    if (rc == MGB_OK && cudaMalloc(&p->f0, bytes0) != cudaSuccess) rc = MGB_ECUDA;
    if (rc != MGB_OK)
    {
        if (rc == MGB_ECUDA)
        {
            set_error("mgb_precond_create: device allocation failed");
            (void)cudaGetLastError();
        }
        mgb_precond_destroy(p);
        return rc;
    }
    // work blocks start zeroed like freshly constructed GridFuncVectors
    for (size_t l = 0; l < p->work.size(); l++)
        cudaMemset(p->work[l], 0, sizeof(float) * sizeg_of(p->grid[l]) * nfunc_max);
    *out = p;
    return MGB_OK;
}

int mgb_precond_destroy(mgb_precond* p)
{
    if (!p) return MGB_OK;
    for (float* q : p->work)
        if (q) cudaFree(q);
    for (float* q : p->rcoarse)
        if (q) cudaFree(q);
    for (float* q : p->newv)
        if (q) cudaFree(q);
    if (p->v0) cudaFree(p->v0);
    if (p->f0) cudaFree(p->f0);
    delete p;
    return MGB_OK;
}

int mgb_precond_vcycle(mgb_precond* p, float* v, const float* f, int nfunc, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(p && v && f, "mgb_precond_vcycle: null pointer");
    MGB_REQUIRE(nfunc >= 0 && nfunc <= p->nfunc_max, "nfunc %d > nfunc_max %d",
        nfunc, p->nfunc_max);
    if (nfunc == 0) return MGB_OK;
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
    const size_t n     = sizeg_of(gr) * nfunc;
    cudaStream_t st    = as_stream(stream);
    int rc;
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
