/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference's
 * orbital-update hot path, in plain C.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this; the
 * product library (libmgmol_b200.so) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks every function
 * here bit-for-bit (integer compare of the float/double bit patterns) against
 * the reference's own compiled sources (oracle/_ref/libmgmol_ref.so, built by
 * oracle/Makefile from /root/reference/src, unmodified), and
 * tests/golden/*.npz holds outputs of that compiled reference for the GPU
 * box, where /root/reference does not exist.  Exception, stated where it
 * applies: the FP64 contractions go through an un-vendored Fortran BLAS in the
 * reference (DGEMM/DSYRK, no pinned version) -- unpinned at the last-bit level;
 * the reference's own double-accumulating loops are restated instead.
 *
 * Compile with -ffp-contract=off: the reference build (g++ -O2, generic
 * x86-64) has no FMA contraction, and bit-exactness depends on that.
 *
 * All file:line citations are into /root/reference/src.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_INV12 (1. / 12.) /* pb/FDkernels.cc:9  */
#define ORC_INV64 (1. / 64.) /* pb/MGkernels.cc:14 */

#define ORC_CAT_(a, b) a##b
#define ORC_CAT(a, b) ORC_CAT_(a, b)

/* f64 first: the f32 instantiation of orc_hpsi calls the f64 ghost helpers
 * for the (always double) potential. */
#define T double
#define FN(name) ORC_CAT(name, _f64)
#include "mgmol_oracle_impl.inc"
#undef T
#undef FN

#define T float
#define FN(name) ORC_CAT(name, _f32)
#include "mgmol_oracle_impl.inc"
#undef T
#undef FN

int orc_lap_constants(int lap_type, const double h[3], double out[3]);

/* ---- non-local Kleinman-Bylander projectors (pinned against the compiled
 *      KBprojectorSparse, tests/test_kb_cpu.py) ---- */
#define T double
#define FN(name) ORC_CAT(name, _f64)
#include "mgmol_oracle_kb.inc"
#undef T
#undef FN
#define T float
#define FN(name) ORC_CAT(name, _f32)
#include "mgmol_oracle_kb.inc"
#undef T
#undef FN

/* ---- Poisson solvers of the Hartree potential ----------- */
#define ORC_POISSON_PART 1
#define T double
#define FN(name) ORC_CAT(name, _f64)
#include "mgmol_oracle_poisson.inc"
#undef T
#undef FN
#define T float
#define FN(name) ORC_CAT(name, _f32)
#include "mgmol_oracle_poisson.inc"
#undef T
#undef FN
#undef ORC_POISSON_PART

/* PCGSolver::preconSolve (PCGSolver.cc:112-162) in POISSONPRECONDTYPE = float:
 * the solver's operator on level 0, Laph2 below; max(4, nu1 + nu2) sweeps on the
 * coarsest level; the work array of the last sweep is what gets restricted; the
 * trailing boundary trade under the condition exactly as written at :161. */
static void orc_pcg_precon(int lap_type, pfield_f32* v, const pfield_f32* f, int level,
    int nlevels, int ldims[][3], int nu1, int nu2, const double ll[3], const int bc[3])
{
    const int last     = level == nlevels;
    const int ncycl    = last ? (4 > nu1 + nu2 ? 4 : nu1 + nu2) : nu1;
    const int lt       = level == 0 ? lap_type : 1;
    const double scale = pf_scale_f32(lt, v);
    pfield_f32 work    = pf_new_f32(v->dims, v->g, ll, 0);
    for (int it = 0; it < ncycl; it++)
        pf_jacobi_f32(lt, v, f, &work, scale, bc);
    if (!last)
    {
        pfield_f32 rcoarse = pf_new_f32(ldims[level + 1], v->g, ll, 0);
        pfield_f32 newv    = pf_new_f32(ldims[level + 1], v->g, ll, 0);
        pf_trade_f32(&work, bc);
        orc_restrict3D_f32(work.dims, work.g, work.u, rcoarse.u, 1);
        rcoarse.upd = 0;
        pf_zero_f32(&newv);
        orc_pcg_precon(lap_type, &newv, &rcoarse, level + 1, nlevels, ldims, nu1, nu2, ll, bc);
        pf_trade_f32(&newv, bc);
        orc_extend3D_f32(work.dims, work.g, newv.u, work.u, 1);
        work.upd = 0;
        pf_axpy_f32(-1., &work, v);
        for (int it = 0; it < nu2; it++)
            pf_jacobi_f32(lt, v, f, &work, scale, bc);
        if (bc[0] != 1 || bc[2] != 1) pf_trade_f32(v, bc);
        pf_free_f32(&rcoarse);
        pf_free_f32(&newv);
    }
    pf_free_f32(&work);
}

#define ORC_POISSON_PART 2
#define T double
#define FN(name) ORC_CAT(name, _f64)
#include "mgmol_oracle_poisson.inc"
#undef T
#undef FN
#define T float
#define FN(name) ORC_CAT(name, _f32)
#include "mgmol_oracle_poisson.inc"
#undef T
#undef FN
#undef ORC_POISSON_PART

/* ---- operator constants -------------------------------------------------- */

/* diagEl / invDiagEl / jacobiFactor of the Lap family:
 * Laph4M pb/Laph4M.h:29-31,80; Laph2 pb/Laph2.h:32-34,95;
 * Laph4 pb/Laph4.h:35-38,106; Laph6 pb/Laph6.h:40-42,101;
 * Laph8 pb/Laph8.h:40-42,100; Laph4MP inherits Laph4M. */
int orc_lap_constants(int lap_type, const double h[3], double out[3])
{
    const double s = 1. / (h[0] * h[0]) + 1. / (h[1] * h[1]) + 1. / (h[2] * h[2]);
    double diag, jf;
    switch (lap_type)
    {
        case 0:
        case 10:
            diag = (4. / 3.) * s;
            jf   = 1. / diag;
            break;
        case 1:
            diag = 2. * s;
            jf   = 6. * (1. / diag) / 7.;
            break;
        case 2:
            diag = 2.5 * s;
            jf   = (1. / diag) / 1.5;
            break;
        case 3:
            diag = (49. / 18.) * s;
            jf   = (1. / diag) / 1.5;
            break;
        case 4:
            diag = (1435. / 504.) * s;
            jf   = (1. / diag) / 1.5;
            break;
        default:
            return 1;
    }
    out[0] = diag;
    out[1] = 1. / diag;
    out[2] = jf;
    return 0;
}

/* OrbitalsPreconditioning<T>::setGamma (OrbitalsPreconditioning.cc:120-145) */
double orc_gamma(double inv_diag, int mg_levels, double vmax, double small_eig)
{
    double gamma = inv_diag;
    for (int ln = 0; ln <= mg_levels; ln++)
        gamma *= 4.;
    return 1.0 / (2.0 / gamma + fabs(vmax - small_eig));
}

/* ---- float multigrid preconditioner -------------------------------------- */

typedef struct
{
    int dims[3];
    double h[3];
    size_t sizeg;
} orc_level;

/* GridFuncVector::applyLap (pb/GridFuncVector.cc:2370-2397) on float blocks */
static int orc_apply_lap_f32(int type, const orc_level* L, int g,
    const float* v, float* w, int nfunc)
{
    switch (type)
    {
        case 0:
            orc_del2_4th_Mehr_f32(L->dims, g, L->h, v, w, nfunc);
            return 0;
        case 1:
            orc_del2_2nd_f32(L->dims, g, L->h, v, w, nfunc);
            return 0;
        case 2:
            orc_del2_4th_f32(L->dims, g, L->h, v, w, nfunc);
            return 0;
        default:
            return 1; /* reference: abort() */
    }
}

/* GridFuncVector::jacobi (pb/GridFuncVector.cc:2416-2425):
 *   applyLap(type, w)  [trades v's boundaries unless flagged up to date]
 *   w -= f             [whole ghosted extent, MPaxpy(-1)]
 *   v += (-omega) w    [whole ghosted extent, MPaxpy]
 * *v_updated mirrors GridFuncVector::updated_boundaries_ of v. */
int orc_jacobi_f32(int type, const int dims[3], const double h[3], int g,
    const int bc[3], float* v, int* v_updated, const float* f, float* w,
    int nfunc, double jacobi_factor)
{
    orc_level L;
    memcpy(L.dims, dims, sizeof(L.dims));
    memcpy(L.h, h, sizeof(L.h));
    L.sizeg = (size_t)(dims[0] + 2 * g) * (dims[1] + 2 * g) * (dims[2] + 2 * g);
    if (!*v_updated) orc_trade_boundaries_f32(dims, g, bc, v, nfunc);
    if (orc_apply_lap_f32(type, &L, g, v, w, nfunc)) return 1;
    orc_axpy_f32(L.sizeg * nfunc, -1., f, w);
    orc_axpy_f32(L.sizeg * nfunc, -1. * jacobi_factor, w, v);
    *v_updated = 0;
    return 0;
}

typedef struct
{
    int max_levels, g, nfunc;
    int bc[3];
    orc_level lev[16];
    double jf[16];
    float* work[16];
    float* rcoarse[16];
    float* newv[16];
    /* localization masks (NULL state: none), tables indexed
     * [level][iloc][color] */
    int subdivx, ncolors, mask_op;
    const int* mstate;
    const long long* mvoff;
    const double* mvalues;
} orc_mg;

/* GridFuncVector::app_mask(level) (pb/GridFuncVector.cc:2428-2438) */
static void orc_mg_mask(const orc_mg* M, int level, float* u)
{
    if (!M->mstate) return;
    const size_t per = (size_t)M->subdivx * M->ncolors;
    orc_app_mask_f32(M->lev[level].dims, M->g, M->subdivx, M->ncolors,
        M->mask_op, M->mstate + level * per, M->mvoff + level * per,
        M->mvalues, u, M->nfunc);
}

/* Preconditioning<float>::mg (Preconditioning.cc:155-216).
 * Returns nonzero where the reference would abort. */
static int orc_mg_cycle(orc_mg* M, float* v, int* v_upd, const float* f,
    int lap_type, int level)
{
    const orc_level* L = &M->lev[level];
    const int g = M->g, nf = M->nfunc;
    const int ncycl = (level == M->max_levels) ? 4 : 2;
    float* w        = M->work[level];
    int w_upd       = 0;

    for (int it = 0; it < ncycl; it++) /* :173-177 */
    {
        if (orc_jacobi_f32(lap_type, L->dims, L->h, g, M->bc, v, v_upd, f, w,
                nf, M->jf[level]))
            return 1;
        orc_mg_mask(M, level, v);
    }
    if (level == M->max_levels) return 0; /* :179 */

    orc_mg_mask(M, level, w); /* :184 LOCALIZATION */

    /* :189 restrict3D trades w first (pb/GridFuncVector.cc:1624-1631);
     * w's flag is false after applyLap (GridFuncVector.h:297) */
    if (!w_upd) orc_trade_boundaries_f32(L->dims, g, M->bc, w, nf);
    float* rc = M->rcoarse[level];
    orc_restrict3D_f32(L->dims, g, w, rc, nf);
    orc_mg_mask(M, level + 1, rc); /* :192 */

    float* nv          = M->newv[level];
    const orc_level* C = &M->lev[level + 1];
    memset(nv, 0, sizeof(float) * C->sizeg * nf); /* :198 resetData */
    int nv_upd = 1;                               /* resetData sets the flag */
    if (orc_mg_cycle(M, nv, &nv_upd, rc, 1, level + 1)) return 1; /* :199 */

    /* :201 extend3D trades the coarse block first (:1633-1641) */
    if (!nv_upd) orc_trade_boundaries_f32(C->dims, g, M->bc, nv, nf);
    orc_extend3D_f32(L->dims, g, nv, w, nf);
    orc_mg_mask(M, level, w); /* :204 */

    orc_axpy_f32(L->sizeg * nf, -1., w, v); /* :206  v -= w */
    *v_upd = 0;                             /* extend3D left w un-traded */

    for (int it = 0; it < 2; it++) /* :209-213 */
    {
        if (orc_jacobi_f32(lap_type, L->dims, L->h, g, M->bc, v, v_upd, f, w,
                nf, M->jf[level]))
            return 1;
        orc_mg_mask(M, level, v);
    }

    /* :215 (the reference tests bc_[0], bc_[2], bc_[2]) */
    if (M->bc[0] != 1 || M->bc[2] != 1 || M->bc[2] != 1)
    {
        if (!*v_upd) orc_trade_boundaries_f32(L->dims, g, M->bc, v, nf);
        *v_upd = 1;
    }
    return 0;
}

/* Preconditioning<float>::mg on caller-provided ghosted float blocks
 * v (in/out) and f.  dims must be divisible by 2^mg_levels. */
int orc_mg_masked_f32(int lap_type, int mg_levels, const int dims[3],
    const double ll[3], const int bc[3], int g, float* v, const float* f,
    int nfunc, int subdivx, int ncolors, int mask_op, const int* mstate,
    const long long* mvoff, const double* mvalues)
{
    orc_mg M;
    memset(&M, 0, sizeof(M));
    M.subdivx = subdivx;
    M.ncolors = ncolors;
    M.mask_op = mask_op;
    M.mstate  = mstate;
    M.mvoff   = mvoff;
    M.mvalues = mvalues;
    if (mg_levels > 14) return 1;
    M.max_levels = mg_levels;
    M.g          = g;
    M.nfunc      = nfunc;
    memcpy(M.bc, bc, sizeof(M.bc));
    for (int l = 0; l <= mg_levels; l++)
    {
        for (int d = 0; d < 3; d++)
        {
            if (l > 0 && (M.lev[l - 1].dims[d] % 2)) return 2;
            M.lev[l].dims[d] = (l == 0) ? dims[d] : M.lev[l - 1].dims[d] / 2;
            M.lev[l].h[d]    = ll[d] / (double)M.lev[l].dims[d];
        }
        M.lev[l].sizeg = (size_t)(M.lev[l].dims[0] + 2 * g)
                         * (M.lev[l].dims[1] + 2 * g)
                         * (M.lev[l].dims[2] + 2 * g);
        double c[3];
        /* Preconditioning.cc:26-27 (level 0: lap_type), :122-123 (coarse: 1) */
        if (orc_lap_constants(l == 0 ? lap_type : 1, M.lev[l].h, c)) return 1;
        M.jf[l]   = c[2];
        M.work[l] = (float*)calloc(M.lev[l].sizeg * nfunc, sizeof(float));
        if (l > 0)
        {
            M.rcoarse[l - 1]
                = (float*)calloc(M.lev[l].sizeg * nfunc, sizeof(float));
            M.newv[l - 1]
                = (float*)calloc(M.lev[l].sizeg * nfunc, sizeof(float));
        }
    }
    int v_upd = 0;
    int rc    = orc_mg_cycle(&M, v, &v_upd, f, lap_type, 0);
    for (int l = 0; l <= mg_levels; l++)
    {
        free(M.work[l]);
        if (l > 0)
        {
            free(M.rcoarse[l - 1]);
            free(M.newv[l - 1]);
        }
    }
    return rc;
}

int orc_mg_f32(int lap_type, int mg_levels, const int dims[3],
    const double ll[3], const int bc[3], int g, float* v, const float* f,
    int nfunc)
{
    return orc_mg_masked_f32(lap_type, mg_levels, dims, ll, bc, g, v, f, nfunc,
        1, nfunc, 0, NULL, NULL, NULL);
}

/* OrbitalsPreconditioning<T>::precond_mg (OrbitalsPreconditioning.cc:87-117):
 * res (no-ghost, ORBDTYPE) <- M^-1 res in float.  dtype: 0 float, 1 double.
 * mstate == NULL: no masks (ExtendedGridOrbitals). */
int orc_precond_mg_masked(int lap_type, int mg_levels, int dtype,
    const int dims[3], const double ll[3], const int bc[3], void* res,
    int nfunc, double gamma, int subdivx, int ncolors, int mask_op,
    const int* mstate, const long long* mvoff, const double* mvalues)
{
    const int mehr = (lap_type == 0 || lap_type == 10);
    const int g    = mehr ? 1 : 2;
    const size_t npt   = (size_t)dims[0] * dims[1] * dims[2];
    const size_t sizeg
        = (size_t)(dims[0] + 2 * g) * (dims[1] + 2 * g) * (dims[2] + 2 * g);
    float* v  = (float*)calloc(sizeg * nfunc, sizeof(float)); /* :99  */
    float* f  = (float*)malloc(sizeof(float) * sizeg * nfunc);
    float* in = (float*)malloc(sizeof(float) * npt * nfunc);
    /* :103 setDataWithGhosts(gfv_work2_): ORBDTYPE -> float (MPcpy) */
    if (dtype == 1)
        for (size_t i = 0; i < npt * nfunc; i++)
            in[i] = (float)((const double*)res)[i];
    else
        memcpy(in, res, sizeof(float) * npt * nfunc);
    orc_add_ghosts_f32(dims, g, in, f, nfunc);
    orc_axpy_f32(sizeg * nfunc, gamma, f, v); /* :104 */
    int rc = orc_mg_masked_f32(lap_type, mg_levels, dims, ll, bc, g, v, f,
        nfunc, subdivx, ncolors, mask_op, mstate, mvoff, mvalues);
    /* :109 setPsi(*gfv_work_): float -> ORBDTYPE, strip ghosts */
    orc_strip_ghosts_f32(dims, g, v, in, nfunc);
    if (dtype == 1)
        for (size_t i = 0; i < npt * nfunc; i++)
            ((double*)res)[i] = (double)in[i];
    else
        memcpy(res, in, sizeof(float) * npt * nfunc);
    free(v);
    free(f);
    free(in);
    return rc;
}

int orc_precond_mg(int lap_type, int mg_levels, int dtype, const int dims[3],
    const double ll[3], const int bc[3], void* res, int nfunc, double gamma)
{
    return orc_precond_mg_masked(lap_type, mg_levels, dtype, dims, ll, bc, res,
        nfunc, gamma, 1, nfunc, 0, NULL, NULL, NULL);
}

int orc_version(void) { return 2; }
