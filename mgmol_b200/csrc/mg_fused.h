// Internal interface of the fused multigrid kernels (mg_fused.cu) used by the
// V-cycle driver in mg_precond.cu.  All blocks are no-ghost float arrays
// [function][x][y][z] with an explicit leading dimension between functions.
#pragma once
#include "common.cuh"
#include "masks.h"

namespace mgb
{

struct MgJacobiArgs
{
    int lap_type;         // MGB_LAP_4M, MGB_LAP_2 or MGB_LAP_4
    const mgb_grid* grid; // dims, h, bc of this level
    const float* in;      // v (or f when scale != 1: v = scale * f on the fly)
    size_t ld_in;
    const float* f;       // right-hand side
    size_t ld_f;
    float* out;           // v' as float (may be null when out64 is set)
    double* out64;        // v' as double (exit of precond_mg for ORBDTYPE double)
    size_t ld_out;
    float* w;             // A v - f (null: not needed)
    size_t ld_w;
    int nfunc;
    double omega;         // jacobiFactor of this level
    double scale;         // 1, or the factor applied to `in` while it is loaded
    int zero_low[3];      // store zeros on the low layer (Dirichlet trade)
    MaskView mask;        // localization mask of this level applied to the
                          // stored v' and w (off == nullptr: none)
    // x-split domains (grid->nproc[0] > 1): the x neighbours' copies of `in`
    // (peer memory, same layout); null where there is no neighbour
    const float* peer_w;
    const float* peer_e;
    // LocGridOrbitals: color slot of my color's orbital on that rank (device
    // arrays, -1 = absent), or null = same slot
    const int* map_w;
    const int* map_e;
    // any px x py x pz decomposition: nb3d[((dx+1)*3 + (dy+1))*3 + (dz+1)] = the copy of `in`
    // of the rank at coord + (dx,dy,dz) (my own across a direction that is not split), every
    // halo read in place; null = not this path
    const float* const* nb3d;
    const void* const* nbz; // z split: the 9 buffers of pushed z-edge columns (comm_zhalo)
};

// true when the fused kernels can run this level (z extent a multiple of 4,
// a tile configuration exists)
bool mg_fused_level_ok(const mgb_grid& g, int lap_type);

int mg_jacobi(const MgJacobiArgs& a, cudaStream_t st);
// `mask`: of the COARSE level (applied to the restricted block)
// nb3d (both transfers): the copies of the source block on the 27 Cartesian neighbours (any
// decomposition, halos read in place; see MgJacobiArgs::nb3d), or null
// w_west: the west neighbour's copy of `w` on an x-split domain (its last
// fine plane is my plane -1), or null
int mg_restrict(const mgb_grid& fine, const float* w, size_t ldf, float* coarse, size_t ldc,
    int nfunc, const MaskView& mask, const float* w_west, const int* map_w, cudaStream_t st,
    const float* const* nb3d = nullptr);
// `mask`: of the fine level (applied to P e before the subtraction)
int mg_prolong_correct(const mgb_grid& fine, const float* coarse, size_t ldc, float* v,
    size_t ldv, int nfunc, const int zero_low[3], const MaskView& mask, const float* coarse_east,
    const int* map_e, cudaStream_t st, const float* const* nb3d = nullptr);
int mg_convert(size_t npt, const double* in, size_t ldi, float* out, size_t ldo, int nfunc,
    cudaStream_t st);
int mg_scale(const mgb_grid& gr, double c, const float* f, size_t ldf, float* v, size_t ldv,
    int nfunc, const int zero_low[3], const MaskView& mask, cudaStream_t st);

} // namespace mgb
