// Internal interface between the mgb_hpsi dispatcher and its three paths.
#pragma once
#include "common.cuh"

namespace mgb
{

struct HpsiArgs
{
    int lap_type; // MGB_LAP_4M / MGB_LAP_4MP / MGB_LAP_4
    int dtype;
    const mgb_grid* grid;
    int g; // stencil radius: 1 Mehrstellen, 2 4th order (src/GridFactory.h:23-51)
    const void* phi;
    size_t ld;
    const double* vtot;
    void* hphi;
    size_t ldh;
    int nfunc;
    const void* xhalo_phi;
    const double* xhalo_v;
    // direct peer reads of the x neighbours' blocks (same layout as phi): the
    // g planes below come from the last planes of peer_w, the g planes above
    // from the first planes of peer_e; null = use xhalo_phi
    const void* peer_w;
    const void* peer_e;
    const int* map_w; // color slot of my color's orbital on that rank (or null)
    const int* map_e;
    // any px x py x pz decomposition, every halo read in place (mgb_hpsi_peer3d):
    // nb3d[((dx+1)*3 + (dy+1))*3 + (dz+1)] = the phi block of the rank at coord + (dx,dy,dz)
    // as mapped here (my own block where a direction is not split), or null = not this path;
    // vghost: ghosted copy of V (width g, boundaries traded)
    const void* const* nb3d;
    const double* vghost;
    const void* const* nbz; // z split: the 9 z-halo column buffers (comm_zhalo), else null
};

// path 1: TMA-pipelined x-streaming kernel (hpsi_fused.cu).  Returns
// MGB_ENOTSUP (without setting an error) when the box is not eligible.
int hpsi_tma(const HpsiArgs& a, cudaStream_t st);
// template arguments and tile configuration of the last TMA launch
const char* hpsi_last_kernel();
// path 2: generic fused kernel (hpsi_generic.cu), bit-exact.
int hpsi_generic(const HpsiArgs& a, cudaStream_t st);
// B u (Mehrstellen right-hand-side operator; mehr2: Laph4MP's B2) on a no-ghost block, bit-exact
// (hpsi_generic.cu); xhalo: [nfunc][2][ny][nz] on x-split boxes, else null
int rhs_generic(int dtype, bool mehr2, const mgb_grid* gr, const void* phi, size_t ld,
    const void* xhalo, void* out, size_t ldo, int nfunc, cudaStream_t st);
// Out = alpha A M + beta Out + gamma D (contractions.cu); D may be null
int gemm_nn_fused(int dtype, size_t m, int n, int k, double alpha, const void* A, size_t lda,
    const double* M, int ldm, double beta, void* Out, size_t ldc, double gamma, const void* D,
    size_t ldd, cudaStream_t st);
// path 3: reference-shaped composition on ghosted blocks (api.cu).
int hpsi_ghosted(const HpsiArgs& a, cudaStream_t st);

// literal Jacobi sweep on ghosted float blocks (fd_ghosted.cu)
int jacobi_literal(int lap_type, const mgb_grid* gr, float* v, const float* f,
    float* w, int nfunc, double omega, cudaStream_t st);

// Laph4::applyWithPot on a ghosted block (fd_ghosted.cu)
int del2_4th_withpot(int dtype, const mgb_grid* gr, const void* v,
    const double* pot, void* out, size_t ldo, int nfunc, cudaStream_t st);

// phases of GridFuncVector::trade_boundaries on a ghosted block (fd_ghosted.cu)
int trade_dirichlet(int dtype, const mgb_grid* gr, void* u, int nfunc, cudaStream_t st);
int trade_wrap(int dtype, const mgb_grid* gr, void* u, int nfunc, int d, cudaStream_t st);
// pack (ghosted -> buf) or unpack a sub-box of every function
int subbox_copy(int dtype, bool pack, const mgb_grid* gr, const int lo[3],
    const int ext[3], void* u, void* buf, int nfunc, cudaStream_t st);

// multi-GPU (comm.cu): the registered array `local` of rank `rank` mapped into
// this process (CUDA IPC over NVLink), a stream-ordered barrier over the ranks,
// and the rank at Cartesian coordinates
const void* peer_view(mgb_comm* c, const void* local, int rank);
int comm_barrier(mgb_comm* c, cudaStream_t st);
// barrier with the x neighbours of `gr` only, through peer memory (falls back
// to comm_barrier when the inboxes cannot be mapped)
int comm_barrier_neighbors(mgb_comm* c, const mgb_grid* gr, cudaStream_t st);
int comm_rank_of(const mgb_grid* gr, int cx, int cy, int cz);
// push the z-edge columns of a block to the z neighbours' dense buffers; nbz[(dx+1)*3 +
// (dy+1)]: the buffers ([function][x][y][32 bytes]) the consuming kernel reads (comm.cu)
int comm_zhalo(mgb_comm* c, const mgb_grid* gr, int es, const void* src, size_t ld, int nfunc,
    cudaStream_t st, const void* nbz[9]);
// LocGridOrbitals on split domains: the color slot of my color's orbital on the
// west / east rank (device arrays, -1 = not held there), null = same slot
void comm_color_maps(mgb_comm* c, const int** map_w, const int** map_e, int* n);

// grow-only device scratch owned by the library (workspaces, never user data)
void* scratch(int slot, size_t bytes);

} // namespace mgb
