/*
 * mgmol_b200 -- C ABI of the B200-native orbital-update hot path of MGmol.
 *
 * This is the drop-in boundary: a C++ host (MGmol's own classes, specialised
 * for MemorySpace::Device) calls these entry points with plain device
 * pointers.  Every entry point names the reference interface it backs
 * (file:line into the MGmol source tree).  See INTEGRATION.md for the
 * reference-side stubs.
 *
 * Conventions
 *  - All array pointers are DEVICE pointers unless the name says host.
 *  - Orbital blocks without ghosts: column-major npt x nfunc, leading dimension
 *    `ld` (>= npt) elements, point (ix,iy,iz) at ix*ny*nz + iy*nz + iz
 *    (BlockVector storage, src/BlockVector.cc:79, src/pb/GridFunc.cc:610-626).
 *  - Ghosted blocks: nfunc x sizeg, function k at k*sizeg, point (ix,iy,iz) at
 *    (ix+g)*inc0 + (iy+g)*inc1 + iz+g with inc1 = nz+2g, inc0 = (ny+2g)*inc1
 *    (pb::GridFuncVector, src/pb/GridFuncVector.h:222-225, src/pb/Grid.cc:80-82).
 *  - dtype: MGB_F32 / MGB_F64 = ORBDTYPE float / double (src/global.h:18-22).
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default
 *    stream).  Calls are asynchronous with respect to the host; like the
 *    reference's single MAGMA queue (src/magma_singleton.h:32-33) all work
 *    of one call is ordered on that one stream.
 *  - Concurrency contract: ONE device and ONE stream of work per process, as
 *    in the reference (one MPI rank = one device = one queue).  The library
 *    keeps grow-only scratch blocks (work arrays of the ghosted composition,
 *    stream-K partial tiles, host-pipeline staging) and a few settings
 *    (mgb_hpsi_force_path, mgb_set_f32_contraction) per PROCESS, not per
 *    stream or handle: two calls that use scratch must not overlap on
 *    different streams or host threads, and the current device must not
 *    change between calls.  Entry points are not
 *    re-entrant; handles (mgb_precond, mgb_masks, mgb_kb, mgb_comm) own their
 *    work blocks and may coexist.
 *  - Return value: MGB_OK (0) or a negative MGB_E* code; mgb_last_error()
 *    gives the message.  The reference aborts on these conditions
 *    (src/pb/Lap.h:37-38, src/pb/FDkernels.h:42-45); the C++ shim is expected
 *    to do the same on a non-zero return.
 *  - There is NO CPU fallback: without a CUDA device every compute entry
 *    point returns MGB_ENODEVICE.
 */
#ifndef MGMOL_B200_H
#define MGMOL_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGB_OK 0
#define MGB_EINVAL (-1)      /* bad argument (reference: assert / abort)      */
#define MGB_ENOTSUP (-2)     /* valid in the reference, not built here        */
#define MGB_ECUDA (-3)       /* CUDA runtime / driver error                   */
#define MGB_ENODEVICE (-4)   /* no CUDA device                                */
#define MGB_ENCCL (-5)       /* NCCL error                                    */

#define MGB_F32 0
#define MGB_F64 1

/* lap_type values of LapFactory<T>::createLap (src/LapFactory.h:26-56) */
#define MGB_LAP_4M 0   /* Laph4M  : 4th-order Mehrstellen (default FDtype) */
#define MGB_LAP_2 1    /* Laph2   : 2nd order (MG coarse levels)           */
#define MGB_LAP_4 2    /* Laph4   : classical 4th order                    */
#define MGB_LAP_6 3    /* Laph6                                             */
#define MGB_LAP_8 4    /* Laph8                                             */
#define MGB_LAP_4MP 10 /* Laph4MP : SPD Mehrstellen (same A, other B)      */

/* FD kernel selector for mgb_fd_apply */
#define MGB_FD_DEL2_4TH_MEHR 0
#define MGB_FD_DEL2_2ND 1
#define MGB_FD_DEL2_4TH 2
#define MGB_FD_DEL2_6TH 3
#define MGB_FD_DEL2_8TH 4
#define MGB_FD_RHS_4TH_MEHR1 100

/*
 * Local box of one rank: the data pb::Grid + pb::PEenv carry
 * (src/pb/Grid.h:24-121, src/pb/PEenv.h:34-203).
 */
typedef struct mgb_grid
{
    int dim[3];    /* local dims nx,ny,nz            (Grid::dim)            */
    int gdim[3];   /* global dims                    (Grid::gdim)           */
    int ghosts;    /* ghost width of ghosted blocks  (Grid::ghost_pt)       */
    double h[3];   /* mesh spacing                   (Grid::hgrid)          */
    int bc[3];     /* 1 periodic, 0 Dirichlet-0      (ct.bcWF)              */
    int nproc[3];  /* ranks per direction            (PEenv::n_mpi_task)    */
    int coord[3];  /* this rank's coordinates        (PEenv::my_mpi)        */
} mgb_grid;

/* communicator of a decomposed run (section "multi-GPU" below) */
typedef struct mgb_comm mgb_comm;

const char* mgb_last_error(void);
int mgb_version(void);
/* number of kernels launched by this library since load (all streams) */
unsigned long long mgb_launch_count(void);
int mgb_device_count(void);

/* ---- memory: MemorySpace::Memory<T,Device> (src/tools/memory_space.h:253-333)
 *      allocate / free / copy_to_dev / copy_to_host / set                    */
int mgb_malloc(void** dptr, size_t bytes);
int mgb_free(void* dptr);
int mgb_copy_to_dev(void* dst_dev, const void* src_host, size_t bytes, void* stream);
int mgb_copy_to_host(void* dst_host, const void* src_dev, size_t bytes, void* stream);
int mgb_copy_dev(void* dst_dev, const void* src_dev, size_t bytes, void* stream);
int mgb_memset(void* dptr, int value, size_t bytes, void* stream);
int mgb_stream_sync(void* stream);

/* ---- batched FD kernels on ghosted blocks
 * FDkernelDel2_{2nd,4th,4th_Mehr,6th,8th}(grid, v, b, nfunc, Device) and
 * FDkernelRHS_4th_Mehr1(grid, v, rhs, rhs_ghosts, nfunc, Device)
 * (src/pb/FDkernels.h:14-75; Host bodies src/pb/FDkernels.cc:18-584).
 * v: nfunc x sizeg ghosted, ghosts already exchanged.  out: interior written,
 * ghosted with grid->ghosts (or `rhs_ghosts` for the RHS kernel; 0 = no-ghost
 * npt x nfunc output).                                                       */
int mgb_fd_apply(int kind, int dtype, const mgb_grid* grid, const void* v,
    void* out, int nfunc, int rhs_ghosts, void* stream);

/* ---- fused local Hamiltonian
 * Hamiltonian<T>::applyLocal(ncolors, phi, hphi) body (src/Hamiltonian.cc:
 * 85-159) and Lap<T>::applyWithPot (src/pb/Lap.h:35, src/pb/FDoper.cc:321-399):
 *   lap_type 0/10: hphi = A_Mehr phi + B (vtot .* phi)     (:106-132)
 *   lap_type 2   : hphi = A_4th  phi + vtot .* phi         (:133-156)
 * phi, hphi: no-ghost npt x nfunc blocks (ld elements apart); vtot: POTDTYPE
 * double[npt] without ghosts (src/Potentials.h:141).  One pass over phi: the
 * ghost-add, halo, V*psi, B, Laplacian, axpy and ghost-strip sweeps of the
 * reference are fused.  On an x-split box this entry takes the neighbour planes
 * from the packed halo buffers below; every decomposition, x slabs included, is
 * served without any packed copy by mgb_hpsi_peer / mgb_hpsi_peer3d (section
 * "multi-GPU"), which read the neighbours' blocks in place.
 * xhalo_phi / xhalo_v: NULL when grid->nproc[0]==1; otherwise device buffers
 * [nfunc][2g][ny][nz] (phi) and [2g][ny][nz] (vtot): the g planes below the
 * box then the g planes above it, as filled by mgb_halo_exchange_x.         */
int mgb_hpsi(int lap_type, int dtype, const mgb_grid* grid, const void* phi,
    size_t ld, const double* vtot, void* hphi, size_t ldh, int nfunc,
    const void* xhalo_phi, const double* xhalo_v, void* stream);

/* B phi on a no-ghost block: pb::Lap<T>::rhs (src/pb/Lap.h:32) = FDoper::
 * rhs_4th_Mehr1 / FDkernelRHS_4th_Mehr1 (src/pb/FDkernels.cc:522-584) for
 * Laph4M after setDataWithGhosts + trade_boundaries, the identity for the
 * non-compact operators; every orbital at once, boundary condition folded into
 * the index (bit-identical).  xhalo_phi: [nfunc][2][ny][nz] on x-split boxes.  */
int mgb_apply_b(int lap_type, int dtype, const mgb_grid* grid, const void* phi,
    size_t ld, void* bphi, size_t ldb, int nfunc, const void* xhalo_phi,
    void* stream);
/* MGmol::computeResidualUsingHPhi (src/MGmol.cc:1227-1287):
 *     res = (B phi) theta - hphi          [Ry]
 * theta = localT (column-major nfunc x nfunc double on the device), B applied
 * when the operator is Mehrstellen (ct.Mehrstellen()).  The reference's four
 * sweeps (ghost-add, B per orbital, multiplyByMatrix, axpy) become B phi into a
 * library workspace and ONE contraction pass whose epilogue subtracts hphi.   */
int mgb_residual(int lap_type, int dtype, const mgb_grid* grid, const void* phi,
    size_t ld, const void* hphi, size_t ldh, const double* theta, int ldt,
    void* res, size_t ldr, int nfunc, const void* xhalo_phi, void* stream);

/* The same operator for a caller whose orbitals live in HOST memory (MGmol's
 * default MemorySpace::Host build: BlockVector storage, src/BlockVector.cc:
 * 138-218): phi_host, vtot_host and hphi_host are host pointers.  The call
 * pipelines host->device copy, fused kernel and device->host copy over blocks
 * of `chunk` orbitals (0 = automatic, ~32 MB) on three internal streams so
 * that both PCIe directions and the GPU are busy at once, and returns when
 * hphi_host holds the result (synchronous, like the reference call).  Pin the
 * buffers once with mgb_host_register (cudaHostRegister) for full PCIe rate;
 * pageable buffers work but are staged by the driver.  Single-rank boxes.   */
int mgb_hpsi_host(int lap_type, int dtype, const mgb_grid* grid, const void* phi_host,
    size_t ld, const double* vtot_host, void* hphi_host, size_t ldh, int nfunc,
    int chunk);
/* The same for one rank of an x-split domain (collective over the
 * communicator): every rank streams ITS sub-box through ITS GPU and PCIe link;
 * the fused kernel of block i reads the neighbours' block i in place from their
 * input rings over NVLink, fenced by the neighbour barrier.                   */
int mgb_hpsi_host_peer(mgb_comm* comm, int lap_type, int dtype, const mgb_grid* grid,
    const void* phi_host, size_t ld, const double* vtot_host, void* hphi_host,
    size_t ldh, int nfunc, int chunk);
int mgb_host_register(void* ptr, size_t bytes);
int mgb_host_unregister(void* ptr);

/* Which implementation mgb_hpsi picked last (for tests and the bench):
 * 1 TMA-pipelined fused kernel, 2 generic fused kernel, 3 ghosted-block
 * composition.                                                              */
int mgb_hpsi_last_path(void);
/* Template arguments and tile configuration of the kernel mgb_hpsi launched last (the
 * signature the ncu captures under profiles/ are keyed by).                    */
const char* mgb_hpsi_last_kernel(void);
/* Force a path (0 = automatic).  Test hook. */
int mgb_hpsi_force_path(int path);

/* ---- GridFuncVector<T,Device> batch operations (src/pb/GridFuncVector.h:
 *      445-479) on ghosted blocks                                           */
/* BlockVector::setDataWithGhosts + GridFunc::assign (src/BlockVector.cc:
 * 489-517, src/pb/GridFunc.cc:577-630): zero the block, copy interiors in,
 * converting in_dtype -> out_dtype.                                         */
int mgb_gfv_set_with_ghosts(int in_dtype, int out_dtype, const mgb_grid* grid,
    const void* noghost, size_t ld, void* ghosted, int nfunc, void* stream);
/* BlockVector::assign(GridFuncVector) + GridFunc::getValues
 * (src/BlockVector.cc:303-311, src/pb/GridFunc.cc:1457-1496)               */
int mgb_gfv_get_values(int in_dtype, int out_dtype, const mgb_grid* grid,
    const void* ghosted, void* noghost, size_t ld, int nfunc, void* stream);
/* GridFuncVector::trade_boundaries (src/pb/GridFuncVector.cc:1544-1622) for
 * directions owned by a single rank: Dirichlet zeroing (src/pb/GridFunc.cc:
 * 2192-2336) then the local periodic wraps in Y, Z, X order (:586-603).
 * Directions with nproc > 1 are exchanged by mgb_halo_exchange_ghosted.     */
int mgb_gfv_trade_boundaries(int dtype, const mgb_grid* grid, void* ghosted,
    int nfunc, void* stream);
/* GridFuncVector::pointwiseProduct (src/pb/GridFuncVector.cc:90-136):
 * out_j = A_j .* V over the whole ghosted extent, V a ghosted double field. */
int mgb_gfv_pointwise_product(int dtype, const mgb_grid* grid, const void* A,
    const double* Vghost, void* out, int nfunc, void* stream);
/* GridFuncVector::axpy / operator-= (src/pb/GridFuncVector.cc:1645-1668) =
 * MPaxpy over nfunc*sizeg elements (n given explicitly).                    */
int mgb_axpy(int dtype, size_t n, double alpha, const void* x, void* y, void* stream);
/* LinearAlgebraUtils::MPscal / MPdot (src/linear_algebra/mputils.cc:53-215) */
int mgb_scal(int dtype, size_t n, double alpha, void* x, void* stream);
int mgb_dot(int dtype, size_t n, const void* x, const void* y, double* result_dev,
    void* stream);
/* ExtendedGridOrbitals::computeDiagonalElementsDotProduct (src/
 * ExtendedGridOrbitals.cc:1085-1106): result[i] = alpha * <x_i, y_i> for every
 * orbital of two no-ghost blocks (alpha = grid.vel()), one launch; the input of
 * Orbitals::dotProduct (dotProductDiagonal, :1205-1213) used by the residual
 * norm (src/MGmol.cc:1316) and by AndersonMix (src/AndersonMix.cc:88-110).   */
int mgb_dot_cols(int dtype, size_t n, int nfunc, double alpha, const void* x,
    size_t ldx, const void* y, size_t ldy, double* result_dev, void* stream);
/* GridFuncVector::jacobi (src/pb/GridFuncVector.cc:2416-2425):
 * w = A v ; w -= f ; v += -omega w, v's ghosts must be up to date.          */
int mgb_gfv_jacobi(int lap_type, const mgb_grid* grid, float* v, const float* f,
    float* w, int nfunc, double omega, void* stream);
/* GridFuncVector::restrict3D / extend3D (src/pb/GridFuncVector.cc:1624-1641,
 * kernels src/pb/MGkernels.cc:28-283).  `fine` describes the fine grid; the
 * coarse grid is dims/2 with the same ghost width (src/pb/Grid.cc:214-231).
 * Ghosts of the source block must be up to date.                            */
int mgb_gfv_restrict3D(int dtype, const mgb_grid* fine, const void* ufine,
    void* ucoarse, int nfunc, void* stream);
int mgb_gfv_extend3D(int dtype, const mgb_grid* fine, const void* ucoarse,
    void* ufine, int nfunc, void* stream);

/* ---- localization masks of LocGridOrbitals
 * The data Map2Masks / GridMask hold (src/Map2Masks.cc:25-61, src/GridMask.h:
 * 41-52), flattened per color: the caller resolves overlapping_gids_[iloc]
 * [color] -> GridMask when it fills the set.  For every multigrid level
 * (0..mg_levels), x-slab iloc (subdivx slabs of dim[0]/subdivx planes, src/
 * Mesh.cc:57-77) and color:
 *   state -1 / 0  zero the slab  (mask_not_zero_ <= 0, or gid == -1:
 *                                 src/Map2Masks.cc:53-58)
 *   state 1       keep
 *   state 2       apply `values` (host pointer, lmasktype `dtype`, the no-ghost
 *                 slab sub0_l x ny_l x nz_l of GridMask::lmask_[level][iloc])
 * op: how values are applied -- MGB_MASK_MULT  u *= mask  (GridMaskMult, src/
 * GridMaskMult.cc:30-88; the "corrected" masks) or MGB_MASK_MAX  clip |u| <=
 * mask (GridMaskMax, src/GridMaskMax.cc:30-90, GridMask::limitAbsValue; the
 * default orbital masks, src/MasksSet.cc:15,158-182).
 * mgb_masks_commit uploads the tables; call it after the last mgb_masks_set
 * (MasksSet::setup / update, src/MasksSet.cc:136-157,206-217).               */
#define MGB_MASK_MULT 0
#define MGB_MASK_MAX 1
typedef struct mgb_masks mgb_masks;
int mgb_masks_create(mgb_masks** out, const mgb_grid* grid, int mg_levels,
    int subdivx, int ncolors, int op);
int mgb_masks_set(mgb_masks* m, int level, int iloc, int color, int state,
    int dtype, const void* values_host);
int mgb_masks_commit(mgb_masks* m);
int mgb_masks_destroy(mgb_masks* m);
/* GridFuncVector::app_mask(level) (src/pb/GridFuncVector.cc:2428-2438) on a
 * ghosted block of the level-`level` grid (dims >> level, `ghosts` ghosts).
 * m == NULL: no-op, like map2masks_ == nullptr.                             */
int mgb_gfv_app_mask(int dtype, const mgb_masks* m, int level, int ghosts,
    void* ghosted, int nfunc, void* stream);
/* LocGridOrbitals::applyMask / app_mask(color, u, level) (src/
 * LocGridOrbitals.cc:427-452,487-509) on a no-ghost block.                  */
int mgb_app_mask(int dtype, const mgb_masks* m, int level, void* noghost,
    size_t ld, int nfunc, void* stream);

/* ---- multigrid preconditioner
 * Preconditioning<float> (src/Preconditioning.h:22-64, .cc:15-216) and
 * OrbitalsPreconditioning<T>::precond_mg / setGamma
 * (src/OrbitalsPreconditioning.cc:87-145).  The handle owns the float work
 * blocks of every level (v, f, work, rcoarse, newv).                        */
typedef struct mgb_precond mgb_precond;
int mgb_precond_create(mgb_precond** out, int lap_type, int mg_levels,
    const mgb_grid* grid, int nfunc_max);
int mgb_precond_destroy(mgb_precond* p);
/* Boxes that are one rank of a decomposition (grid->nproc): the communicator the
 * V-cycle exchanges ghosts over -- every GridFuncVector::trade_boundaries of
 * Preconditioning<float>::mg (MPI_Isend/Irecv in the reference, src/pb/
 * GridFuncVector.cc:1544-1622).  Any px x py x pz decomposition runs the fused
 * kernels, reading the neighbours' boundary planes, rows and (pushed) columns of
 * the work blocks in place over NVLink; boxes the fused kernels do not take
 * (mixed boundary conditions, blocks the neighbours cannot map) run the literal
 * sequence with the packed Y -> Z -> X exchange.  Calls are then collective
 * over the communicator.                                                     */
int mgb_precond_set_comm(mgb_precond* p, mgb_comm* comm);
/* OrbitalsPreconditioning::setup with currentMasks != nullptr (src/
 * OrbitalsPreconditioning.cc:59-67 -> GridFuncVector::setMasks): every
 * app_mask of Preconditioning<float>::mg (src/Preconditioning.cc:176,184,192,
 * 204,212) uses this set.  NULL removes it.  The set must outlive the calls. */
int mgb_precond_set_masks(mgb_precond* p, const mgb_masks* m);
/* res (no-ghost, dtype) <- M^-1 res : v0 = gamma*res, mg(v, f=res) in float */
int mgb_precond_mg(mgb_precond* p, int dtype, void* res, size_t ld, int nfunc,
    double gamma, void* stream);
/* Implementation selector (test / profiling hook): 0 automatic, 1 the literal
 * reference-shaped sequence on ghosted blocks (bit-identical to the reference
 * CPU build), 2 the fused kernels on no-ghost blocks (one pass per Jacobi
 * sweep; float difference-form stencils, agrees with the reference to ~1e-7
 * relative).  mgb_precond_last_mode: what the last mgb_precond_mg used.      */
int mgb_precond_set_mode(mgb_precond* p, int mode);
int mgb_precond_last_mode(mgb_precond* p);
/* Preconditioning<float>::mg on caller-owned ghosted float blocks           */
int mgb_precond_vcycle(mgb_precond* p, float* v, const float* f, int nfunc,
    void* stream);
/* diagEl, invDiagEl, jacobiFactor of pb::Lap (src/pb/Laph4M.h:80,
 * Laph4.h:106, Laph2.h:95); out[3] on host.                                 */
int mgb_lap_constants(int lap_type, const double h[3], double out[3]);
/* OrbitalsPreconditioning::setGamma arithmetic (host)                       */
double mgb_gamma(double inv_diag, int mg_levels, double vmax, double small_eig);

/* ---- non-local Kleinman-Bylander projectors: the sparse projector
 *      vectors of KBprojectorSparse (src/KBprojectorSparse.h:39-52) and the two grid-sized
 *      steps that follow Hamiltonian::applyLocal in MGmol::getHpsiAndTheta
 *      (src/computeHij.cc:404-455).                                               */
typedef struct mgb_kb mgb_kb;
/* dtype: KBPROJDTYPE = ORBDTYPE (src/global.h:38); npt: points of the local box        */
int mgb_kb_create(mgb_kb** out, int dtype, size_t npt);
/* One ion overlapping the box (Ions::overlappingNL_ions order): its node list nlindex
 * (positions in the no-ghost orbital storage; KBprojectorSparse::nlindex_), nproj value
 * arrays of size_nl values each (host, `dtype`), coeff[p] = kbcoeff_p * sign_p
 * (src/get_vnlpsi.cc:48-74).  *first_row: the row of its first projector in kbpsi.      */
int mgb_kb_add_ion(mgb_kb* kb, int size_nl, const int* nlindex, int nproj, const void* proj,
    const double* coeff, int* first_row);
int mgb_kb_commit(mgb_kb* kb); /* upload; builds the point-major view of the scatter       */
int mgb_kb_nrows(const mgb_kb* kb);
int mgb_kb_destroy(mgb_kb* kb);
/* KBPsiMatrixSparse::computeKBpsi (src/KBPsiMatrixSparse.cc:136-212; computeLocalElement,
 * src/KBPsiMatrixInterface.cc:20-60): kbpsi[row * nfunc + f] = vel * <beta_row | psi_f>,
 * double, device.  psi: the orbitals, or B phi for the Mehrstellen flag.  The sum over the
 * ranks (globalSumKBpsi) is mgb_allreduce_sum_f64 on kbpsi.                             */
int mgb_kb_psi(const mgb_kb* kb, int dtype, double vel, const void* psi, size_t ld, int nfunc,
    double* kbpsi, void* stream);
/* get_vnlpsi (src/get_vnlpsi.cc:24-87) for every function: vnlpsi_f = sum over the ions, in
 * order, of (T)(sum_p kbpsi[row_p, f] coeff_p beta_p) with the roundings of axpySKet /
 * axpyKet (src/KBprojectorSparse.cc:865-908).  add != 0: out_f += vnlpsi_f, the MPaxpy of
 * computeHnlPhiAndAdd2HPhi (src/computeHij.cc:346-372: H phi += V_nl phi); add == 0:
 * out_f = vnlpsi_f (Mehrstellen: the caller applies B, mgb_apply_b, then adds).          */
int mgb_kb_vnlpsi(const mgb_kb* kb, int dtype, const double* kbpsi, void* out, size_t ldo,
    int nfunc, int add, void* stream);

/* ---- dense contractions: LinearAlgebraUtils<Device>::MPgemm / MPsyrk /
 *      MPgemmNN (src/linear_algebra/mputils.cc:295-1067), as the Orbitals
 *      classes call them.                                                   */
/* C(m x n, double, ldc) = alpha * A^T B + beta C, A: k x m (lda), B: k x n
 * (ldb), both `dtype`, accumulated in double.  ExtendedGridOrbitals::
 * computeLocalProduct (src/ExtendedGridOrbitals.cc:1049-1083), getLocalOverlap
 * / computeGram (:985-1010) with A==B, addDotWithNcol2Matrix (:1704-1752).  */
int mgb_gemm_tn(int dtype, int m, int n, size_t k, double alpha, const void* A,
    size_t lda, const void* B, size_t ldb, double beta, double* C, int ldc,
    void* stream);
/* ORBDTYPE float operands: 0 (default) = FP32-class tensor tiles, error-
 * compensated 3xTF32 with chunked exact accumulation (agrees with the
 * reference's double-accumulating loops to ~1e-6 of |a||b|, bar 1e-5) on the
 * 5th-generation tensor cores (tcgen05.mma kind::tf32, TMEM accumulators) for
 * mgb_gemm_tn / mgb_syrk_t, their slab variants and mgb_gemm_nn; 1 = the FP64
 * DMMA kernels on widened operands (the reference's products and sums); 2 =
 * the 3xTF32 arithmetic on mma.sync (the previous generation's kernels).
 * Process-wide.  */
int mgb_set_f32_contraction(int mode);
/* Test hook (host only, no device): the stream-K decomposition of a
 * contraction -- segments (cta, tile, first k-iteration, one past the last) for
 * `ncta` CTAs, K slabs of `kc` points, diagonal-tile cost `cd`/16.            */
int mgb_debug_tn_plan(int syrk, int m, int n, size_t k, int nbatch, int kc, int ncta,
    int cd, long long* segs, int max_segs, int* nsegs, long long* nkt_out,
    int* ntiles_out, int* ndiag_out);
/* Gram: C = alpha * A^T A (full symmetric matrix written, as
 * syrk('l','t') + fillUpperWithLower do, src/local_matrices/LocalMatrices.cc:
 * 210-247)                                                                  */
int mgb_syrk_t(int dtype, int n, size_t k, double alpha, const void* A,
    size_t lda, double* C, int ldc, void* stream);
/* The same per x-slab, as LocGridOrbitals does (src/LocGridOrbitals.cc:
 * 1504-1530 getLocalOverlap: LocalMatrices::syrk(iloc, loc_numpt, psi + iloc *
 * loc_numpt, lda); :1554-1604 computeLocalProduct: LocalMatrices::gemm(iloc,
 * ...)): slab s uses rows [s*k, (s+1)*k) of every column of A (and B) and
 * writes the s-th matrix of C (nslabs matrices, ldc*n doubles apart:
 * LocalMatrices storage, src/local_matrices/LocalMatrices.h:37-62).  One
 * launch for all slabs.                                                      */
int mgb_gemm_tn_slabs(int dtype, int m, int n, size_t k, int nslabs, double alpha,
    const void* A, size_t lda, const void* B, size_t ldb, double beta, double* C,
    int ldc, void* stream);
int mgb_syrk_t_slabs(int dtype, int n, size_t k, int nslabs, double alpha,
    const void* A, size_t lda, double* C, int ldc, void* stream);
/* Cout(m x n, dtype, ldc) = alpha * A(m x k, dtype, lda) * M(k x n, double,
 * ldm) + beta*Cout : ExtendedGridOrbitals::multiplyByMatrix
 * (src/ExtendedGridOrbitals.cc:448-498), MPgemmNN.  Cout must not alias A.  */
int mgb_gemm_nn(int dtype, size_t m, int n, int k, double alpha, const void* A,
    size_t lda, const double* M, int ldm, double beta, void* Cout, size_t ldc,
    void* stream);

/* Rho::computeRhoSubdomainUsingBlas3 (src/Rho.cc:359-448) for one x-slab of
 * nrows points: rho[i] += sum_j (Phi1 X)[i][j] * phi2[i][j], X = localX (column-
 * major nfunc x nfunc double on the device), rho RHODTYPE double.  The
 * contraction runs in point chunks through a library workspace (MPgemmNN
 * rounding to ORBDTYPE kept), the accumulation visits j in the reference's
 * order.                                                                     */
int mgb_rho_blas3(int dtype, size_t nrows, int nfunc, const void* phi1, size_t ld1,
    const double* X, int ldx, const void* phi2, size_t ld2, double* rho,
    void* stream);

/* ---- Hartree Poisson solvers on one scalar field ------
 * solver MGB_POISSON_MG  = pb::SolverLap<Lap<T>,T>::solve (src/pb/SolverLap.cc:
 *                          62-72): pb::Mgm (src/pb/Mgm.h:21-112) over pb::Vcycle
 *                          (src/pb/Vcycle.h:29-250), average0 when fully periodic;
 *        MGB_POISSON_PCG = PCGSolver<Lap<T>,T>::solve (src/PCGSolver.cc:165-252),
 *                          float multigrid preconditioner (preconSolve, :112-162).
 * lap_type MGB_LAP_4M / MGB_LAP_2 / MGB_LAP_4; grid->bc 0 or 1; single rank
 * (the solvers choose their own ghost width, grid->ghosts is not used).  vh (in:
 * initial guess, out: solution) and rho: no-ghost device arrays of `dtype`.
 * The parameters are those of Solver::setup (nu1, nu2, max_sweeps, tol,
 * max_nlevels).  stats (host, 5 doubles, may be null): converged, nb_sweeps,
 * final_residual, final_relative_residual, residual_reduction (-1 where the
 * solver has no such getter).  The control flow runs on the host over this
 * library's grid kernels (include/mgmol_b200_poisson.hpp holds it as C++
 * templates); the call returns after the solution is in vh.                  */
#define MGB_POISSON_MG 0
#define MGB_POISSON_PCG 1
int mgb_poisson_solve(int solver, int lap_type, int dtype, const mgb_grid* grid, void* vh,
    const void* rho, int nu1, int nu2, int max_sweeps, double tol, int max_nlevels,
    double* stats);

/* ---- multi-GPU: one process per GPU, 3-D block decomposition of pb::PEenv.
 * The communicator wraps NCCL; the unique id (128 bytes) is created on rank 0
 * with mgb_comm_unique_id and distributed by the caller (MPI_Bcast in MGmol,
 * torch.distributed in the tests).                                          */
int mgb_comm_unique_id(void* id128);
int mgb_comm_create(mgb_comm** out, const void* id128, int rank, int nranks);
int mgb_comm_destroy(mgb_comm* c);
/* MGmol_MPI::allreduce(double*, n, MPI_SUM) at src/ExtendedGridOrbitals.cc:
 * 1746 and ReplicatedMatrix::consolidate (src/ReplicatedMatrix.cc:108-127):
 * in-place sum of a device array over all ranks.                            */
int mgb_allreduce_sum_f64(mgb_comm* c, double* data, size_t n, void* stream);
/* Stream-ordered barrier over the ranks (a one-element all-reduce).            */
int mgb_comm_barrier(mgb_comm* c, void* stream);
/* Synchronises the device and reports whether a neighbour barrier of the in-place
 * halo paths (a flag exchange through peer memory, used instead of a collective
 * because a halo only depends on the two x neighbours) ever timed out.        */
int mgb_comm_check(mgb_comm* c);
/* Direct peer reads over NVLink.  Every rank registers ITS array (collective:
 * each rank passes its own pointer, e.g. its orbital block) and the library
 * maps the other ranks' arrays into this process with CUDA IPC
 * (cudaIpcGetMemHandle / cudaIpcOpenMemHandle, peer access over NVLink /
 * NVSwitch).  Kernels then read the neighbours' boundary planes in place:
 * this replaces the pack / MPI_Isend / unpack of initiate/finishEastWestComm
 * (src/pb/GridFuncVector.cc:1195-1256) for arrays that stay resident.
 * Returns MGB_ENOTSUP if an array is not in IPC-exportable (cudaMalloc)
 * memory; the packed exchange below then remains available.                 */
int mgb_peer_register(mgb_comm* c, const void* ptr, void* stream);
int mgb_peer_unregister(mgb_comm* c, const void* ptr);
/* LocGridOrbitals on an x-split domain: the reference's halo packets are
 * addressed by global orbital id, not by color slot -- the sender tags every
 * face with gid_[0][color] (west-bound) / gid_[nsubdivx-1][color] (east-bound)
 * and the receiver stores it in ITS slot of that gid, if its boundary slab
 * holds it (src/pb/GridFuncVector.cc:1225-1246, 1374-1419).  For the in-place
 * peer reads this is a per-color index translation: map_west[c] / map_east[c]
 * = the color slot on the west / east rank whose boundary slab holds the
 * orbital of my color c in my boundary slab, or -1 (then the ghost planes read
 * as zero: what a freshly reset ghosted block keeps in the reference).  Host
 * arrays of ncolors ints; NULL, NULL restores the identity
 * (ExtendedGridOrbitals).  The caller builds them from its own and its
 * neighbours' overlapping_gids (one MPI_Sendrecv of ncolors ints).           */
int mgb_peer_set_color_maps(mgb_comm* c, const int* map_west, const int* map_east,
    int ncolors);
/* mgb_hpsi on an x-split domain with the neighbours' boundary planes of phi
 * read directly from their registered blocks (no x-halo buffer for phi; V's
 * halo xhalo_v is exchanged once per potential update with
 * mgb_halo_exchange_x).  Collective: rank barrier, fused kernel, rank barrier. */
int mgb_hpsi_peer(mgb_comm* c, int lap_type, int dtype, const mgb_grid* grid,
    const void* phi, size_t ld, const double* vtot, void* hphi, size_t ldh,
    int nfunc, const double* xhalo_v, void* stream);
/* The same on ANY px x py x pz decomposition (pb::PEenv::geom, src/pb/PEenv.cc:335-, gives
 * 2 x 2 x 2 for a cubic grid on 8 ranks): the g planes, rows and columns outside the box --
 * and the edge lines the Mehrstellen stencil reaches -- are fetched by the kernel's TMA
 * producer from the blocks of the up to 26 neighbouring ranks, in place over NVLink
 * (GridFuncVector::trade_boundaries' N/S, U/D and E/W exchanges, src/pb/GridFuncVector.cc:
 * 232-315, 748-830, 1195-1256, without any packed copy).  phi must be registered
 * (mgb_peer_register) and have the same shape on every rank; orbitals keep their slots
 * (ExtendedGridOrbitals).  vghost: the potential as a ghosted double field of the stencil's
 * width whose boundaries were traded (mgb_gfv_set_with_ghosts + mgb_halo_exchange_ghosted once
 * per potential update: gfpot of src/Hamiltonian.cc:108-111); may be null for MGB_LAP_4,
 * which multiplies by V at the centre only.  comm may be null on a single rank (every
 * neighbour is the box itself: the periodic wrap).                                     */
/* With MGB_HPSI_TIMING set in the environment: mean phase times (z push, barrier, kernel,
 * barrier) of the last mgb_hpsi_peer3d calls of this process, to stderr.  Development aid. */
void mgb_hpsi_timing_report(int rank);
int mgb_hpsi_peer3d(mgb_comm* comm, int lap_type, int dtype, const mgb_grid* grid,
    const void* phi, size_t ld, const double* vtot, const double* vghost, void* hphi, size_t ldh,
    int nfunc, void* stream);
/* x-direction halo for the fused H path: sends this rank's first/last g
 * planes of every function to the west/east neighbours and fills
 * xhalo[nfunc][2g][ny][nz] (replaces initiate/finishEastWestComm,
 * src/pb/GridFuncVector.cc:1195-1256, for no-ghost blocks).                 */
int mgb_halo_exchange_x(mgb_comm* c, int dtype, const mgb_grid* grid, int g,
    const void* noghost, size_t ld, void* xhalo, int nfunc, void* stream);
/* GridFuncVector::trade_boundaries for a ghosted block on a general
 * px x py x pz decomposition: Y, then Z, then X faces (src/pb/
 * GridFuncVector.cc:1544-1622) so edges and corners propagate.              */
int mgb_halo_exchange_ghosted(mgb_comm* c, int dtype, const mgb_grid* grid,
    void* ghosted, int nfunc, void* stream);

/* The same gid addressing for the packed exchange (LocGridOrbitals on any
 * decomposition): maps[dir][side][iloc][color] (dir 0 x, 1 y, 2 z; side 0 = the
 * ghosts filled by the low neighbour, 1 = by the high neighbour; host array of
 * 6*subdivx*ncolors ints) = the color of the SENDING rank whose x-slab iloc holds
 * the orbital of my color in that slab, or -1: that slab of the ghost layer is
 * left as it is, exactly as the reference's receiver skips it
 * (src/pb/GridFuncVector.cc:461-513 north/south, :958-998 up/down, :1374-1419
 * east/west; for x only iloc 0 (side 0) / subdivx-1 (side 1) is read).  NULL
 * restores slot-for-slot copies.                                             */
int mgb_halo_set_color_maps(mgb_comm* c, int subdivx, int ncolors, const int* maps);

#ifdef __cplusplus
}
#endif

#endif /* MGMOL_B200_H */
