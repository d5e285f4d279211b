// C-ABI glue: error state, device memory (MemorySpace::Memory<T,Device>), the
// mgb_hpsi dispatcher and its reference-shaped composition on ghosted blocks.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "hpsi.h"

namespace mgb
{

std::atomic<unsigned long long> g_launch_count{ 0 };
static thread_local char g_err[512] = "";
static int g_last_path  = 0;
static int g_force_path = 0;

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int require_device()
{
    static int state = 0; // 0 unknown, 1 ok, -1 none
    if (state == 0)
    {
        int n         = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        state         = (e == cudaSuccess && n > 0) ? 1 : -1;
        if (state < 0) (void)cudaGetLastError();
    }
    if (state < 0)
    {
        set_error("no CUDA device: mgmol_b200 has no CPU fallback");
        return MGB_ENODEVICE;
    }
    return MGB_OK;
}

int check_grid(const mgb_grid* gr)
{
    MGB_REQUIRE(gr != nullptr, "grid is null");
    for (int d = 0; d < 3; d++)
    {
        MGB_REQUIRE(gr->dim[d] > 0 && gr->dim[d] < 10000,
            "grid dim[%d]=%d out of range", d, gr->dim[d]); // pb/Grid.cc:33-35
        MGB_REQUIRE(gr->h[d] > 1.e-8, "grid h[%d] too small", d);
        MGB_REQUIRE(gr->bc[d] == 0 || gr->bc[d] == 1,
            "bc[%d]=%d: only periodic (1) and Dirichlet-0 (0) orbitals' "
            "boundary conditions are supported",
            d, gr->bc[d]);
        MGB_REQUIRE(gr->nproc[d] >= 1 && gr->coord[d] >= 0
                        && gr->coord[d] < gr->nproc[d],
            "bad process coordinates in direction %d", d);
    }
    MGB_REQUIRE(gr->ghosts >= 0 && gr->ghosts < 10, "bad ghost width %d",
        gr->ghosts); // pb/Grid.h:80
    return MGB_OK;
}

// grow-only scratch slots
static void* g_scratch[12]     = { nullptr };
static size_t g_scratch_sz[12] = { 0 };
static std::mutex g_scratch_mu;
void* scratch(int slot, size_t bytes)
{
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    if (bytes <= g_scratch_sz[slot]) return g_scratch[slot];
    if (g_scratch[slot]) cudaFree(g_scratch[slot]); // implicit device sync
    g_scratch[slot]    = nullptr;
    g_scratch_sz[slot] = 0;
    void* p            = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess)
    {
        set_error("scratch allocation of %zu bytes failed", bytes);
        (void)cudaGetLastError();
        return nullptr;
    }
    g_scratch[slot]    = p;
    g_scratch_sz[slot] = bytes;
    return p;
}

// Path 3: the reference's own sequence (src/Hamiltonian.cc:101-156) on ghosted
// blocks, built from the bit-exact batch kernels.  Used for mixed boundary
// conditions and boxes the fused kernels do not take; y/z splits additionally
// need mgb_halo_exchange_ghosted between set_with_ghosts and the stencils and
// are driven by the host wrapper, not from here.
int hpsi_ghosted(const HpsiArgs& a, cudaStream_t st)
{
    const mgb_grid* gr0 = a.grid;
    MGB_REQUIRE(gr0->nproc[0] == 1 && gr0->nproc[1] == 1 && gr0->nproc[2] == 1,
        "mgb_hpsi: the ghosted-block composition handles single-rank boxes; "
        "for split directions exchange halos with mgb_halo_exchange_ghosted "
        "and call the mgb_gfv_* / mgb_fd_apply entry points");
    mgb_grid gr = *gr0;
    gr.ghosts   = a.g;
    Box b       = box_of(&gr, a.g);
    const size_t es   = a.dtype == MGB_F64 ? 8 : 4;
    // sub-buffers on 16-byte boundaries (the double potential block follows three float blocks)
    const size_t blk  = ((size_t)b.sizeg * a.nfunc * es + 15) & ~(size_t)15;
    void* stream      = (void*)st;
    unsigned char* ws = (unsigned char*)scratch(1, 3 * blk + (size_t)b.sizeg * 8);
    if (!ws) return MGB_ECUDA;
    void* gphi   = ws;
    void* w1     = ws + blk;
    void* work1  = ws + 2 * blk;
    double* gpot = (double*)(ws + 3 * blk);
    int rc;
    // phi.setDataWithGhosts(); phi.trade_boundaries();            (:101-102)
    if ((rc = mgb_gfv_set_with_ghosts(
             a.dtype, a.dtype, &gr, a.phi, a.ld, gphi, a.nfunc, stream)))
        return rc;
    if ((rc = mgb_gfv_trade_boundaries(a.dtype, &gr, gphi, a.nfunc, stream)))
        return rc;
    if (a.lap_type == MGB_LAP_4)
        // per-orbital Laph4::applyWithPot                          (:138-155)
        return del2_4th_withpot(
            a.dtype, &gr, gphi, a.vtot, a.hphi, a.ldh, a.nfunc, st);
    // gfpot.assign(vtot); gfpot.trade_boundaries();               (:108-111)
    if ((rc = mgb_gfv_set_with_ghosts(
             MGB_F64, MGB_F64, &gr, a.vtot, (size_t)b.npt, gpot, 1, stream)))
        return rc;
    if ((rc = mgb_gfv_trade_boundaries(MGB_F64, &gr, gpot, 1, stream))) return rc;
    // gfvw1.pointwiseProduct(gfvphi, gfpot)                        (:117)
    if ((rc = mgb_gfv_pointwise_product(a.dtype, &gr, gphi, gpot, w1, a.nfunc, stream)))
        return rc;
    // gfv_work1 (fresh, zeroed) = B * gfvw1                        (:119-122)
    if ((rc = mgb_memset(work1, 0, blk, stream))) return rc;
    if ((rc = mgb_fd_apply(
             MGB_FD_RHS_4TH_MEHR1, a.dtype, &gr, w1, work1, a.nfunc, a.g, stream)))
        return rc;
    // gfvw1 = -Lap phi                                             (:127)
    if ((rc = mgb_fd_apply(
             MGB_FD_DEL2_4TH_MEHR, a.dtype, &gr, gphi, w1, a.nfunc, 0, stream)))
        return rc;
    // gfv_work1.axpy(1., gfvw1)                                    (:129)
    if ((rc = mgb_axpy(a.dtype, (size_t)b.sizeg * a.nfunc, 1., w1, work1, stream)))
        return rc;
    // hphi.setPsi(gfv_work1)                                       (:131)
    return mgb_gfv_get_values(
        a.dtype, a.dtype, &gr, work1, a.hphi, a.ldh, a.nfunc, stream);
}

} // namespace mgb

using namespace mgb;

extern "C"
{

const char* mgb_last_error(void) { return g_err; }
int mgb_version(void) { return 100; }
unsigned long long mgb_launch_count(void)
{
    return g_launch_count.load(std::memory_order_relaxed);
}
int mgb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

int mgb_malloc(void** dptr, size_t bytes)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(dptr, "mgb_malloc: null out pointer");
    MGB_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return MGB_OK;
}
int mgb_free(void* dptr)
{
    if (!dptr) return MGB_OK;
    MGB_CUDA(cudaFree(dptr));
    return MGB_OK;
}
int mgb_copy_to_dev(void* dst, const void* src, size_t bytes, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
    return MGB_OK;
}
int mgb_copy_to_host(void* dst, const void* src, size_t bytes, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
    return MGB_OK;
}
int mgb_copy_dev(void* dst, const void* src, size_t bytes, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return MGB_OK;
}
int mgb_memset(void* dptr, int value, size_t bytes, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_CUDA(cudaMemsetAsync(dptr, value, bytes, as_stream(stream)));
    return MGB_OK;
}
int mgb_stream_sync(void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return MGB_OK;
}

int mgb_hpsi_last_path(void) { return g_last_path; }
const char* mgb_hpsi_last_kernel(void) { return g_last_path == 1 ? hpsi_last_kernel() : (g_last_path == 2 ? "k_hpsi_generic" : "ghosted composition"); }
int mgb_hpsi_force_path(int path)
{
    MGB_REQUIRE(path >= 0 && path <= 3, "mgb_hpsi_force_path: path %d", path);
    g_force_path = path;
    return MGB_OK;
}

static int hpsi_entry(int lap_type, int dtype, const mgb_grid* grid, const void* phi,
    size_t ld, const double* vtot, void* hphi, size_t ldh, int nfunc,
    const void* xhalo_phi, const double* xhalo_v, const void* peer_w, const void* peer_e,
    const int* map_w, const int* map_e, void* stream, const void* const* nb3d = nullptr,
    const double* vghost = nullptr, const void* const* nbz = nullptr)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(grid)) return rc;
    MGB_REQUIRE(phi && vtot && hphi, "mgb_hpsi: null pointer");
    MGB_REQUIRE(phi != hphi, "mgb_hpsi: hphi must not alias phi");
    MGB_REQUIRE(dtype == MGB_F32 || dtype == MGB_F64, "mgb_hpsi: bad dtype %d", dtype);
    MGB_REQUIRE(nfunc >= 0, "mgb_hpsi: nfunc < 0");
    // Only Laph4M, Laph4MP and Laph4 can apply H: the other operators have no
    // applyWithPot and the reference aborts (src/pb/Lap.h:35-39).
    MGB_REQUIRE(lap_type == MGB_LAP_4M || lap_type == MGB_LAP_4MP
                    || lap_type == MGB_LAP_4,
        "mgb_hpsi: lap_type %d has no applyWithPot (Lap::applyWithPot aborts)",
        lap_type);
    const size_t npt = (size_t)grid->dim[0] * grid->dim[1] * grid->dim[2];
    MGB_REQUIRE(ld >= npt && ldh >= npt, "mgb_hpsi: leading dimension < npt");
    if (nfunc == 0) return MGB_OK;

    HpsiArgs a;
    a.lap_type  = lap_type;
    a.dtype     = dtype;
    a.grid      = grid;
    a.g         = (lap_type == MGB_LAP_4) ? 2 : 1;
    a.phi       = phi;
    a.ld        = ld;
    a.vtot      = vtot;
    a.hphi      = hphi;
    a.ldh       = ldh;
    a.nfunc     = nfunc;
    a.xhalo_phi = xhalo_phi;
    a.xhalo_v   = xhalo_v;
    a.peer_w    = peer_w;
    a.peer_e    = peer_e;
    a.map_w     = map_w;
    a.map_e     = map_e;
    a.nb3d      = nb3d;
    a.vghost    = vghost;
    a.nbz       = nbz;
    cudaStream_t st = as_stream(stream);
    if (nb3d)
    {
        // every halo read in place from the neighbours' blocks: the TMA kernel only
        MGB_REQUIRE(grid->bc[0] == grid->bc[1] && grid->bc[1] == grid->bc[2],
            "mgb_hpsi_peer3d: mixed boundary conditions are served by the ghosted-block entry points");
        const int rc = hpsi_tma(a, st);
        if (rc == MGB_OK) g_last_path = 1;
        if (rc == MGB_ENOTSUP)
            set_error("mgb_hpsi_peer3d: box not eligible for the TMA kernel (nz <= 256, z rows a "
                      "multiple of 32 bytes, 16-byte aligned blocks)");
        return rc;
    }

    const bool uniform_bc = (grid->bc[0] == grid->bc[1] && grid->bc[1] == grid->bc[2]);
    const bool yz_single  = grid->nproc[1] == 1 && grid->nproc[2] == 1;
    MGB_REQUIRE(nb3d || grid->nproc[0] == 1 || ((xhalo_phi || (peer_w && peer_e)) && xhalo_v),
        "mgb_hpsi: x is split over %d ranks but no x-halo buffers were given",
        grid->nproc[0]);
    MGB_REQUIRE(grid->dim[0] >= a.g && grid->dim[1] >= a.g && grid->dim[2] >= a.g,
        "mgb_hpsi: local dims smaller than the stencil radius");

    if (uniform_bc && yz_single)
    {
        if (g_force_path == 0 || g_force_path == 1)
        {
            int rc = hpsi_tma(a, st);
            if (rc == MGB_OK)
            {
                g_last_path = 1;
                return MGB_OK;
            }
            if (rc != MGB_ENOTSUP) return rc;
            if (g_force_path == 1)
            {
                set_error("mgb_hpsi: box not eligible for the TMA kernel");
                return MGB_ENOTSUP;
            }
        }
        if (g_force_path == 0 || g_force_path == 2)
        {
            g_last_path = 2;
            return hpsi_generic(a, st);
        }
    }
    if (g_force_path == 1 || g_force_path == 2)
    {
        set_error("mgb_hpsi: forced path %d cannot serve this box", g_force_path);
        return MGB_ENOTSUP;
    }
    MGB_REQUIRE(yz_single,
        "mgb_hpsi: y/z are split over ranks; use the ghosted-block entry "
        "points with mgb_halo_exchange_ghosted");
    g_last_path = 3;
    return hpsi_ghosted(a, st);
}

int mgb_apply_b(int lap_type, int dtype, const mgb_grid* grid, const void* phi, size_t ld,
    void* bphi, size_t ldb, int nfunc, const void* xhalo_phi, void* stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(grid)) return rc;
    MGB_REQUIRE(phi && bphi && phi != bphi, "mgb_apply_b: null or aliased pointers");
    MGB_REQUIRE(dtype == MGB_F32 || dtype == MGB_F64, "mgb_apply_b: bad dtype %d", dtype);
    const size_t npt = (size_t)grid->dim[0] * grid->dim[1] * grid->dim[2];
    MGB_REQUIRE(ld >= npt && ldb >= npt && nfunc >= 0, "mgb_apply_b: bad dimensions");
    if (nfunc == 0) return MGB_OK;
    const size_t es = dtype == MGB_F64 ? 8 : 4;
    cudaStream_t st = as_stream(stream);
    if (lap_type != MGB_LAP_4M && lap_type != MGB_LAP_4MP)
    {
        // FDoper::rhs of the non-compact operators is the identity
        // (src/pb/FDoper.h:68-79: B = 1)
        MGB_CUDA(cudaMemcpy2DAsync(bphi, ldb * es, phi, ld * es, npt * es, (size_t)nfunc,
            cudaMemcpyDeviceToDevice, st));
        return MGB_OK;
    }
    const bool uniform = grid->bc[0] == grid->bc[1] && grid->bc[1] == grid->bc[2];
    MGB_REQUIRE(uniform && grid->nproc[1] == 1 && grid->nproc[2] == 1,
        "mgb_apply_b: mixed boundary conditions and y/z splits go through "
        "mgb_gfv_set_with_ghosts + trade + mgb_fd_apply(MGB_FD_RHS_4TH_MEHR1)");
    MGB_REQUIRE(grid->nproc[0] == 1 || xhalo_phi,
        "mgb_apply_b: x is split but no x-halo buffer was given");
    // Laph4M: B (rhs_4th_Mehr1); Laph4MP: B2 (rhs_4th_Mehr2, src/pb/Laph4MP.h:48-51)
    return rhs_generic(
        dtype, lap_type == MGB_LAP_4MP, grid, phi, ld, xhalo_phi, bphi, ldb, nfunc, st);
}

int mgb_residual(int lap_type, int dtype, const mgb_grid* grid, const void* phi, size_t ld,
    const void* hphi, size_t ldh, const double* theta, int ldt, void* res, size_t ldr,
    int nfunc, const void* xhalo_phi, void* stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(grid)) return rc;
    MGB_REQUIRE(phi && hphi && theta && res, "mgb_residual: null pointer");
    MGB_REQUIRE(res != phi && res != hphi, "mgb_residual: res must not alias phi or hphi");
    MGB_REQUIRE(dtype == MGB_F32 || dtype == MGB_F64, "mgb_residual: bad dtype %d", dtype);
    const size_t npt = (size_t)grid->dim[0] * grid->dim[1] * grid->dim[2];
    MGB_REQUIRE(ld >= npt && ldh >= npt && ldr >= npt && ldt >= nfunc && nfunc >= 0,
        "mgb_residual: bad dimensions");
    if (nfunc == 0) return MGB_OK;
    cudaStream_t st = as_stream(stream);
    const void* a   = phi;
    size_t lda      = ld;
    if (lap_type == MGB_LAP_4M || lap_type == MGB_LAP_4MP)
    {
        // applyB (ct.Mehrstellen()): tmp = B phi  (src/MGmol.cc:1250-1260)
        const size_t es = dtype == MGB_F64 ? 8 : 4;
        void* tmp       = scratch(6, npt * (size_t)nfunc * es);
        if (!tmp) return MGB_ECUDA;
        if (int rc = mgb_apply_b(lap_type, dtype, grid, phi, ld, tmp, npt, nfunc, xhalo_phi, stream))
            return rc;
        a   = tmp;
        lda = npt;
    }
    // res = (B phi) theta - hphi: multiplyByMatrix(localT, res) then
    // res.axpy(-1., hphi)  (:1272, :1284), one pass
    return gemm_nn_fused(dtype, npt, nfunc, nfunc, 1., a, lda, theta, ldt, 0., res, ldr, -1.,
        hphi, ldh, st);
}

int mgb_hpsi(int lap_type, int dtype, const mgb_grid* grid, const void* phi,
    size_t ld, const double* vtot, void* hphi, size_t ldh, int nfunc,
    const void* xhalo_phi, const double* xhalo_v, void* stream)
{
    return hpsi_entry(lap_type, dtype, grid, phi, ld, vtot, hphi, ldh, nfunc, xhalo_phi,
        xhalo_v, nullptr, nullptr, nullptr, nullptr, stream);
}

int mgb_hpsi_peer(mgb_comm* comm, int lap_type, int dtype, const mgb_grid* grid,
    const void* phi, size_t ld, const double* vtot, void* hphi, size_t ldh, int nfunc,
    const double* xhalo_v, void* stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(grid)) return rc;
    MGB_REQUIRE(comm && phi, "mgb_hpsi_peer: null pointer");
    MGB_REQUIRE(grid->nproc[0] > 1 && grid->nproc[1] == 1 && grid->nproc[2] == 1,
        "mgb_hpsi_peer: the direct peer path serves x-split domains");
    MGB_REQUIRE(grid->dim[0] * grid->nproc[0] == grid->gdim[0],
        "mgb_hpsi_peer: x must be split evenly (neighbours' blocks have my shape)");
    const int west = comm_rank_of(grid, grid->coord[0] - 1, grid->coord[1], grid->coord[2]);
    const int east = comm_rank_of(grid, grid->coord[0] + 1, grid->coord[1], grid->coord[2]);
    const void* pw = peer_view(comm, phi, west);
    const void* pe = peer_view(comm, phi, east);
    if (!pw || !pe)
    {
        set_error("mgb_hpsi_peer: phi is not registered with mgb_peer_register (or the "
                  "neighbours' blocks cannot be mapped)");
        return MGB_ENOTSUP;
    }
    const int *map_w = nullptr, *map_e = nullptr;
    int map_n = 0;
    comm_color_maps(comm, &map_w, &map_e, &map_n);
    MGB_REQUIRE(!map_w || map_n >= nfunc,
        "mgb_hpsi_peer: color maps cover %d colors, the block has %d", map_n, nfunc);
    cudaStream_t st = as_stream(stream);
    // every rank's phi is complete before anybody reads boundary planes ...
    if (int rc = comm_barrier_neighbors(comm, grid, st)) return rc;
    const int force = g_force_path;
    g_force_path    = 1; // only the TMA kernel reads peers
    const int rc    = hpsi_entry(lap_type, dtype, grid, phi, ld, vtot, hphi, ldh, nfunc,
        nullptr, xhalo_v, pw, pe, map_w, map_e, stream);
    g_force_path = force;
    if (rc) return rc;
    // ... and nobody overwrites its phi while a neighbour still reads it
    return comm_barrier_neighbors(comm, grid, st);
}


static cudaEvent_t g_tev[32][5];
static unsigned long long g_tcall = 0;

int mgb_hpsi_peer3d(mgb_comm* comm, int lap_type, int dtype, const mgb_grid* grid,
    const void* phi, size_t ld, const double* vtot, const double* vghost, void* hphi, size_t ldh,
    int nfunc, void* stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(grid)) return rc;
    MGB_REQUIRE(phi, "mgb_hpsi_peer3d: null pointer");
    const int np = grid->nproc[0] * grid->nproc[1] * grid->nproc[2];
    MGB_REQUIRE(comm || np == 1, "mgb_hpsi_peer3d: a decomposed grid needs the communicator");
    for (int d = 0; d < 3; d++)
        MGB_REQUIRE(grid->dim[d] * grid->nproc[d] == grid->gdim[d],
            "mgb_hpsi_peer3d: direction %d must be split evenly (neighbours' blocks have my shape)", d);
    MGB_REQUIRE(lap_type == MGB_LAP_4 || vghost,
        "mgb_hpsi_peer3d: the Mehrstellen operator needs the ghosted copy of the potential");
    // the blocks of the 26 neighbours (myself across a direction that is not split; null
    // beyond a non-periodic end -- never read: those boxes are out of range by coordinate)
    const void* nb[27];
    for (int dx = -1; dx <= 1; dx++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dz = -1; dz <= 1; dz++)
            {
                const int i = ((dx + 1) * 3 + (dy + 1)) * 3 + (dz + 1);
                const int r = comm_rank_of(
                    grid, grid->coord[0] + dx, grid->coord[1] + dy, grid->coord[2] + dz);
                nb[i] = (np == 1) ? phi : peer_view(comm, phi, r);
                if (!nb[i])
                {
                    set_error("mgb_hpsi_peer3d: phi is not registered with mgb_peer_register (or "
                              "rank %d's block cannot be mapped)", r);
                    return MGB_ENOTSUP;
                }
            }
    cudaStream_t st = as_stream(stream);
    // MGB_HPSI_TIMING: phase times of the last calls (push, barrier, kernel, barrier), printed
    // by mgb_hpsi_timing_report -- development aid for multi-rank runs
    static const bool timing = getenv("MGB_HPSI_TIMING") != nullptr;
    cudaEvent_t* ev = nullptr;
    if (timing)
    {
        if (!g_tev[0][0])
            for (auto& row : g_tev)
                for (auto& e : row)
                    cudaEventCreate(&e);
        ev = g_tev[g_tcall++ % 32];
        cudaEventRecord(ev[0], st);
    }
    // z split: the z-edge columns are pushed into the z neighbours' dense buffers first
    const void* nbz[9];
    // MGB_ZHALO=inplace: the kernel reads the columns from the z neighbours' blocks itself
    const char* zmode = getenv("MGB_ZHALO");
    const bool zsplit = (grid->nproc[2] > 1 || getenv("MGB_HPSI_FORCE_ZBOXES") != nullptr)
                        && !(zmode && zmode[0] == 'i');
    if (zsplit)
        if (int rc = comm_zhalo(comm, grid, dtype == MGB_F64 ? 8 : 4, phi, ld, nfunc, st, nbz))
            return rc;
    // every rank's phi is complete (and its columns have landed) before anybody reads
    // boundary layers ...
    if (ev) cudaEventRecord(ev[1], st);
    if (int rc = comm_barrier_neighbors(comm, grid, st)) return rc;
    if (ev) cudaEventRecord(ev[2], st);
    const int rc = hpsi_entry(lap_type, dtype, grid, phi, ld, vtot, hphi, ldh, nfunc, nullptr,
        nullptr, nullptr, nullptr, nullptr, nullptr, stream, nb, vghost, zsplit ? nbz : nullptr);
    if (rc) return rc;
    if (ev) cudaEventRecord(ev[3], st);
    // ... and nobody overwrites its phi while a neighbour still reads it
    const int rc2 = comm_barrier_neighbors(comm, grid, st);
    if (ev) cudaEventRecord(ev[4], st);
    return rc2;
}

void mgb_hpsi_timing_report(int rank)
{
    if (!g_tev[0][0] || g_tcall < 4) return;
    cudaDeviceSynchronize();
    const int n = g_tcall < 32 ? (int)g_tcall : 32;
    double acc[4] = { 0, 0, 0, 0 };
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 4; k++)
        {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, g_tev[i][k], g_tev[i][k + 1]);
            acc[k] += ms / n;
        }
    fprintf(stderr, "[mgb timing] rank %d: z push %.3f ms, barrier %.3f ms, kernel %.3f ms, barrier %.3f ms "
                    "(mean of the last %d mgb_hpsi_peer3d calls)\n", rank, acc[0], acc[1], acc[2], acc[3], n);
}

/* ---- host-buffer entry: H2D copy, fused kernel and D2H copy pipelined over
 * blocks of orbitals on three streams (full-duplex PCIe) ------------------- */
int mgb_host_register(void* ptr, size_t bytes)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(ptr && bytes, "mgb_host_register: null pointer or zero size");
    MGB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return MGB_OK;
}
int mgb_host_unregister(void* ptr)
{
    if (!ptr) return MGB_OK;
    MGB_CUDA(cudaHostUnregister(ptr));
    return MGB_OK;
}

static int hpsi_host_impl(mgb_comm* comm, int lap_type, int dtype, const mgb_grid* grid,
    const void* phi_host, size_t ld, const double* vtot_host, void* hphi_host, size_t ldh,
    int nfunc, int chunk)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(grid)) return rc;
    MGB_REQUIRE(phi_host && vtot_host && hphi_host, "mgb_hpsi_host: null pointer");
    MGB_REQUIRE(dtype == MGB_F32 || dtype == MGB_F64, "mgb_hpsi_host: bad dtype %d", dtype);
    MGB_REQUIRE(nfunc >= 0 && chunk >= 0, "mgb_hpsi_host: negative count");
    const bool split = grid->nproc[0] > 1;
    MGB_REQUIRE(grid->nproc[1] == 1 && grid->nproc[2] == 1 && (!split || comm),
        "mgb_hpsi_host: the host-buffer entry serves single-rank boxes and, with a "
        "communicator (mgb_hpsi_host_peer), x-split domains");
    MGB_REQUIRE(!split || grid->dim[0] * grid->nproc[0] == grid->gdim[0],
        "mgb_hpsi_host_peer: x must be split evenly");
    const size_t npt = (size_t)grid->dim[0] * grid->dim[1] * grid->dim[2];
    MGB_REQUIRE(ld >= npt && ldh >= npt, "mgb_hpsi_host: leading dimension < npt");
    if (nfunc == 0) return MGB_OK;

    constexpr int NS = 3; // ring slots per direction
    static cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    static cudaEvent_t ev_in[NS], ev_k[NS], ev_out[NS];
    if (!s_in)
    {
        MGB_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
        MGB_CUDA(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking));
        MGB_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        for (int i = 0; i < NS; i++)
        {
            MGB_CUDA(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
            MGB_CUDA(cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming));
            MGB_CUDA(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
        }
    }
    const size_t es = dtype == MGB_F64 ? 8 : 4;
    // padded so that every slot and every orbital inside it is 16-byte aligned
    const size_t ldd = (npt + 3) / 4 * 4;
    if (chunk == 0)
    {
        // ~32 MB per block: still at full PCIe rate, and the pipeline fill and
        // drain (one block in, one block out) cost 2 / (number of blocks)
        chunk = (int)((size_t)(32u << 20) / (ldd * es));
        if (chunk < 1) chunk = 1;
    }
    if (chunk > nfunc) chunk = nfunc;
    const size_t slot_bytes = (size_t)chunk * ldd * es;
    const int g             = (lap_type == MGB_LAP_4) ? 2 : 1;
    const size_t plane      = (size_t)grid->dim[1] * grid->dim[2];
    unsigned char* din_old  = (unsigned char*)scratch(3, 0);
    unsigned char* din  = (unsigned char*)scratch(3, NS * slot_bytes);
    unsigned char* dout = (unsigned char*)scratch(4, NS * slot_bytes);
    double* dv          = (double*)scratch(5, (npt + 2 * g * plane) * sizeof(double));
    if (!din || !dout || !dv) return MGB_ECUDA;
    MGB_CUDA(cudaMemcpyAsync(dv, vtot_host, npt * sizeof(double), cudaMemcpyHostToDevice, s_in));
    double* xv = dv + npt; // x halo of the potential (2g planes)
    const unsigned char *din_w = nullptr, *din_e = nullptr;
    if (split)
    {
        // the input ring is what the neighbours read in place: publish it (again
        // if it was re-allocated; collective -- every rank runs the same sizes)
        if (din_old && din != din_old) mgb_peer_unregister(comm, din_old); // ring was re-sized
        if (din != din_old || !peer_view(comm, din, comm_rank_of(grid, grid->coord[0], 0, 0)))
            if (int rc = mgb_peer_register(comm, din, (void*)s_in)) return rc;
        din_w = (const unsigned char*)peer_view(
            comm, din, comm_rank_of(grid, grid->coord[0] - 1, grid->coord[1], grid->coord[2]));
        din_e = (const unsigned char*)peer_view(
            comm, din, comm_rank_of(grid, grid->coord[0] + 1, grid->coord[1], grid->coord[2]));
        if (!din_w || !din_e)
        {
            set_error("mgb_hpsi_host_peer: the neighbours' input rings cannot be mapped");
            return MGB_ENOTSUP;
        }
        if (int rc = mgb_halo_exchange_x(comm, MGB_F64, grid, g, dv, npt, xv, 1, (void*)s_in))
            return rc;
    }

    const int nchunks = (nfunc + chunk - 1) / chunk;
    for (int i = 0; i < nchunks; i++)
    {
        const int slot = i % NS;
        const int f0   = i * chunk;
        const int nf   = (nfunc - f0 < chunk) ? nfunc - f0 : chunk;
        unsigned char* in  = din + (size_t)slot * slot_bytes;
        unsigned char* out = dout + (size_t)slot * slot_bytes;
        // the kernel that last read this input slot must be done
        if (i >= NS) MGB_CUDA(cudaStreamWaitEvent(s_in, ev_k[slot], 0));
        if (ld == ldd && ldd == npt) // contiguous: one linear copy
            MGB_CUDA(cudaMemcpyAsync(in, (const unsigned char*)phi_host + (size_t)f0 * ld * es,
                (size_t)nf * npt * es, cudaMemcpyHostToDevice, s_in));
        else
            MGB_CUDA(cudaMemcpy2DAsync(in, ldd * es,
                (const unsigned char*)phi_host + (size_t)f0 * ld * es, ld * es, npt * es,
                (size_t)nf, cudaMemcpyHostToDevice, s_in));
        MGB_CUDA(cudaEventRecord(ev_in[slot], s_in));
        MGB_CUDA(cudaStreamWaitEvent(s_k, ev_in[slot], 0));
        // the copy-out that last read this output slot must be done
        if (i >= NS) MGB_CUDA(cudaStreamWaitEvent(s_k, ev_out[slot], 0));
        if (!split)
        {
            if (int rc = mgb_hpsi(lap_type, dtype, grid, in, ldd, dv, out, ldd, nf, nullptr,
                    nullptr, (void*)s_k))
                return rc;
        }
        else
        {
            // every rank's block i has landed in its slot -> fused kernel reading
            // the neighbours' slots in place -> nobody refills a slot early
            if (int rc = comm_barrier_neighbors(comm, grid, s_k)) return rc;
            const int force = g_force_path;
            g_force_path    = 1;
            const int rc    = hpsi_entry(lap_type, dtype, grid, in, ldd, dv, out, ldd, nf, nullptr,
                xv, din_w + (size_t)slot * slot_bytes, din_e + (size_t)slot * slot_bytes, nullptr,
                nullptr, (void*)s_k);
            g_force_path = force;
            if (rc) return rc;
            if (int rc2 = comm_barrier_neighbors(comm, grid, s_k)) return rc2;
        }
        MGB_CUDA(cudaEventRecord(ev_k[slot], s_k));
        MGB_CUDA(cudaStreamWaitEvent(s_out, ev_k[slot], 0));
        if (ldh == ldd && ldd == npt)
            MGB_CUDA(cudaMemcpyAsync((unsigned char*)hphi_host + (size_t)f0 * ldh * es, out,
                (size_t)nf * npt * es, cudaMemcpyDeviceToHost, s_out));
        else
            MGB_CUDA(cudaMemcpy2DAsync((unsigned char*)hphi_host + (size_t)f0 * ldh * es, ldh * es,
                out, ldd * es, npt * es, (size_t)nf, cudaMemcpyDeviceToHost, s_out));
        MGB_CUDA(cudaEventRecord(ev_out[slot], s_out));
    }
    MGB_CUDA(cudaStreamSynchronize(s_out));
    MGB_CUDA(cudaStreamSynchronize(s_k));
    MGB_CUDA(cudaStreamSynchronize(s_in));
    return MGB_OK;
}

int mgb_hpsi_host(int lap_type, int dtype, const mgb_grid* grid, const void* phi_host,
    size_t ld, const double* vtot_host, void* hphi_host, size_t ldh, int nfunc,
    int chunk)
{
    MGB_REQUIRE(grid && grid->nproc[0] == 1,
        "mgb_hpsi_host: x is split; use mgb_hpsi_host_peer with the communicator");
    return hpsi_host_impl(nullptr, lap_type, dtype, grid, phi_host, ld, vtot_host, hphi_host,
        ldh, nfunc, chunk);
}

int mgb_hpsi_host_peer(mgb_comm* comm, int lap_type, int dtype, const mgb_grid* grid,
    const void* phi_host, size_t ld, const double* vtot_host, void* hphi_host, size_t ldh,
    int nfunc, int chunk)
{
    MGB_REQUIRE(comm, "mgb_hpsi_host_peer: null communicator");
    return hpsi_host_impl(comm, lap_type, dtype, grid, phi_host, ld, vtot_host, hphi_host, ldh,
        nfunc, chunk);
}

} // extern "C"
