// Multi-GPU exchange for the hot path: one process per GPU, the 3-D block
// decomposition of pb::PEenv (src/pb/PEenv.cc:56-139), NCCL over NVLink for
//   * the x-direction halo planes feeding the fused H kernel,
//   * the Y -> Z -> X ghost exchange of a ghosted block
//     (GridFuncVector::trade_boundaries, src/pb/GridFuncVector.cc:1544-1622),
//   * the all-reduce of partial N x N matrices
//     (MGmol_MPI::allreduce at src/ExtendedGridOrbitals.cc:1746).
// NCCL is loaded with dlopen so that the library also loads on a single-GPU
// host that has no NCCL (under PyTorch the already-loaded libnccl.so.2 is
// reused).
#include <cuda.h>
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "hpsi.h"

// minimal NCCL surface (matches nccl.h 2.x; no header dependency at build)
typedef struct ncclComm* ncclComm_t;
typedef struct
{
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
enum
{
    kNcclInt8    = 0,
    kNcclFloat32 = 7,
    kNcclFloat64 = 8,
    kNcclSum     = 0
};

namespace mgb
{
struct Nccl
{
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*)                                  = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)           = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t)                                     = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t,
        cudaStream_t)                                                           = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t,
        cudaStream_t)                                                           = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t)     = nullptr;
    ncclResult_t (*GroupStart)()                                                = nullptr;
    ncclResult_t (*GroupEnd)()                                                  = nullptr;
    const char* (*GetErrorString)(ncclResult_t)                                 = nullptr;
};

static Nccl* nccl()
{
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char* nm : names)
        {
            n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (n.handle) break;
        }
        if (!n.handle) return;
#define MGB_SYM(field, name) *(void**)(&n.field) = dlsym(n.handle, name)
        MGB_SYM(GetUniqueId, "ncclGetUniqueId");
        MGB_SYM(CommInitRank, "ncclCommInitRank");
        MGB_SYM(CommDestroy, "ncclCommDestroy");
        MGB_SYM(AllReduce, "ncclAllReduce");
        MGB_SYM(AllGather, "ncclAllGather");
        MGB_SYM(Send, "ncclSend");
        MGB_SYM(Recv, "ncclRecv");
        MGB_SYM(GroupStart, "ncclGroupStart");
        MGB_SYM(GroupEnd, "ncclGroupEnd");
        MGB_SYM(GetErrorString, "ncclGetErrorString");
#undef MGB_SYM
    });
    if (!n.handle || !n.GetUniqueId || !n.CommInitRank || !n.AllReduce || !n.Send
        || !n.Recv || !n.GroupStart || !n.GroupEnd)
        return nullptr;
    return &n;
}

#define MGB_NCCL(call)                                                         \
    do                                                                         \
    {                                                                          \
        ncclResult_t r__ = (call);                                             \
        if (r__ != 0)                                                          \
        {                                                                      \
            ::mgb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,     \
                nccl()->GetErrorString ? nccl()->GetErrorString(r__) : "?");   \
            return MGB_ENCCL;                                                  \
        }                                                                      \
    } while (0)

} // namespace mgb

// one array registered for direct peer reads: what every rank published
struct PeerEntry
{
    std::vector<cudaIpcMemHandle_t> handle; // per rank: IPC handle of the allocation
    std::vector<unsigned long long> offset; // per rank: offset of the array inside it
    std::vector<void*> mapped;              // per rank: the array in MY address space
    std::vector<int> opened_id;             // per rank: which opened allocation (-1: none)
    size_t my_bytes;                        // bytes from my array's start to the end of its allocation
};

// one peer allocation opened with cudaIpcOpenMemHandle, shared by every
// registered array that lives in it
struct OpenedAlloc
{
    int rank;
    cudaIpcMemHandle_t handle;
    void* base;
    int refs;
};

struct mgb_comm
{
    ncclComm_t comm;
    int rank, nranks;
    void* buf[4];
    size_t buf_sz[4];
    float* flag;                                  // 1-element all-reduce = rank barrier
    // neighbour barrier over peer memory: inbox[r] = epoch last signalled by rank r
    // (only my Cartesian neighbours ever write theirs)
    unsigned long long* inbox;
    unsigned long long** d_peer_inbox; // device table [nranks]: the neighbours' inboxes mapped here
    int nb_key[9];                     // nproc, coord, bc the table was built for
    int nb_mode;                       // 0 not agreed yet, 1 peer flags, 2 NCCL all-reduce (every rank)
    int nb_count;                      // distinct neighbour ranks
    unsigned long long epoch;
    // gid-addressed packed exchange: hmaps[dir][side][iloc][color] = the sending
    // neighbour's color whose slab holds my color's orbital, or -1 (device)
    int* hmaps;
    int hm_subdivx, hm_ncolors;
    int* map_w;                                   // my color -> west / east rank's color
    int* map_e;                                   // (device, ncolors each) or null = same
    int map_n;
    // z-halo columns of a decomposed box: what my z neighbours pushed here (see comm_zhalo)
    void* zbuf;
    size_t zbuf_bytes;
    std::map<const void*, PeerEntry>* peers;      // local array -> peer views
    std::vector<OpenedAlloc>* opened; // slots are reused, never compacted
};

namespace mgb
{
static void* comm_buf(mgb_comm* c, int i, size_t bytes)
{
    if (bytes <= c->buf_sz[i]) return c->buf[i];
    if (c->buf[i]) cudaFree(c->buf[i]);
    c->buf[i]    = nullptr;
    c->buf_sz[i] = 0;
    if (cudaMalloc(&c->buf[i], bytes) != cudaSuccess)
    {
        set_error("comm buffer allocation of %zu bytes failed", bytes);
        (void)cudaGetLastError();
        return nullptr;
    }
    c->buf_sz[i] = bytes;
    return c->buf[i];
}

// rank of the process at Cartesian coordinates (row-major, as
// MPI_Cart_create orders them; src/pb/PEenv.cc:89)
static int rank_of(const mgb_grid* gr, int cx, int cy, int cz)
{
    const int px = gr->nproc[0], py = gr->nproc[1], pz = gr->nproc[2];
    cx = (cx + px) % px;
    cy = (cy + py) % py;
    cz = (cz + pz) % pz;
    return (cx * py + cy) * pz + cz;
}

// first / last g planes of every function of a no-ghost block -> packed
// [nfunc][g][ny*nz] send buffers (one launch packs both sides)
template <typename T>
__global__ void k_pack_x(int nx, long long plane, int g, long long ld,
    const T* __restrict__ u, T* __restrict__ lo, T* __restrict__ hi)
{
    const long long per = (long long)g * plane;
    const long long t   = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= per) return;
    const int f   = blockIdx.y;
    const T* src  = u + (long long)f * ld;
    lo[(long long)f * per + t] = src[t];                                  // x = 0..g-1
    hi[(long long)f * per + t] = src[(long long)(nx - g) * plane + t];    // x = nx-g..nx-1
}
// received buffers -> xhalo[nfunc][2g][ny*nz] (g planes below, g planes above)
template <typename T>
__global__ void k_unpack_x(long long plane, int g, const T* __restrict__ from_west,
    const T* __restrict__ from_east, T* __restrict__ xhalo, int have_w, int have_e)
{
    const long long per = (long long)g * plane;
    const long long t   = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= per) return;
    const int f = blockIdx.y;
    T* dst      = xhalo + (long long)f * 2 * per;
    dst[t]       = have_w ? from_west[(long long)f * per + t] : (T)0;
    dst[per + t] = have_e ? from_east[(long long)f * per + t] : (T)0;
}

template <typename T>
static int halo_x_t(mgb_comm* c, const mgb_grid* gr, int g, const T* u, size_t ld,
    T* xhalo, int nfunc, cudaStream_t st)
{
    Nccl* N = nccl();
    const int nx = gr->dim[0];
    const long long plane = (long long)gr->dim[1] * gr->dim[2];
    const size_t cnt      = (size_t)nfunc * g * plane;
    const size_t bytes    = cnt * sizeof(T);
    T* s_w = (T*)comm_buf(c, 0, bytes);
    T* s_e = (T*)comm_buf(c, 1, bytes);
    T* r_w = (T*)comm_buf(c, 2, bytes);
    T* r_e = (T*)comm_buf(c, 3, bytes);
    if (!s_w || !s_e || !r_w || !r_e) return MGB_ECUDA;
    const bool periodic = gr->bc[0] == 1;
    const bool have_w   = periodic || gr->coord[0] > 0;
    const bool have_e   = periodic || gr->coord[0] < gr->nproc[0] - 1;
    const int west = rank_of(gr, gr->coord[0] - 1, gr->coord[1], gr->coord[2]);
    const int east = rank_of(gr, gr->coord[0] + 1, gr->coord[1], gr->coord[2]);
    const int dt   = sizeof(T) == 8 ? kNcclFloat64 : kNcclFloat32;
    for (int f0 = 0; f0 < nfunc; f0 += 65535)
    {
        const int nf = (nfunc - f0 < 65535) ? nfunc - f0 : 65535;
        dim3 grid((unsigned)(((long long)g * plane + 255) / 256), (unsigned)nf);
        k_pack_x<T><<<grid, 256, 0, st>>>(nx, plane, g, (long long)ld,
            u + (size_t)f0 * ld, s_w + (size_t)f0 * g * plane,
            s_e + (size_t)f0 * g * plane);
        MGB_LAUNCHED("k_pack_x");
    }
    // Sends in the order west, east; receives in the order east, west: with two
    // ranks both neighbours are the same peer and NCCL pairs the k-th send with
    // the peer's k-th receive.
    MGB_NCCL(N->GroupStart());
    if (have_w) MGB_NCCL(N->Send(s_w, cnt, dt, west, c->comm, st));
    if (have_e) MGB_NCCL(N->Send(s_e, cnt, dt, east, c->comm, st));
    if (have_e) MGB_NCCL(N->Recv(r_e, cnt, dt, east, c->comm, st));
    if (have_w) MGB_NCCL(N->Recv(r_w, cnt, dt, west, c->comm, st));
    MGB_NCCL(N->GroupEnd());
    for (int f0 = 0; f0 < nfunc; f0 += 65535)
    {
        const int nf = (nfunc - f0 < 65535) ? nfunc - f0 : 65535;
        dim3 grid((unsigned)(((long long)g * plane + 255) / 256), (unsigned)nf);
        k_unpack_x<T><<<grid, 256, 0, st>>>(plane, g, r_w + (size_t)f0 * g * plane,
            r_e + (size_t)f0 * g * plane, xhalo + (size_t)f0 * 2 * g * plane,
            have_w ? 1 : 0, have_e ? 1 : 0);
        MGB_LAUNCHED("k_unpack_x");
    }
    return MGB_OK;
}

// Unpack a received face gid-addressed (src/pb/GridFuncVector.cc:461-513,
// 958-998, 1374-1419): my color f takes, x-slab by x-slab, the face of the
// sender's color map[iloc][f]; slabs whose orbital the sender does not hold
// keep their ghost values, as in the reference.
template <typename T>
__global__ void k_unpack_mapped(Box b, int lo0, int lo1, int lo2, int e0, int e1, int e2,
    T* __restrict__ u, const T* __restrict__ buf, const int* __restrict__ map, int ncolors,
    int subdivx, int sub0, int fixed_iloc)
{
    const long long per = (long long)e0 * e1 * e2;
    const long long t   = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= per) return;
    const int f = blockIdx.y;
    const int k = (int)(t % e2);
    const int j = (int)((t / e2) % e1);
    const int i = (int)(t / ((long long)e2 * e1));
    int iloc    = fixed_iloc;
    if (iloc < 0)
    {
        iloc = (lo0 + i - b.g) / sub0; // faces of the y / z exchange span interior x only
        if (iloc < 0) iloc = 0;
        if (iloc >= subdivx) iloc = subdivx - 1;
    }
    const int src = map[iloc * ncolors + f];
    if (src < 0) return;
    u[(long long)f * b.sizeg + (long long)(lo0 + i) * b.incx + (long long)(lo1 + j) * b.incy
        + (lo2 + k)]
        = buf[(long long)src * per + t];
}

static int unpack_mapped(mgb_comm* c, int dtype, const mgb_grid* gr, const int lo[3],
    const int ext[3], void* u, const void* buf, int nfunc, int d, int side, cudaStream_t st)
{
    Box b = box_of(gr, gr->ghosts);
    const long long per = (long long)ext[0] * ext[1] * ext[2];
    if (per == 0 || nfunc == 0) return MGB_OK;
    MGB_REQUIRE(nfunc <= c->hm_ncolors && nfunc <= 65535,
        "gid-addressed exchange: %d functions, color maps cover %d", nfunc, c->hm_ncolors);
    const int* map = c->hmaps + (size_t)(d * 2 + side) * c->hm_subdivx * c->hm_ncolors;
    const int sub0 = gr->dim[0] / c->hm_subdivx;
    const int fixed = d == 0 ? (side == 0 ? 0 : c->hm_subdivx - 1) : -1;
    dim3 grid((unsigned)((per + 255) / 256), (unsigned)nfunc, 1);
    if (dtype == MGB_F64)
        k_unpack_mapped<double><<<grid, 256, 0, st>>>(b, lo[0], lo[1], lo[2], ext[0], ext[1],
            ext[2], (double*)u, (const double*)buf, map, c->hm_ncolors, c->hm_subdivx, sub0, fixed);
    else
        k_unpack_mapped<float><<<grid, 256, 0, st>>>(b, lo[0], lo[1], lo[2], ext[0], ext[1],
            ext[2], (float*)u, (const float*)buf, map, c->hm_ncolors, c->hm_subdivx, sub0, fixed);
    MGB_LAUNCHED("k_unpack_mapped");
    return MGB_OK;
}

// One direction of the ghosted exchange: faces with the reference's extents
// (Y: interior x and z; Z: interior x, all y; X: whole planes).
static int exchange_dir(mgb_comm* c, int dtype, const mgb_grid* gr, void* u,
    int nfunc, int d, cudaStream_t st)
{
    Nccl* N     = nccl();
    const int g = gr->ghosts;
    const int n[3]  = { gr->dim[0], gr->dim[1], gr->dim[2] };
    // extents of a face slab in ghosted coordinates
    int lo[3], ext[3];
    lo[0] = g, ext[0] = n[0];
    lo[1] = g, ext[1] = n[1];
    lo[2] = g, ext[2] = n[2];
    if (d == 2) lo[1] = 0, ext[1] = n[1] + 2 * g;
    if (d == 0) lo[1] = 0, ext[1] = n[1] + 2 * g, lo[2] = 0, ext[2] = n[2] + 2 * g;
    ext[d] = g;
    const size_t es  = dtype == MGB_F64 ? 8 : 4;
    const size_t cnt = (size_t)nfunc * ext[0] * ext[1] * ext[2];
    void* s_lo = comm_buf(c, 0, cnt * es); // my low interior layers -> low neighbour
    void* s_hi = comm_buf(c, 1, cnt * es);
    void* r_lo = comm_buf(c, 2, cnt * es); // from low neighbour -> my low ghosts
    void* r_hi = comm_buf(c, 3, cnt * es);
    if (!s_lo || !s_hi || !r_lo || !r_hi) return MGB_ECUDA;
    const bool periodic = gr->bc[d] == 1;
    const bool have_lo  = periodic || gr->coord[d] > 0;
    const bool have_hi  = periodic || gr->coord[d] < gr->nproc[d] - 1;
    int cl[3] = { gr->coord[0], gr->coord[1], gr->coord[2] };
    int ch[3] = { gr->coord[0], gr->coord[1], gr->coord[2] };
    cl[d] -= 1;
    ch[d] += 1;
    const int rlo = rank_of(gr, cl[0], cl[1], cl[2]);
    const int rhi = rank_of(gr, ch[0], ch[1], ch[2]);
    const int dt  = dtype == MGB_F64 ? kNcclFloat64 : kNcclFloat32;
    int rc;
    int l[3] = { lo[0], lo[1], lo[2] };
    l[d]     = g; // first interior layers
    if ((rc = subbox_copy(dtype, true, gr, l, ext, u, s_lo, nfunc, st))) return rc;
    l[d] = n[d]; // last interior layers (ghosted index n+g-g)
    if ((rc = subbox_copy(dtype, true, gr, l, ext, u, s_hi, nfunc, st))) return rc;
    MGB_NCCL(N->GroupStart());
    if (have_lo) MGB_NCCL(N->Send(s_lo, cnt, dt, rlo, c->comm, st));
    if (have_hi) MGB_NCCL(N->Send(s_hi, cnt, dt, rhi, c->comm, st));
    if (have_hi) MGB_NCCL(N->Recv(r_hi, cnt, dt, rhi, c->comm, st));
    if (have_lo) MGB_NCCL(N->Recv(r_lo, cnt, dt, rlo, c->comm, st));
    MGB_NCCL(N->GroupEnd());
    l[d] = 0; // low ghosts
    if (have_lo)
    {
        rc = c->hmaps ? unpack_mapped(c, dtype, gr, l, ext, u, r_lo, nfunc, d, 0, st)
                      : subbox_copy(dtype, false, gr, l, ext, u, r_lo, nfunc, st);
        if (rc) return rc;
    }
    l[d] = n[d] + g; // high ghosts
    if (have_hi)
    {
        rc = c->hmaps ? unpack_mapped(c, dtype, gr, l, ext, u, r_hi, nfunc, d, 1, st)
                      : subbox_copy(dtype, false, gr, l, ext, u, r_hi, nfunc, st);
        if (rc) return rc;
    }
    return MGB_OK;
}

// the array `local` of rank `rank` as seen from this process (nullptr if it
// was not registered or cannot be mapped)
const void* peer_view(mgb_comm* c, const void* local, int rank)
{
    if (!c || !c->peers) return nullptr;
    auto it = c->peers->find(local);
    if (it == c->peers->end())
    {
        // a position inside a registered array (e.g. a ring slot of a registered
        // workspace): same displacement inside the neighbour's array
        for (auto& kv : *c->peers)
        {
            const char* base = (const char*)kv.first;
            if ((const char*)local > base && (const char*)local < base + kv.second.my_bytes)
            {
                const void* pb = peer_view(c, kv.first, rank);
                return pb ? (const char*)pb + ((const char*)local - base) : nullptr;
            }
        }
        return nullptr;
    }
    PeerEntry& e = it->second;
    if (rank == c->rank) return local;
    if (e.mapped[rank]) return e.mapped[rank];
    int id = -1;
    for (size_t i = 0; i < c->opened->size(); i++)
    {
        OpenedAlloc& o = (*c->opened)[i];
        if (o.refs > 0 && o.rank == rank
            && memcmp(&o.handle, &e.handle[rank], sizeof(cudaIpcMemHandle_t)) == 0)
            id = (int)i;
    }
    if (id < 0)
    {
        void* base = nullptr;
        if (cudaIpcOpenMemHandle(&base, e.handle[rank], cudaIpcMemLazyEnablePeerAccess)
            != cudaSuccess)
        {
            (void)cudaGetLastError();
            return nullptr;
        }
        OpenedAlloc o;
        o.rank   = rank;
        o.handle = e.handle[rank];
        o.base   = base;
        o.refs   = 0;
        for (size_t i = 0; i < c->opened->size() && id < 0; i++)
            if ((*c->opened)[i].refs == 0) id = (int)i;
        if (id < 0)
        {
            c->opened->push_back(o);
            id = (int)c->opened->size() - 1;
        }
        else
            (*c->opened)[id] = o;
    }
    (*c->opened)[id].refs++;
    e.opened_id[rank] = id;
    e.mapped[rank]    = (char*)(*c->opened)[id].base + e.offset[rank];
    return e.mapped[rank];
}

int comm_barrier(mgb_comm* c, cudaStream_t st)
{
    if (!c || c->nranks == 1) return MGB_OK;
    MGB_NCCL(nccl()->AllReduce(c->flag, c->flag, 1, kNcclFloat32, kNcclSum, c->comm, st));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return MGB_OK;
}

int comm_rank_of(const mgb_grid* gr, int cx, int cy, int cz) { return rank_of(gr, cx, cy, cz); }

// Barrier with the Cartesian neighbours only (every rank within one step in each
// split direction: the ranks a halo -- faces, and the edges / corners the
// Mehrstellen stencil reaches -- can come from), through peer memory: publish my
// epoch in every neighbour's inbox (remote stores over NVLink), then wait until
// every neighbour has published it in mine.  One tiny kernel instead of a
// collective.  The wait FAILS CLOSED: after ~10 s without the neighbour it
// records the epoch in the flag slot and traps, which poisons the context --
// every later call of this process returns MGB_ECUDA instead of computing on
// stale halo planes.
__global__ void k_neighbor_sync(unsigned long long* inbox, unsigned long long* const* peer_inbox,
    int myrank, int nranks, unsigned long long epoch)
{
    const int r = threadIdx.x;
    if (r >= nranks) return;
    unsigned long long* dst = peer_inbox[r];
    if (!dst) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst + myrank), "l"(epoch) : "memory");
    const long long t0 = clock64();
    unsigned long long v;
    do
    {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(inbox + r) : "memory");
        if (v < epoch && clock64() - t0 > 20000000000LL)
        {
            inbox[nranks] = epoch; // which barrier (readable only from a debugger: the trap ends the context)
            __threadfence_system();
            __trap();
        }
    } while (v < epoch);
}

// the distinct ranks within one step of `gr`'s coordinates in the split directions
static void neighbour_ranks(const mgb_grid* gr, int myrank, std::vector<int>& out)
{
    out.clear();
    for (int dx = -1; dx <= 1; dx++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dz = -1; dz <= 1; dz++)
            {
                const int d[3] = { dx, dy, dz };
                bool ok        = dx || dy || dz;
                for (int k = 0; k < 3 && ok; k++)
                {
                    if (!d[k]) continue;
                    if (gr->nproc[k] == 1) ok = false; // not split: own data
                    const int c = gr->coord[k] + d[k];
                    if (gr->bc[k] != 1 && (c < 0 || c >= gr->nproc[k])) ok = false; // domain ends
                }
                if (!ok) continue;
                const int r = rank_of(gr, gr->coord[0] + dx, gr->coord[1] + dy, gr->coord[2] + dz);
                if (r == myrank) continue;
                bool seen = false;
                for (int q : out)
                    seen = seen || q == r;
                if (!seen) out.push_back(r);
            }
}

int comm_barrier_neighbors(mgb_comm* c, const mgb_grid* gr, cudaStream_t st)
{
    if (!c || c->nranks == 1) return MGB_OK;
    if (getenv("MGB_NCCL_BARRIER")) return comm_barrier(c, st);
    int key[9];
    for (int k = 0; k < 3; k++)
    {
        key[k]     = gr->nproc[k];
        key[3 + k] = gr->coord[k];
        key[6 + k] = gr->bc[k];
    }
    if (c->nb_mode == 0 || memcmp(key, c->nb_key, sizeof(key)) != 0)
    {
        // collective, once per decomposition: every rank publishes its inbox, maps
        // its neighbours', and all ranks agree on ONE kind of barrier (a rank that
        // fell back to the all-reduce alone would leave its neighbours spinning)
        int ok = 1;
        if (!c->inbox)
        {
            const size_t bytes = (size_t)(c->nranks + 1) * sizeof(unsigned long long);
            MGB_CUDA(cudaMalloc(&c->inbox, bytes));
            MGB_CUDA(cudaMemset(c->inbox, 0, bytes));
            MGB_CUDA(cudaMalloc(&c->d_peer_inbox, c->nranks * sizeof(void*)));
            const int rc = mgb_peer_register(c, c->inbox, (void*)st);
            if (rc == MGB_ENOTSUP)
                ok = 0;
            else if (rc)
                return rc;
        }
        std::vector<int> nb;
        neighbour_ranks(gr, c->rank, nb);
        std::vector<unsigned long long*> tab(c->nranks, nullptr);
        for (int r : nb)
        {
            tab[r] = ok ? (unsigned long long*)peer_view(c, c->inbox, r) : nullptr;
            if (!tab[r]) ok = 0;
        }
        // agreement (and: nobody signals before every inbox exists and is zero)
        float* flag = (float*)comm_buf(c, 0, sizeof(float));
        if (!flag) return MGB_ECUDA;
        const float mine = ok ? 0.f : 1.f;
        float total      = 0.f;
        MGB_CUDA(cudaMemcpyAsync(flag, &mine, sizeof(float), cudaMemcpyHostToDevice, st));
        MGB_NCCL(nccl()->AllReduce(flag, flag, 1, kNcclFloat32, kNcclSum, c->comm, st));
        MGB_CUDA(cudaMemcpyAsync(&total, flag, sizeof(float), cudaMemcpyDeviceToHost, st));
        MGB_CUDA(cudaStreamSynchronize(st));
        c->nb_mode = total == 0.f ? 1 : 2;
        memcpy(c->nb_key, key, sizeof(key));
        c->nb_count = (int)nb.size();
        if (c->nb_mode == 1)
            MGB_CUDA(cudaMemcpyAsync(c->d_peer_inbox, tab.data(), c->nranks * sizeof(void*),
                cudaMemcpyHostToDevice, st));
        MGB_CUDA(cudaStreamSynchronize(st)); // tab is a local
    }
    if (c->nb_mode == 2) return comm_barrier(c, st);
    if (c->nb_count == 0) return MGB_OK;
    c->epoch++;
    k_neighbor_sync<<<1, ((c->nranks + 31) / 32) * 32, 0, st>>>(
        c->inbox, c->d_peer_inbox, c->rank, c->nranks, c->epoch);
    MGB_LAUNCHED("k_neighbor_sync");
    return MGB_OK;
}

// ---------------------------------------------------------------------------
// z-halo columns.  z is the contiguous direction: the G columns a stencil needs from the z
// neighbour are 16 bytes out of every row, and fetching them in place costs one NVLink
// transaction per row (measured: -30 % on a z-split box).  Instead every rank PUSHES the
// first and the last 16 bytes of each of its rows into its z neighbours' dense buffers
// ([side][function][x][y][16 bytes], posted remote stores), and the stencil kernels take
// the columns from the local buffer -- or, for the halo rows / planes of the Mehrstellen
// edges, from the buffers of the x / y neighbours -- with dense TMA boxes.
// ---------------------------------------------------------------------------
__global__ void k_zpack(long long rows_per_f, int nfunc, int nz_bytes, long long ld_bytes,
    const unsigned char* __restrict__ src, unsigned char* __restrict__ low_nb,
    unsigned char* __restrict__ high_nb, unsigned char* __restrict__ mine, int zero_lo, int zero_hi)
{
    const long long total = rows_per_f * nfunc;
    const uint4 zero      = make_uint4(0u, 0u, 0u, 0u);
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < total;
         r += (long long)gridDim.x * blockDim.x)
    {
        const long long f = r / rows_per_f, q = r - f * rows_per_f;
        const unsigned char* row = src + f * ld_bytes + q * nz_bytes;
        // my first columns are the columns ABOVE my low neighbour's box (second half of its
        // entry), my last ones the columns BELOW my high neighbour's (first half)
        if (low_nb) *reinterpret_cast<uint4*>(low_nb + r * 32 + 16) = *reinterpret_cast<const uint4*>(row);
        if (high_nb)
            *reinterpret_cast<uint4*>(high_nb + r * 32)
                = *reinterpret_cast<const uint4*>(row + nz_bytes - 16);
        // beyond a Dirichlet end of the domain the columns are zero
        if (zero_lo) *reinterpret_cast<uint4*>(mine + r * 32) = zero;
        if (zero_hi) *reinterpret_cast<uint4*>(mine + r * 32 + 16) = zero;
    }
}

static void* g_zbuf_single = nullptr; // one rank (test hook): no communicator to own it
static size_t g_zbuf_single_bytes = 0;

// Push the z-edge columns of src (nfunc functions of `es`-byte elements, leading dimension
// ld) to the z neighbours and return in nbz[(dx+1)*3 + (dy+1)] the buffer of the rank at
// coord + (dx, dy, 0) as mapped here: [function][x][y][32 bytes], the 16 bytes below z = 0
// then the 16 bytes above z = nz-1.  The caller places a neighbour barrier between this and
// the consuming kernel, and one after that kernel (nobody pushes into a buffer that is
// still being read).
int comm_zhalo(mgb_comm* c, const mgb_grid* gr, int es, const void* src, size_t ld, int nfunc,
    cudaStream_t st, const void* nbz[9])
{
    const long long rows = (long long)gr->dim[0] * gr->dim[1];
    const size_t bytes   = (size_t)rows * nfunc * 32;
    const bool multi     = c && c->nranks > 1;
    void** buf           = multi ? &c->zbuf : &g_zbuf_single;
    size_t* cap          = multi ? &c->zbuf_bytes : &g_zbuf_single_bytes;
    if (*cap < bytes)
    {
        // collective on a decomposed box: every rank runs the same sizes
        if (*buf)
        {
            // every rank closes its mappings of the old buffers before anybody frees one
            if (multi)
            {
                mgb_peer_unregister(c, *buf);
                if (int rc = comm_barrier(c, st)) return rc;
            }
            MGB_CUDA(cudaStreamSynchronize(st));
            cudaFree(*buf);
            *buf = nullptr;
            *cap = 0;
        }
        MGB_CUDA(cudaMalloc(buf, bytes));
        MGB_CUDA(cudaMemsetAsync(*buf, 0, bytes, st));
        *cap = bytes;
        if (multi)
            if (int rc = mgb_peer_register(c, *buf, (void*)st)) return rc;
    }
    unsigned char* lowv  = nullptr;
    unsigned char* highv = nullptr;
    for (int dx = -1; dx <= 1; dx++)
        for (int dy = -1; dy <= 1; dy++)
        {
            const int r = rank_of(gr, gr->coord[0] + dx, gr->coord[1] + dy, gr->coord[2]);
            const void* v = multi ? peer_view(c, *buf, r) : *buf;
            if (!v)
            {
                set_error("the z-halo buffer of rank %d cannot be mapped (CUDA IPC)", r);
                return MGB_ENOTSUP;
            }
            nbz[(dx + 1) * 3 + (dy + 1)] = v;
        }
    const bool per = gr->bc[2] == 1;
    const bool has_lo = per || gr->coord[2] > 0, has_hi = per || gr->coord[2] < gr->nproc[2] - 1;
    if (has_lo)
    {
        const int r = rank_of(gr, gr->coord[0], gr->coord[1], gr->coord[2] - 1);
        lowv = multi ? (unsigned char*)peer_view(c, *buf, r) : (unsigned char*)*buf;
        if (!lowv) return MGB_ENOTSUP;
    }
    if (has_hi)
    {
        const int r = rank_of(gr, gr->coord[0], gr->coord[1], gr->coord[2] + 1);
        highv = multi ? (unsigned char*)peer_view(c, *buf, r) : (unsigned char*)*buf;
        if (!highv) return MGB_ENOTSUP;
    }
    const long long total = rows * nfunc;
    const int blocks      = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    k_zpack<<<blocks, 256, 0, st>>>(rows, nfunc, gr->dim[2] * es, (long long)ld * es,
        (const unsigned char*)src, lowv, highv, (unsigned char*)*buf, has_lo ? 0 : 1,
        has_hi ? 0 : 1);
    MGB_LAUNCHED("k_zpack");
    return MGB_OK;
}

void comm_color_maps(mgb_comm* c, const int** map_w, const int** map_e, int* n)
{
    *map_w = c ? c->map_w : nullptr;
    *map_e = c ? c->map_e : nullptr;
    *n     = c ? c->map_n : 0;
}

} // namespace mgb

using namespace mgb;

extern "C"
{

int mgb_comm_barrier(mgb_comm* c, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(c, "mgb_comm_barrier: null communicator");
    return comm_barrier(c, as_stream(stream));
}

int mgb_comm_check(mgb_comm* c)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(c, "mgb_comm_check: null communicator");
    // a neighbour barrier that timed out trapped: the synchronize reports it
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess)
    {
        set_error("mgb_comm_check: %s (a neighbour barrier that times out traps: a rank did "
                  "not reach it)", cudaGetErrorString(e));
        return MGB_ECUDA;
    }
    return MGB_OK;
}

int mgb_peer_register(mgb_comm* c, const void* ptr, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(c && ptr, "mgb_peer_register: null pointer");
    Nccl* N = nccl();
    MGB_REQUIRE(N && N->AllGather, "mgb_peer_register: ncclAllGather not available");
    cudaStream_t st = as_stream(stream);
    struct Rec
    {
        cudaIpcMemHandle_t h;
        unsigned long long off;
        int ok, pad;
    } mine;
    memset(&mine, 0, sizeof(mine));
    CUdeviceptr base = 0;
    size_t size      = 0;
    mine.ok = 1;
    typedef CUresult (*PFN_range)(CUdeviceptr*, size_t*, CUdeviceptr);
    static PFN_range get_range = nullptr;
    if (!get_range)
    {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q)
                == cudaSuccess
            && q == cudaDriverEntryPointSuccess)
            get_range = (PFN_range)fp;
    }
    if (!get_range || get_range(&base, &size, (CUdeviceptr)ptr) != CUDA_SUCCESS)
        mine.ok = 0;
    else if (cudaIpcGetMemHandle(&mine.h, (void*)base) != cudaSuccess)
    {
        // e.g. memory from a virtual-memory (expandable segments) allocator
        (void)cudaGetLastError();
        mine.ok = 0;
    }
    mine.off = (unsigned long long)((CUdeviceptr)ptr - base);
    // publish (collective; every rank must call with its own array)
    Rec* dev = (Rec*)comm_buf(c, 0, sizeof(Rec) * (c->nranks + 1));
    if (!dev) return MGB_ECUDA;
    MGB_CUDA(cudaMemcpyAsync(dev, &mine, sizeof(Rec), cudaMemcpyHostToDevice, st));
    MGB_NCCL(N->AllGather(dev, dev + 1, sizeof(Rec), kNcclInt8, c->comm, st));
    std::vector<Rec> all(c->nranks);
    MGB_CUDA(cudaMemcpyAsync(
        all.data(), dev + 1, sizeof(Rec) * c->nranks, cudaMemcpyDeviceToHost, st));
    MGB_CUDA(cudaStreamSynchronize(st));
    bool ok = true;
    for (auto& r : all)
        ok = ok && r.ok;
    if (!ok)
    {
        set_error("mgb_peer_register: an array is not in CUDA-IPC-exportable memory "
                  "(cudaMalloc); use the packed halo exchange instead");
        return MGB_ENOTSUP;
    }
    PeerEntry e;
    e.handle.resize(c->nranks);
    e.offset.resize(c->nranks);
    e.mapped.assign(c->nranks, nullptr);
    e.opened_id.assign(c->nranks, -1);
    e.my_bytes = size - (size_t)mine.off;
    mgb_peer_unregister(c, ptr); // a stale entry of a freed array at this address
    for (int r = 0; r < c->nranks; r++)
    {
        e.handle[r] = all[r].h;
        e.offset[r] = all[r].off;
    }
    (*c->peers)[ptr] = e;
    return MGB_OK;
}

int mgb_peer_set_color_maps(mgb_comm* c, const int* map_west, const int* map_east, int ncolors)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(c, "mgb_peer_set_color_maps: null communicator");
    if (c->map_w) cudaFree(c->map_w);
    if (c->map_e) cudaFree(c->map_e);
    c->map_w = c->map_e = nullptr;
    c->map_n = 0;
    if (!map_west && !map_east) return MGB_OK;
    MGB_REQUIRE(map_west && map_east && ncolors > 0,
        "mgb_peer_set_color_maps: both maps and ncolors > 0 are needed");
    MGB_CUDA(cudaMalloc(&c->map_w, sizeof(int) * ncolors));
    MGB_CUDA(cudaMalloc(&c->map_e, sizeof(int) * ncolors));
    MGB_CUDA(cudaMemcpy(c->map_w, map_west, sizeof(int) * ncolors, cudaMemcpyHostToDevice));
    MGB_CUDA(cudaMemcpy(c->map_e, map_east, sizeof(int) * ncolors, cudaMemcpyHostToDevice));
    c->map_n = ncolors;
    return MGB_OK;
}

int mgb_halo_set_color_maps(mgb_comm* c, int subdivx, int ncolors, const int* maps)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(c, "mgb_halo_set_color_maps: null communicator");
    if (c->hmaps) cudaFree(c->hmaps);
    c->hmaps      = nullptr;
    c->hm_subdivx = c->hm_ncolors = 0;
    if (!maps) return MGB_OK;
    MGB_REQUIRE(subdivx > 0 && ncolors > 0, "mgb_halo_set_color_maps: bad table shape");
    const size_t n = (size_t)6 * subdivx * ncolors;
    MGB_CUDA(cudaMalloc(&c->hmaps, sizeof(int) * n));
    MGB_CUDA(cudaMemcpy(c->hmaps, maps, sizeof(int) * n, cudaMemcpyHostToDevice));
    c->hm_subdivx = subdivx;
    c->hm_ncolors = ncolors;
    return MGB_OK;
}

int mgb_peer_unregister(mgb_comm* c, const void* ptr)
{
    MGB_REQUIRE(c && ptr, "mgb_peer_unregister: null pointer");
    auto it = c->peers->find(ptr);
    if (it == c->peers->end()) return MGB_OK;
    // close the peer allocations nobody else uses any more (after my kernels
    // that may still read them); the owner may free its array only after every
    // rank has done this (see mgb_precond_destroy)
    bool any = false;
    for (int id : it->second.opened_id)
        any = any || id >= 0;
    if (any) cudaDeviceSynchronize();
    for (int id : it->second.opened_id)
        if (id >= 0 && --(*c->opened)[id].refs == 0)
            cudaIpcCloseMemHandle((*c->opened)[id].base);
    c->peers->erase(it);
    return MGB_OK;
}

int mgb_comm_unique_id(void* id128)
{
    MGB_REQUIRE(id128, "mgb_comm_unique_id: null pointer");
    Nccl* N = nccl();
    if (!N)
    {
        set_error("NCCL (libnccl.so.2) could not be loaded");
        return MGB_ENCCL;
    }
    ncclUniqueId id;
    MGB_NCCL(N->GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return MGB_OK;
}

int mgb_comm_create(mgb_comm** out, const void* id128, int rank, int nranks)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(out && id128, "mgb_comm_create: null pointer");
    MGB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank %d/%d", rank, nranks);
    Nccl* N = nccl();
    if (!N)
    {
        set_error("NCCL (libnccl.so.2) could not be loaded");
        return MGB_ENCCL;
    }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    mgb_comm* c = new mgb_comm();
    memset(c, 0, sizeof(*c));
    c->rank   = rank;
    c->nranks = nranks;
    c->peers  = new std::map<const void*, PeerEntry>();
    c->opened = new std::vector<OpenedAlloc>();
    if (cudaMalloc(&c->flag, sizeof(float)) != cudaSuccess
        || cudaMemset(c->flag, 0, sizeof(float)) != cudaSuccess)
    {
        set_error("mgb_comm_create: device allocation failed");
        delete c->peers;
        delete c->opened;
        delete c;
        return MGB_ECUDA;
    }
    ncclResult_t r = N->CommInitRank(&c->comm, nranks, id, rank);
    if (r != 0)
    {
        set_error("ncclCommInitRank failed: %s",
            N->GetErrorString ? N->GetErrorString(r) : "?");
        cudaFree(c->flag);
        delete c->peers;
        delete c->opened;
        delete c;
        return MGB_ENCCL;
    }
    *out = c;
    return MGB_OK;
}

int mgb_comm_destroy(mgb_comm* c)
{
    if (!c) return MGB_OK;
    Nccl* N = nccl();
    for (int i = 0; i < 4; i++)
        if (c->buf[i]) cudaFree(c->buf[i]);
    if (c->flag) cudaFree(c->flag);
    if (c->inbox)
    {
        mgb_peer_unregister(c, c->inbox);
        cudaFree(c->inbox);
    }
    if (c->d_peer_inbox) cudaFree(c->d_peer_inbox);
    if (c->zbuf)
    {
        mgb_peer_unregister(c, c->zbuf);
        cudaFree(c->zbuf);
    }
    if (c->map_w) cudaFree(c->map_w);
    if (c->map_e) cudaFree(c->map_e);
    if (c->hmaps) cudaFree(c->hmaps);
    if (c->opened)
        for (auto& o : *c->opened)
            if (o.refs > 0) cudaIpcCloseMemHandle(o.base);
    delete c->opened;
    delete c->peers;
    if (N && c->comm) N->CommDestroy(c->comm);
    delete c;
    return MGB_OK;
}

int mgb_allreduce_sum_f64(mgb_comm* c, double* data, size_t n, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(c && data, "mgb_allreduce_sum_f64: null pointer");
    if (n == 0 || c->nranks == 1) return MGB_OK;
    MGB_NCCL(nccl()->AllReduce(
        data, data, n, kNcclFloat64, kNcclSum, c->comm, as_stream(stream)));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return MGB_OK;
}

int mgb_halo_exchange_x(mgb_comm* c, int dtype, const mgb_grid* grid, int g,
    const void* noghost, size_t ld, void* xhalo, int nfunc, void* stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(grid)) return rc;
    MGB_REQUIRE(c && noghost && xhalo, "mgb_halo_exchange_x: null pointer");
    MGB_REQUIRE(g >= 1 && g <= grid->dim[0], "mgb_halo_exchange_x: bad halo width %d", g);
    MGB_REQUIRE(grid->nproc[0] * grid->nproc[1] * grid->nproc[2] == c->nranks,
        "mgb_halo_exchange_x: grid decomposition does not match the communicator");
    MGB_REQUIRE(grid->nproc[0] > 1, "mgb_halo_exchange_x: x is not split");
    if (nfunc == 0) return MGB_OK;
    if (dtype == MGB_F64)
        return halo_x_t<double>(c, grid, g, (const double*)noghost, ld,
            (double*)xhalo, nfunc, as_stream(stream));
    if (dtype == MGB_F32)
        return halo_x_t<float>(c, grid, g, (const float*)noghost, ld, (float*)xhalo,
            nfunc, as_stream(stream));
    set_error("mgb_halo_exchange_x: bad dtype");
    return MGB_EINVAL;
}

int mgb_halo_exchange_ghosted(mgb_comm* c, int dtype, const mgb_grid* grid,
    void* ghosted, int nfunc, void* stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_grid(grid)) return rc;
    MGB_REQUIRE(ghosted, "mgb_halo_exchange_ghosted: null pointer");
    MGB_REQUIRE(dtype == MGB_F32 || dtype == MGB_F64, "bad dtype");
    const int np = grid->nproc[0] * grid->nproc[1] * grid->nproc[2];
    MGB_REQUIRE(np == 1 || (c && np == c->nranks),
        "mgb_halo_exchange_ghosted: decomposition does not match the communicator");
    if (nfunc == 0 || grid->ghosts == 0) return MGB_OK;
    cudaStream_t st = as_stream(stream);
    int rc;
    // zero the Dirichlet ghosts and first low layers  (GridFuncVector.cc:1554-1557)
    if ((rc = trade_dirichlet(dtype, grid, ghosted, nfunc, st))) return rc;
    const int order[3] = { 1, 2, 0 }; // Y (north/south), Z (up/down), X (east/west)
    for (int o = 0; o < 3; o++)
    {
        const int d = order[o];
        if (grid->nproc[d] > 1)
            rc = exchange_dir(c, dtype, grid, ghosted, nfunc, d, st);
        else
            rc = trade_wrap(dtype, grid, ghosted, nfunc, d, st);
        if (rc) return rc;
    }
    return MGB_OK;
}

} // extern "C"
