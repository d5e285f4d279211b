// Non-local Kleinman-Bylander projectors on the device: the two
// grid-sized steps the reference runs right after Hamiltonian::applyLocal
// (src/computeHij.cc:404-455),
//
//   kbpsi[row, f] = vel <beta_row | psi_f>        KBPsiMatrixSparse::computeKBpsi
//                                                 (src/KBPsiMatrixSparse.cc:136-212,
//                                                  src/KBPsiMatrixInterface.cc:20-60)
//   H phi_f += sum_ions sum_p alpha_p beta_p      get_vnlpsi (src/get_vnlpsi.cc:24-87) +
//                                                 computeHnlPhiAndAdd2HPhi
//                                                 (src/computeHij.cc:294-375)
//
// on the sparse projector vectors KBprojectorSparse holds (node list + values per ion,
// src/KBprojectorSparse.h:39-52).  The projector set is committed once per ionic step; both
// kernels then run over all orbitals in one launch each.
//
//   k_kb_psi     one block per (ion, function): gather psi on the ion's nodes (coalesced in
//                the node list, scattered in psi), nproj dot products in double, block
//                reduction.
//   k_kb_vnlpsi  POINT-major: at commit the (ion, node) pairs are sorted by grid point
//                (stable in ion order), so one thread owns one touched grid point of one
//                function, sums the ions' contributions in the reference's order and
//                rounding (zero start, (T) casts of src/KBprojectorSparse.cc:865-908) and
//                adds the sum to H phi -- the reference's memset + per-ion scatter +
//                MPaxpy over the whole grid become one pass over the touched points only.
#include <algorithm>
#include <numeric>
#include <vector>

#include "common.cuh"

struct mgb_kb
{
    int dtype;
    size_t npt;
    bool committed;
    // host staging
    std::vector<long long> node0; // per ion: first node; back = total
    std::vector<int> row0;        // per ion: first projector row; back = total
    std::vector<long long> val0;  // per ion: first value (nproj * size_nl values per ion)
    std::vector<int> nlindex;
    std::vector<unsigned char> proj; // values, dtype
    std::vector<double> coeff;       // per row: kbcoeff * sign
    // device
    long long* d_node0;
    int* d_row0;
    long long* d_val0;
    int* d_nlindex;
    void* d_proj;
    double* d_coeff;
    // point-major view: touched points, their entry ranges, entries = (ion, local node)
    long long npoints, nentries;
    int* d_point;
    long long* d_ent0;
    int* d_ent_ion;
    int* d_ent_node;
    int max_nproj;
};

namespace mgb
{

template <typename T, int MAXP>
__global__ void k_kb_psi(const long long* __restrict__ node0, const int* __restrict__ row0,
    const long long* __restrict__ val0, const int* __restrict__ nlindex,
    const T* __restrict__ proj, double vel, const T* __restrict__ psi, long long ld, int nfunc,
    double* __restrict__ kbpsi)
{
    const int j = blockIdx.x, f = blockIdx.y;
    const long long n0 = node0[j], n = node0[j + 1] - n0;
    const int np = row0[j + 1] - row0[j];
    const int* idx = nlindex + n0;
    const T* pr    = proj + val0[j];
    const T* ps    = psi + (long long)f * ld;
    double s[MAXP];
#pragma unroll
    for (int p = 0; p < MAXP; p++)
        s[p] = 0.;
    for (long long i = threadIdx.x; i < n; i += blockDim.x)
    {
        const double v = (double)ps[idx[i]];
#pragma unroll
        for (int p = 0; p < MAXP; p++)
            if (p < np) s[p] += v * (double)pr[(long long)p * n + i];
    }
    __shared__ double red[MAXP][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int p = 0; p < MAXP; p++)
    {
        double v = s[p];
        for (int o = 16; o > 0; o >>= 1)
            v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[p][w] = v;
    }
    __syncthreads();
    if (w == 0)
    {
        const int nw = blockDim.x >> 5;
#pragma unroll
        for (int p = 0; p < MAXP; p++)
        {
            double v = lane < nw ? red[p][lane] : 0.;
            for (int o = 16; o > 0; o >>= 1)
                v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0 && p < np) kbpsi[(long long)(row0[j] + p) * nfunc + f] = vel * v;
        }
    }
}

template <typename T>
__global__ void k_kb_vnlpsi(long long npoints, const int* __restrict__ point,
    const long long* __restrict__ ent0, const int* __restrict__ ent_ion,
    const int* __restrict__ ent_node, const long long* __restrict__ node0,
    const int* __restrict__ row0, const long long* __restrict__ val0,
    const T* __restrict__ proj, const double* __restrict__ coeff,
    const double* __restrict__ kbpsi, T* __restrict__ out, long long ldo, int nfunc, int add)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= npoints) return;
    const int f = blockIdx.y;
    T v = (T)0; // get_vnlpsi's memset
    for (long long e = ent0[t]; e < ent0[t + 1]; e++)
    {
        const int j = ent_ion[e], i = ent_node[e];
        const long long n = node0[j + 1] - node0[j];
        const int r0 = row0[j], np = row0[j + 1] - r0;
        const T* pr = proj + val0[j] + i;
        if (np == 1)
        {
            // axpySKet: dst += (T)(proj * alpha)
            const double alpha = kbpsi[(long long)r0 * nfunc + f] * coeff[r0];
            v += (T)__dmul_rn((double)pr[0], alpha);
        }
        else
        {
            // axpyKet: work = 0; work += (T)(alpha_p * proj_p) per projector; dst += (T)work
            T w = (T)0;
            for (int p = 0; p < np; p++)
            {
                const double alpha = kbpsi[(long long)(r0 + p) * nfunc + f] * coeff[r0 + p];
                w += (T)__dmul_rn(alpha, (double)pr[(long long)p * n]);
            }
            v += w;
        }
    }
    T* o = out + (long long)f * ldo + point[t];
    // MPaxpy(numpt, 1., vnlpsi, hpsi) (src/computeHij.cc:361-363), or the plain vector
    *o = add ? (T)(*o + v) : v;
}

template <typename T>
__global__ void k_kb_zero_rows(long long n, T* __restrict__ out, long long ldo)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[(long long)blockIdx.y * ldo + t] = (T)0;
}

template <typename V>
static int upload(V** dev, const std::vector<V>& h)
{
    if (*dev) cudaFree(*dev);
    *dev = nullptr;
    MGB_CUDA(cudaMalloc(dev, (h.size() ? h.size() : 1) * sizeof(V)));
    if (h.size())
        MGB_CUDA(cudaMemcpy(*dev, h.data(), h.size() * sizeof(V), cudaMemcpyHostToDevice));
    return MGB_OK;
}

} // namespace mgb

using namespace mgb;

extern "C"
{

int mgb_kb_create(mgb_kb** out, int dtype, size_t npt)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(out, "mgb_kb_create: null pointer");
    MGB_REQUIRE(dtype == MGB_F32 || dtype == MGB_F64, "mgb_kb_create: bad dtype %d", dtype);
    MGB_REQUIRE(npt > 0 && npt < ((size_t)1 << 31), "mgb_kb_create: npt out of range");
    mgb_kb* kb = new mgb_kb();
    kb->dtype  = dtype;
    kb->npt    = npt;
    kb->committed = false;
    kb->node0.push_back(0);
    kb->row0.push_back(0);
    kb->val0.push_back(0);
    kb->d_node0 = nullptr, kb->d_row0 = nullptr, kb->d_val0 = nullptr, kb->d_nlindex = nullptr;
    kb->d_proj = nullptr, kb->d_coeff = nullptr, kb->d_point = nullptr, kb->d_ent0 = nullptr;
    kb->d_ent_ion = nullptr, kb->d_ent_node = nullptr;
    kb->npoints = kb->nentries = 0;
    kb->max_nproj = 0;
    *out = kb;
    return MGB_OK;
}

int mgb_kb_add_ion(mgb_kb* kb, int size_nl, const int* nlindex, int nproj, const void* proj,
    const double* coeff, int* first_row)
{
    MGB_REQUIRE(kb && (size_nl == 0 || (nlindex && proj)) && coeff, "mgb_kb_add_ion: null pointer");
    MGB_REQUIRE(size_nl >= 0 && nproj >= 1 && nproj <= 16,
        "mgb_kb_add_ion: %d nodes, %d projectors (1..16 per ion)", size_nl, nproj);
    for (int i = 0; i < size_nl; i++)
        MGB_REQUIRE(nlindex[i] >= 0 && (size_t)nlindex[i] < kb->npt,
            "mgb_kb_add_ion: node %d outside the box", nlindex[i]);
    const size_t es = kb->dtype == MGB_F64 ? 8 : 4;
    if (first_row) *first_row = kb->row0.back();
    kb->nlindex.insert(kb->nlindex.end(), nlindex, nlindex + size_nl);
    const unsigned char* p = (const unsigned char*)proj;
    kb->proj.insert(kb->proj.end(), p, p + (size_t)nproj * size_nl * es);
    kb->coeff.insert(kb->coeff.end(), coeff, coeff + nproj);
    kb->node0.push_back(kb->node0.back() + size_nl);
    kb->row0.push_back(kb->row0.back() + nproj);
    kb->val0.push_back(kb->val0.back() + (long long)nproj * size_nl);
    if (nproj > kb->max_nproj) kb->max_nproj = nproj;
    kb->committed = false;
    return MGB_OK;
}

int mgb_kb_nrows(const mgb_kb* kb) { return kb ? kb->row0.back() : 0; }

int mgb_kb_commit(mgb_kb* kb)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(kb, "mgb_kb_commit: null handle");
    int rc;
    if ((rc = upload(&kb->d_node0, kb->node0))) return rc;
    if ((rc = upload(&kb->d_row0, kb->row0))) return rc;
    if ((rc = upload(&kb->d_val0, kb->val0))) return rc;
    if ((rc = upload(&kb->d_nlindex, kb->nlindex))) return rc;
    if ((rc = upload(&kb->d_coeff, kb->coeff))) return rc;
    if ((rc = upload((unsigned char**)&kb->d_proj, kb->proj))) return rc;
    // point-major view: entries (ion, local node) sorted by grid point, stable in ion order
    const long long ne = kb->node0.back();
    std::vector<long long> order(ne);
    std::iota(order.begin(), order.end(), 0LL);
    std::stable_sort(order.begin(), order.end(),
        [&](long long a, long long b) { return kb->nlindex[a] < kb->nlindex[b]; });
    std::vector<int> ion_of(ne);
    for (size_t j = 0; j + 1 < kb->node0.size(); j++)
        for (long long e = kb->node0[j]; e < kb->node0[j + 1]; e++)
            ion_of[e] = (int)j;
    std::vector<int> point, ent_ion(ne), ent_node(ne);
    std::vector<long long> ent0;
    for (long long k = 0; k < ne; k++)
    {
        const long long e = order[k];
        if (k == 0 || kb->nlindex[e] != kb->nlindex[order[k - 1]])
        {
            point.push_back(kb->nlindex[e]);
            ent0.push_back(k);
        }
        ent_ion[k]  = ion_of[e];
        ent_node[k] = (int)(e - kb->node0[ion_of[e]]);
    }
    ent0.push_back(ne);
    kb->npoints  = (long long)point.size();
    kb->nentries = ne;
    if ((rc = upload(&kb->d_point, point))) return rc;
    if ((rc = upload(&kb->d_ent0, ent0))) return rc;
    if ((rc = upload(&kb->d_ent_ion, ent_ion))) return rc;
    if ((rc = upload(&kb->d_ent_node, ent_node))) return rc;
    kb->committed = true;
    return MGB_OK;
}

int mgb_kb_destroy(mgb_kb* kb)
{
    if (!kb) return MGB_OK;
    for (void* p : { (void*)kb->d_node0, (void*)kb->d_row0, (void*)kb->d_val0, (void*)kb->d_nlindex,
             kb->d_proj, (void*)kb->d_coeff, (void*)kb->d_point, (void*)kb->d_ent0,
             (void*)kb->d_ent_ion, (void*)kb->d_ent_node })
        if (p) cudaFree(p);
    delete kb;
    return MGB_OK;
}

int mgb_kb_psi(const mgb_kb* kb, int dtype, double vel, const void* psi, size_t ld, int nfunc,
    double* kbpsi, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(kb && psi && kbpsi, "mgb_kb_psi: null pointer");
    MGB_REQUIRE(kb->committed, "mgb_kb_psi: mgb_kb_commit was not called");
    MGB_REQUIRE(dtype == kb->dtype, "mgb_kb_psi: KBPROJDTYPE is ORBDTYPE (src/global.h:38)");
    MGB_REQUIRE(ld >= kb->npt && nfunc >= 0 && nfunc <= 65535, "mgb_kb_psi: bad dimensions");
    const int nions = (int)kb->node0.size() - 1;
    if (nfunc == 0 || nions == 0) return MGB_OK;
    cudaStream_t st = as_stream(stream);
    dim3 grid((unsigned)nions, (unsigned)nfunc);
#define MGB_KB_PSI(T, MP)                                                                     \
    k_kb_psi<T, MP><<<grid, 256, 0, st>>>(kb->d_node0, kb->d_row0, kb->d_val0, kb->d_nlindex,   \
        (const T*)kb->d_proj, vel, (const T*)psi, (long long)ld, nfunc, kbpsi)
    if (dtype == MGB_F64)
    {
        if (kb->max_nproj <= 4)
            MGB_KB_PSI(double, 4);
        else
            MGB_KB_PSI(double, 16);
    }
    else
    {
        if (kb->max_nproj <= 4)
            MGB_KB_PSI(float, 4);
        else
            MGB_KB_PSI(float, 16);
    }
#undef MGB_KB_PSI
    MGB_LAUNCHED("k_kb_psi");
    return MGB_OK;
}

int mgb_kb_vnlpsi(const mgb_kb* kb, int dtype, const double* kbpsi, void* out, size_t ldo,
    int nfunc, int add, void* stream)
{
    if (int rc = require_device()) return rc;
    MGB_REQUIRE(kb && kbpsi && out, "mgb_kb_vnlpsi: null pointer");
    MGB_REQUIRE(kb->committed, "mgb_kb_vnlpsi: mgb_kb_commit was not called");
    MGB_REQUIRE(dtype == kb->dtype, "mgb_kb_vnlpsi: KBPROJDTYPE is ORBDTYPE (src/global.h:38)");
    MGB_REQUIRE(ldo >= kb->npt && nfunc >= 0 && nfunc <= 65535, "mgb_kb_vnlpsi: bad dimensions");
    if (nfunc == 0) return MGB_OK;
    cudaStream_t st = as_stream(stream);
    if (!add)
    {
        // get_vnlpsi's memset of the whole function
        dim3 g((unsigned)((kb->npt + 255) / 256), (unsigned)nfunc);
        if (dtype == MGB_F64)
            k_kb_zero_rows<double><<<g, 256, 0, st>>>((long long)kb->npt, (double*)out, (long long)ldo);
        else
            k_kb_zero_rows<float><<<g, 256, 0, st>>>((long long)kb->npt, (float*)out, (long long)ldo);
        MGB_LAUNCHED("k_kb_zero_rows");
    }
    if (kb->npoints == 0) return MGB_OK;
    dim3 grid((unsigned)((kb->npoints + 255) / 256), (unsigned)nfunc);
    if (dtype == MGB_F64)
        k_kb_vnlpsi<double><<<grid, 256, 0, st>>>(kb->npoints, kb->d_point, kb->d_ent0,
            kb->d_ent_ion, kb->d_ent_node, kb->d_node0, kb->d_row0, kb->d_val0,
            (const double*)kb->d_proj, kb->d_coeff, kbpsi, (double*)out, (long long)ldo, nfunc, add);
    else
        k_kb_vnlpsi<float><<<grid, 256, 0, st>>>(kb->npoints, kb->d_point, kb->d_ent0,
            kb->d_ent_ion, kb->d_ent_node, kb->d_node0, kb->d_row0, kb->d_val0,
            (const float*)kb->d_proj, kb->d_coeff, kbpsi, (float*)out, (long long)ldo, nfunc, add);
    MGB_LAUNCHED("k_kb_vnlpsi");
    return MGB_OK;
}

} // extern "C"
