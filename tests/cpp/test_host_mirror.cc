// C++ parity test of the host mirror (include/mgmol_b200.hpp) against the CPU
// oracle (oracle/libmgmol_oracle.so), written the way MGmol's driver uses the
// classes: getHpsiAndTheta (src/computeHij.cc:404-455) then
// OrbitalsPreconditioning::precond_mg (src/ABPG.cc:73-140).
//
//   g++ -std=c++11 -O1 -I include tests/cpp/test_host_mirror.cc \
//       -L mgmol_b200 -lmgmol_b200 -L oracle -lmgmol_oracle -o test_host_mirror
// Run by tests/test_cpp_host.py (GPU marker).  Exit code 0 = parity.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mgmol_b200.hpp"

extern "C"
{
int orc_hpsi_f64(int lap_type, const int dims[3], const double ll[3], const int bc[3],
    const double* phi, const double* vtot, double* hphi, int nfunc);
int orc_hpsi_f32(int lap_type, const int dims[3], const double ll[3], const int bc[3],
    const float* phi, const double* vtot, float* hphi, int nfunc);
int orc_precond_mg(int lap_type, int mg_levels, int dtype, const int dims[3],
    const double ll[3], const int bc[3], void* res, int nfunc, double gamma);
int orc_precond_mg_masked(int lap_type, int mg_levels, int dtype, const int dims[3],
    const double ll[3], const int bc[3], void* res, int nfunc, double gamma, int subdivx,
    int ncolors, int mask_op, const int* mstate, const long long* mvoff, const double* mvalues);
void orc_app_mask_noghost_f64(const int dims[3], int subdivx, int ncolors, int op,
    const int* state, const long long* voff, const double* values, double* u, size_t ld,
    int nfunc);
void orc_app_mask_noghost_f32(const int dims[3], int subdivx, int ncolors, int op,
    const int* state, const long long* voff, const double* values, float* u, size_t ld,
    int nfunc);
void orc_kb_psi_f64(int nions, const long long* node0, const int* row0, const int* nlindex,
    const double* proj, double vel, const double* psi, size_t ld, int nfunc, double* kbpsi);
void orc_kb_psi_f32(int nions, const long long* node0, const int* row0, const int* nlindex,
    const float* proj, double vel, const float* psi, size_t ld, int nfunc, double* kbpsi);
void orc_kb_vnlpsi_f64(int nions, const long long* node0, const int* row0, const int* nlindex,
    const double* proj, const double* coeff, const double* kbpsi, size_t npt, double* out,
    size_t ldo, int nfunc, int add);
void orc_kb_vnlpsi_f32(int nions, const long long* node0, const int* row0, const int* nlindex,
    const float* proj, const double* coeff, const double* kbpsi, size_t npt, float* out,
    size_t ldo, int nfunc, int add);
void orc_add_ghosts_f64(const int dims[3], int g, const double* noghost, double* ghosted, int nfunc);
void orc_add_ghosts_f32(const int dims[3], int g, const float* noghost, float* ghosted, int nfunc);
void orc_trade_boundaries_f64(const int dims[3], int g, const int bc[3], double* u, int nfunc);
void orc_trade_boundaries_f32(const int dims[3], int g, const int bc[3], float* u, int nfunc);
void orc_rhs_4th_Mehr1_f64(const int dims[3], int g, const double* v, double* rhs, int rhs_ghosts,
    int nfunc);
void orc_rhs_4th_Mehr1_f32(const int dims[3], int g, const float* v, float* rhs, int rhs_ghosts,
    int nfunc);
void orc_gemm_nn_f64(int m, int n, int k, double alpha, const double* a, int lda, const double* b,
    int ldb, double beta, double* c, int ldc);
void orc_gemm_nn_f32(int m, int n, int k, double alpha, const float* a, int lda, const double* b,
    int ldb, double beta, float* c, int ldc);
void orc_gemm_tn_f64(int m, int n, int k, double alpha, const double* a, int lda,
    const double* b, int ldb, double* c, int ldc);
void orc_gemm_tn_f32(int m, int n, int k, double alpha, const float* a, int lda,
    const float* b, int ldb, double* c, int ldc);
}

using namespace mgmol_b200;

static int oracle_hpsi(int lap, const int dims[3], const double ll[3], const int bc[3],
    const double* phi, const double* v, double* out, int n)
{
    return orc_hpsi_f64(lap, dims, ll, bc, phi, v, out, n);
}
static int oracle_hpsi(int lap, const int dims[3], const double ll[3], const int bc[3],
    const float* phi, const double* v, float* out, int n)
{
    return orc_hpsi_f32(lap, dims, ll, bc, phi, v, out, n);
}
static void oracle_gemm_tn(int m, int n, int k, double alpha, const double* a, const double* b,
    double* c)
{
    orc_gemm_tn_f64(m, n, k, alpha, a, k, b, k, c, m);
}
static void oracle_gemm_tn(int m, int n, int k, double alpha, const float* a, const float* b,
    double* c)
{
    orc_gemm_tn_f32(m, n, k, alpha, a, k, b, k, c, m);
}

// B psi (lap 0) or psi, then (B psi) theta with the oracle's MPgemmNN
static void oracle_bphi_theta(int lap, const int dims[3], const int bc[3], const double* phi,
    const double* theta, double* out, int n, size_t npt)
{
    std::vector<double> b(phi, phi + npt * n);
    if (lap == 0)
    {
        std::vector<double> g((size_t)(dims[0] + 2) * (dims[1] + 2) * (dims[2] + 2) * n);
        orc_add_ghosts_f64(dims, 1, phi, g.data(), n);
        orc_trade_boundaries_f64(dims, 1, bc, g.data(), n);
        orc_rhs_4th_Mehr1_f64(dims, 1, g.data(), b.data(), 0, n);
    }
    orc_gemm_nn_f64((int)npt, n, n, 1., b.data(), (int)npt, theta, n, 0., out, (int)npt);
}
static void oracle_bphi_theta(int lap, const int dims[3], const int bc[3], const float* phi,
    const double* theta, float* out, int n, size_t npt)
{
    std::vector<float> b(phi, phi + npt * n);
    if (lap == 0)
    {
        std::vector<float> g((size_t)(dims[0] + 2) * (dims[1] + 2) * (dims[2] + 2) * n);
        orc_add_ghosts_f32(dims, 1, phi, g.data(), n);
        orc_trade_boundaries_f32(dims, 1, bc, g.data(), n);
        orc_rhs_4th_Mehr1_f32(dims, 1, g.data(), b.data(), 0, n);
    }
    orc_gemm_nn_f32((int)npt, n, n, 1., b.data(), (int)npt, theta, n, 0., out, (int)npt);
}

// deterministic inputs (LCG; no <random> so the numbers never depend on libstdc++)
static double lcg(unsigned long long& s)
{
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double)(s >> 11) / 9007199254740992.0 - 0.5;
}

template <typename T>
static int run(const int lap_type, const double tol, const double mg_tol)
{
    const int dims[3]      = { 16, 24, 32 };
    const unsigned gdim[3] = { 16, 24, 32 };
    const double ll[3]     = { 4.0, 6.0, 8.0 };
    const int bc[3]        = { 1, 1, 1 };
    const int N            = 5;
    const size_t npt       = (size_t)dims[0] * dims[1] * dims[2];
    std::vector<T> phi(npt * N), hphi(npt * N), href(npt * N);
    std::vector<double> vtot(npt);
    unsigned long long seed = 1234 + lap_type;
    for (int j = 0; j < N; j++)
        for (int ix = 0; ix < dims[0]; ix++)
            for (int iy = 0; iy < dims[1]; iy++)
                for (int iz = 0; iz < dims[2]; iz++)
                    phi[j * npt + ((size_t)ix * dims[1] + iy) * dims[2] + iz]
                        = (T)(std::cos(2 * M_PI * (j + 1) * ix / dims[0])
                                  * std::sin(2 * M_PI * iz / dims[2] + 0.3 * iy)
                              + 0.1 * lcg(seed));
    for (size_t i = 0; i < npt; i++)
        vtot[i] = -0.75 + 0.5 * lcg(seed);

    Grid grid(gdim, ll, ghostsFor(lap_type), bc);
    ExtendedGridOrbitals<T> orbitals(grid, N);
    orbitals.setPsi(phi.data());
    Hamiltonian<T> hamiltonian;
    hamiltonian.setup(grid, lap_type);
    hamiltonian.potential().setVtot(vtot.data());

    int fails = 0;
    // --- H psi ---------------------------------------------------------------
    const ExtendedGridOrbitals<T>& hl = hamiltonian.applyLocal(orbitals);
    hl.getPsiHost(hphi.data());
    oracle_hpsi(lap_type, dims, ll, bc, phi.data(), vtot.data(), href.data(), N);
    double worst = 0.;
    for (int j = 0; j < N; j++)
    {
        double scale = 0., err = 0.;
        for (size_t i = 0; i < npt; i++)
        {
            scale = std::fmax(scale, std::fabs((double)href[j * npt + i]));
            err   = std::fmax(err, std::fabs((double)hphi[j * npt + i] - (double)href[j * npt + i]));
        }
        worst = std::fmax(worst, err / scale);
    }
    std::printf("lap %2d %s  H psi       rel err %.3e (tol %.0e)\n", lap_type,
        sizeof(T) == 8 ? "f64" : "f32", worst, tol);
    if (!(worst <= tol)) fails++;
    // the cache: same indices -> no recomputation, same object
    const ExtendedGridOrbitals<T>& hl2 = hamiltonian.applyLocal(orbitals);
    if (&hl2 != &hl) fails++;

    // --- Phi^T H Phi -----------------------------------------------------------
    DeviceMemory<double> hij_dev((size_t)N * N);
    std::vector<double> hij(N * N), hij_ref(N * N);
    hamiltonian.addHlocalij(orbitals, orbitals, hij_dev.data());
    hij_dev.copy_to_host(hij.data(), hij.size());
    oracle_gemm_tn(N, N, (int)npt, grid.vel(), phi.data(), href.data(), hij_ref.data());
    double hmax = 0., herr = 0.;
    for (int i = 0; i < N * N; i++)
    {
        hmax = std::fmax(hmax, std::fabs(hij_ref[i]));
        herr = std::fmax(herr, std::fabs(hij[i] - hij_ref[i]));
    }
    const double htol = sizeof(T) == 8 ? 1e-11 : 1e-4;
    std::printf("lap %2d %s  Phi^T H Phi rel err %.3e (tol %.0e)\n", lap_type,
        sizeof(T) == 8 ? "f64" : "f32", herr / hmax, htol);
    if (!(herr <= htol * hmax)) fails++;

    // --- residual (B psi) theta - H psi  (MGmol::computeResidualUsingHPhi) ---------
    {
        std::vector<double> theta(N * N);
        for (int i = 0; i < N; i++)
            for (int j = 0; j <= i; j++)
                theta[i * N + j] = theta[j * N + i] = 0.4 * lcg(seed);
        DeviceMemory<double> theta_dev((size_t)N * N);
        theta_dev.copy_to_dev(theta.data(), theta.size());
        ExtendedGridOrbitals<T> resid(grid, N);
        computeResidualUsingHPhi(*hamiltonian.lapOper(), orbitals, hl, theta_dev.data(), resid);
        std::vector<T> got(npt * N), rref(npt * N);
        resid.getPsiHost(got.data());
        oracle_bphi_theta(lap_type, dims, bc, phi.data(), theta.data(), rref.data(), N, npt);
        double rmax = 0., rerr = 0.;
        for (size_t i = 0; i < npt * N; i++)
        {
            const double r = (double)rref[i] - (double)href[i]; // res.axpy(-1., hphi)
            rmax           = std::fmax(rmax, std::fabs(r));
            rerr           = std::fmax(rerr, std::fabs((double)got[i] - r));
        }
        const double rtol = sizeof(T) == 8 ? 1e-12 : 3e-6;
        std::printf("lap %2d %s  residual    rel err %.3e (tol %.0e)\n", lap_type,
            sizeof(T) == 8 ? "f64" : "f32", rerr / rmax, rtol);
        if (!(rerr <= rtol * rmax)) fails++;
        // residual norm through the per-orbital dot products (dotProductDiagonal)
        const double nrm2 = resid.dotProduct(resid);
        double ref2 = 0.;
        for (size_t i = 0; i < npt * N; i++)
            ref2 += (double)got[i] * (double)got[i];
        ref2 *= grid.vel();
        if (!(std::fabs(nrm2 - ref2) <= 1e-10 * ref2)) fails++;
    }

    // --- preconditioned residual ------------------------------------------------
    ExtendedGridOrbitals<T> res(grid, N);
    res.setPsi(href.data());
    OrbitalsPreconditioning<T> precond;
    precond.setup(res, 2, (short)lap_type);
    precond.setGamma(0.3);
    precond.precond_mg(res);
    std::vector<T> pres(npt * N), pref(href);
    res.getPsiHost(pres.data());
    orc_precond_mg(lap_type, 2, dtype_of<T>::value, dims, ll, bc, pref.data(), N, 0.3);
    worst = 0.;
    for (int j = 0; j < N; j++)
    {
        double scale = 0., err = 0.;
        for (size_t i = 0; i < npt; i++)
        {
            scale = std::fmax(scale, std::fabs((double)pref[j * npt + i]));
            err   = std::fmax(err, std::fabs((double)pres[j * npt + i] - (double)pref[j * npt + i]));
        }
        worst = std::fmax(worst, err / scale);
    }
    std::printf("lap %2d %s  precond_mg  rel err %.3e (tol %.0e)\n", lap_type,
        sizeof(T) == 8 ? "f64" : "f32", worst, mg_tol);
    if (!(worst <= mg_tol)) fails++;

    // --- computeMatB (src/ExtendedGridOrbitals.cc:901-967): vel Phi^T (B Phi), and
    // addHlocal2matrix (src/Hamiltonian.cc:163-212): hij += vel Phi^T (H_loc Phi)
    {
        std::vector<T> bref(npt * N);
        std::vector<double> ident(N * N, 0.), mb(N * N), mb_ref(N * N);
        for (int i = 0; i < N; i++)
            ident[i + i * N] = 1.;
        oracle_bphi_theta(lap_type, dims, bc, phi.data(), ident.data(), bref.data(), N, npt);
        oracle_gemm_tn(N, N, (int)npt, grid.vel(), phi.data(), bref.data(), mb_ref.data());
        DeviceMemory<double> mb_dev((size_t)N * N), acc_dev((size_t)N * N), wk_dev((size_t)N * N);
        orbitals.computeMatB(orbitals, *hamiltonian.lapOper(), mb_dev.data(), nullptr, nullptr, 2);
        mb_dev.copy_to_host(mb.data(), mb.size());
        double bmax = 0., berr = 0.;
        for (int i = 0; i < N * N; i++)
        {
            bmax = std::fmax(bmax, std::fabs(mb_ref[i]));
            berr = std::fmax(berr, std::fabs(mb[i] - mb_ref[i]));
        }
        // hij = 2 * (Phi^T H Phi) after two accumulating calls
        acc_dev.set(0);
        hamiltonian.addHlocal2matrix(orbitals, orbitals, acc_dev.data(), wk_dev.data());
        hamiltonian.addHlocal2matrix(orbitals, orbitals, acc_dev.data(), wk_dev.data(), true);
        std::vector<double> acc(N * N);
        acc_dev.copy_to_host(acc.data(), acc.size());
        double aerr = 0.;
        for (int i = 0; i < N * N; i++)
            aerr = std::fmax(aerr, std::fabs(acc[i] - 2. * hij[i]));
        const double btol = sizeof(T) == 8 ? 1e-11 : 1e-4;
        std::printf("lap %2d %s  matB        rel err %.3e, addHlocal2matrix %.3e (tol %.0e)\n", lap_type,
            sizeof(T) == 8 ? "f64" : "f32", berr / bmax, aerr / hmax, btol);
        if (!(berr <= btol * bmax && aerr <= btol * hmax)) fails++;
    }

    // --- orthonormalizeLoewdin: afterwards the Gram matrix is the identity, and
    // the transform is the symmetric S^-1/2 of the oracle's Gram matrix
    {
        ExtendedGridOrbitals<T> orth(grid, N);
        orth.setPsi(phi.data());
        const int idx = orth.getIterativeIndex();
        std::vector<double> P(N * N), S(N * N), G(N * N);
        orth.orthonormalizeLoewdin(P.data());
        if (orth.getIterativeIndex() <= idx) fails++;
        oracle_gemm_tn(N, N, (int)npt, grid.vel(), phi.data(), phi.data(), S.data());
        // P S P = I
        double perr = 0.;
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++)
            {
                double v = 0.;
                for (int k = 0; k < N; k++)
                    for (int l = 0; l < N; l++)
                        v += P[i + k * N] * S[k + l * N] * P[l + j * N];
                perr = std::fmax(perr, std::fabs(v - (i == j ? 1. : 0.)));
            }
        DeviceMemory<double> g_dev((size_t)N * N);
        orth.computeGram(g_dev.data());
        g_dev.copy_to_host(G.data(), G.size());
        double gerr = 0.;
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++)
                gerr = std::fmax(gerr, std::fabs(G[i + j * N] - (i == j ? 1. : 0.)));
        const double ptol = sizeof(T) == 8 ? 1e-11 : 2e-5;
        std::printf("lap %2d %s  Loewdin     |PSP-I| %.3e |G-I| %.3e (tol %.0e)\n", lap_type,
            sizeof(T) == 8 ? "f64" : "f32", perr, gerr, ptol);
        if (!(perr <= ptol && gerr <= ptol)) fails++;
    }
    return fails;
}

static void oracle_app_mask(const int dims[3], int subdivx, int nc, int op, const int* st,
    const long long* vo, const double* val, double* u, size_t ld)
{
    orc_app_mask_noghost_f64(dims, subdivx, nc, op, st, vo, val, u, ld, nc);
}
static void oracle_app_mask(const int dims[3], int subdivx, int nc, int op, const int* st,
    const long long* vo, const double* val, float* u, size_t ld)
{
    orc_app_mask_noghost_f32(dims, subdivx, nc, op, st, vo, val, u, ld, nc);
}

// LocGridOrbitals the way MGmol uses it: colors with per-slab gids, masks from
// a centre and a radius (here: a ball profile per color and level), applyMask,
// per-slab overlap matrices, masked preconditioner.
template <typename T>
static int run_localized(const int lap_type, const int op, const double mg_tol)
{
    const int dims[3]      = { 16, 16, 32 };
    const unsigned gdim[3] = { 16, 16, 32 };
    const double ll[3]     = { 4.0, 4.0, 8.0 };
    const int bc[3]        = { 1, 1, 1 };
    const int N = 4, subdivx = 2, levels = 2, numst = 6;
    const size_t npt = (size_t)dims[0] * dims[1] * dims[2];
    const std::vector<std::vector<int>> gids = { { 0, 1, -1, 3 }, { 2, 1, 4, 5 } };
    unsigned long long seed = 99 + lap_type + 7 * op;
    std::vector<T> phi(npt * N);
    for (size_t i = 0; i < phi.size(); i++)
        phi[i] = (T)lcg(seed);

    Grid grid(gdim, ll, ghostsFor(lap_type), bc);
    Masks masks(grid, levels, subdivx, N, op);
    // oracle tables [level][iloc][color]
    std::vector<int> st((levels + 1) * subdivx * N);
    std::vector<long long> vo(st.size(), -1);
    std::vector<double> pool;
    for (int l = 0; l <= levels; l++)
    {
        const int d[3] = { dims[0] >> l, dims[1] >> l, dims[2] >> l };
        const int s0   = d[0] / subdivx;
        for (int iloc = 0; iloc < subdivx; iloc++)
            for (int c = 0; c < N; c++)
            {
                const size_t t = ((size_t)l * subdivx + iloc) * N + c;
                const int gid  = gids[iloc][c];
                if (gid == -1)
                    st[t] = 0;
                else if (gid == 3)
                    st[t] = 1;
                else
                {
                    st[t] = 2;
                    vo[t] = (long long)pool.size();
                    std::vector<double> v((size_t)s0 * d[1] * d[2]);
                    for (int ix = 0; ix < s0; ix++)
                        for (int iy = 0; iy < d[1]; iy++)
                            for (int iz = 0; iz < d[2]; iz++)
                            {
                                const double x = (ix + iloc * s0 + 0.5) / d[0] - 0.3 - 0.1 * gid;
                                const double y = (iy + 0.5) / d[1] - 0.5;
                                const double z = (iz + 0.5) / d[2] - 0.2 * gid;
                                const double r = std::sqrt(x * x + y * y + z * z) / 0.45;
                                v[((size_t)ix * d[1] + iy) * d[2] + iz]
                                    = r < 0.75 ? 1. : (r < 1. ? (float)(4. * (1. - r)) : 0.);
                            }
                    pool.insert(pool.end(), v.begin(), v.end());
                    masks.set(l, iloc, c, 2, v.data());
                    continue;
                }
                masks.set<double>(l, iloc, c, (short)st[t]);
            }
    }
    masks.commit();

    int fails = 0;
    LocGridOrbitals<T> orbitals(grid, numst, gids, &masks);
    orbitals.setPsi(phi.data());
    orbitals.applyMask();
    std::vector<T> got(npt * N), ref(phi);
    orbitals.getPsiHost(got.data());
    oracle_app_mask(dims, subdivx, N, op, st.data(), vo.data(), pool.data(), ref.data(), npt);
    for (size_t i = 0; i < got.size(); i++)
        if (got[i] != ref[i])
        {
            fails++;
            break;
        }
    std::printf("lap %2d %s op %d  applyMask   %s\n", lap_type, sizeof(T) == 8 ? "f64" : "f32", op,
        fails ? "DIFFERS" : "bit-identical");

    // per-slab overlap: ss[iloc] = vel Phi_iloc^T Phi_iloc
    DeviceMemory<double> ss_dev((size_t)subdivx * N * N);
    std::vector<double> ss(subdivx * N * N), ss_ref(N * N);
    orbitals.getLocalOverlap(ss_dev.data());
    ss_dev.copy_to_host(ss.data(), ss.size());
    const size_t lnp = npt / subdivx;
    double serr = 0., smax = 0.;
    for (int iloc = 0; iloc < subdivx; iloc++)
    {
        std::vector<T> slab(lnp * N);
        for (int c = 0; c < N; c++)
            for (size_t i = 0; i < lnp; i++)
                slab[c * lnp + i] = ref[c * npt + iloc * lnp + i];
        oracle_gemm_tn(N, N, (int)lnp, grid.vel(), slab.data(), slab.data(), ss_ref.data());
        for (int i = 0; i < N * N; i++)
        {
            smax = std::fmax(smax, std::fabs(ss_ref[i]));
            serr = std::fmax(serr, std::fabs(ss[iloc * N * N + i] - ss_ref[i]));
        }
    }
    const double stol = sizeof(T) == 8 ? 1e-12 : 3e-6; // float: 3xTF32 tensor tiles
    std::printf("lap %2d %s op %d  slab Gram   rel err %.3e (tol %.0e)\n", lap_type,
        sizeof(T) == 8 ? "f64" : "f32", op, serr / smax, stol);
    if (!(serr <= stol * smax)) fails++;

    // masked preconditioner
    OrbitalsPreconditioning<T> precond;
    precond.setup(orbitals, levels, (short)lap_type, &masks);
    precond.setGamma(0.3);
    precond.precond_mg(orbitals);
    orbitals.getPsiHost(got.data());
    orc_precond_mg_masked(lap_type, levels, dtype_of<T>::value, dims, ll, bc, ref.data(), N, 0.3,
        subdivx, N, op, st.data(), vo.data(), pool.data());
    double worst = 0.;
    for (int j = 0; j < N; j++)
    {
        double scale = 0., err = 0.;
        for (size_t i = 0; i < npt; i++)
        {
            scale = std::fmax(scale, std::fabs((double)ref[j * npt + i]));
            err   = std::fmax(err, std::fabs((double)got[j * npt + i] - (double)ref[j * npt + i]));
        }
        worst = std::fmax(worst, scale > 0 ? err / scale : err);
    }
    std::printf("lap %2d %s op %d  precond_mg  rel err %.3e (tol %.0e)\n", lap_type,
        sizeof(T) == 8 ? "f64" : "f32", op, worst, mg_tol);
    if (!(worst <= mg_tol)) fails++;
    return fails;
}


// KBProjectors (row f3): the sparse projectors of three synthetic ions -- balls of grid
// points around a centre, one or four projectors each, two of them overlapping -- through
// computeKBpsi and computeHnlPhiAndAdd2HPhi against the oracle's restatement of
// KBPsiMatrixSparse::computeKBpsi / get_vnlpsi (oracle/mgmol_oracle_kb.inc).
static void oracle_kb_psi(int nions, const long long* node0, const int* row0, const int* idx,
    const double* proj, double vel, const double* psi, size_t ld, int nf, double* kbpsi)
{
    orc_kb_psi_f64(nions, node0, row0, idx, proj, vel, psi, ld, nf, kbpsi);
}
static void oracle_kb_psi(int nions, const long long* node0, const int* row0, const int* idx,
    const float* proj, double vel, const float* psi, size_t ld, int nf, double* kbpsi)
{
    orc_kb_psi_f32(nions, node0, row0, idx, proj, vel, psi, ld, nf, kbpsi);
}
static void oracle_kb_vnlpsi(int nions, const long long* node0, const int* row0, const int* idx,
    const double* proj, const double* coeff, const double* kbpsi, size_t npt, double* out,
    size_t ldo, int nf, int add)
{
    orc_kb_vnlpsi_f64(nions, node0, row0, idx, proj, coeff, kbpsi, npt, out, ldo, nf, add);
}
static void oracle_kb_vnlpsi(int nions, const long long* node0, const int* row0, const int* idx,
    const float* proj, const double* coeff, const double* kbpsi, size_t npt, float* out,
    size_t ldo, int nf, int add)
{
    orc_kb_vnlpsi_f32(nions, node0, row0, idx, proj, coeff, kbpsi, npt, out, ldo, nf, add);
}

template <typename T>
static int run_kb()
{
    const int dims[3]      = { 16, 24, 32 };
    const unsigned gdim[3] = { 16, 24, 32 };
    const double ll[3]     = { 4.0, 6.0, 8.0 };
    const int N            = 5;
    const size_t npt       = (size_t)dims[0] * dims[1] * dims[2];
    unsigned long long seed = 4321;
    std::vector<T> phi(npt * N), h0(npt * N);
    for (size_t i = 0; i < phi.size(); i++)
    {
        phi[i] = (T)(lcg(seed) + 0.2);
        h0[i]  = (T)lcg(seed);
    }
    const int centre[3][3] = { { 4, 6, 8 }, { 6, 7, 10 }, { 15, 23, 31 } }; // the last wraps
    const int nproj[3]     = { 1, 4, 4 };
    std::vector<long long> node0(1, 0);
    std::vector<int> row0(1, 0), nlindex;
    std::vector<T> proj;
    std::vector<double> coeff;
    Grid grid(gdim, ll, 1);
    KBProjectors<T> kb(grid);
    int fails = 0;
    for (int j = 0; j < 3; j++)
    {
        std::vector<int> idx;
        for (int dx = -3; dx <= 3; dx++)
            for (int dy = -3; dy <= 3; dy++)
                for (int dz = -3; dz <= 3; dz++)
                    if (dx * dx + dy * dy + dz * dz <= 9)
                    {
                        const int ix = (centre[j][0] + dx + dims[0]) % dims[0];
                        const int iy = (centre[j][1] + dy + dims[1]) % dims[1];
                        const int iz = (centre[j][2] + dz + dims[2]) % dims[2];
                        idx.push_back((ix * dims[1] + iy) * dims[2] + iz);
                    }
        std::vector<T> pr((size_t)nproj[j] * idx.size());
        for (size_t i = 0; i < pr.size(); i++)
            pr[i] = (T)lcg(seed);
        std::vector<double> cf(nproj[j]);
        for (int p = 0; p < nproj[j]; p++)
            cf[p] = (p & 1 ? -1. : 1.) * (0.5 + 0.1 * p);
        const int row = kb.addIon(idx, pr.data(), cf);
        if (row != row0.back()) fails++;
        nlindex.insert(nlindex.end(), idx.begin(), idx.end());
        proj.insert(proj.end(), pr.begin(), pr.end());
        coeff.insert(coeff.end(), cf.begin(), cf.end());
        node0.push_back(node0.back() + (long long)idx.size());
        row0.push_back(row0.back() + nproj[j]);
    }
    kb.commit();
    const int nrows = kb.nrows();
    if (nrows != row0.back()) fails++;

    ExtendedGridOrbitals<T> orbitals(grid, N), hphi(grid, N);
    orbitals.setPsi(phi.data());
    hphi.setPsi(h0.data());
    DeviceMemory<double> kbpsi_dev((size_t)nrows * N);
    kb.computeKBpsi(orbitals, kbpsi_dev.data());
    std::vector<double> kbpsi((size_t)nrows * N), kref((size_t)nrows * N);
    kbpsi_dev.copy_to_host(kbpsi.data(), kbpsi.size());
    oracle_kb_psi(3, node0.data(), row0.data(), nlindex.data(), proj.data(), grid.vel(), phi.data(),
        npt, N, kref.data());
    double scale = 0., kerr = 0.;
    for (size_t i = 0; i < kref.size(); i++)
        scale = std::fmax(scale, std::fabs(kref[i]));
    for (size_t i = 0; i < kref.size(); i++)
        kerr = std::fmax(kerr, std::fabs(kbpsi[i] - kref[i]) / scale);
    const double ktol = sizeof(T) == 8 ? 1e-13 : 1e-6;
    // the scatter alone: fed with the oracle's projections; float bit for bit (the (T)
    // roundings of axpySKet / axpyKet are reproduced), double to 1e-15
    kbpsi_dev.copy_to_dev(kref.data(), kref.size());
    const int idx0 = hphi.getIterativeIndex();
    kb.computeHnlPhiAndAdd2HPhi(kbpsi_dev.data(), hphi);
    if (hphi.getIterativeIndex() <= idx0) fails++;
    std::vector<T> got(npt * N), exp(h0);
    hphi.getPsiHost(got.data());
    oracle_kb_vnlpsi(3, node0.data(), row0.data(), nlindex.data(), proj.data(), coeff.data(),
        kref.data(), npt, exp.data(), npt, N, 1);
    double verr = 0., vmax = 0.;
    size_t nbits = 0;
    for (size_t i = 0; i < got.size(); i++)
    {
        vmax = std::fmax(vmax, std::fabs((double)exp[i]));
        verr = std::fmax(verr, std::fabs((double)got[i] - (double)exp[i]));
        if (got[i] != exp[i]) nbits++;
    }
    std::printf("       %s  KB projectors  kbpsi %.3e (tol %.0e)  H phi += Vnl phi %.3e, %zu values "
                "differ\n",
        sizeof(T) == 8 ? "f64" : "f32", kerr, ktol, verr / vmax, nbits);
    if (!(kerr <= ktol)) fails++;
    if (sizeof(T) == 4 ? nbits != 0 : !(verr <= 1e-15 * vmax)) fails++;
    return fails;
}

int main()
{
    if (mgb_device_count() < 1)
    {
        std::fprintf(stderr, "no CUDA device\n");
        return 77;
    }
    int fails = 0;
    for (int lap : { 0, 2 })
    {
        fails += run<double>(lap, 1e-12, 5e-6);
        fails += run<float>(lap, 1e-5, 5e-6);
        for (int op : { MGB_MASK_MULT, MGB_MASK_MAX })
        {
            fails += run_localized<double>(lap, op, 5e-6);
            fails += run_localized<float>(lap, op, 5e-6);
        }
    }
    fails += run_kb<double>();
    fails += run_kb<float>();
    std::printf(fails ? "FAILED (%d)\n" : "ok\n", fails);
    return fails ? 1 : 0;
}
