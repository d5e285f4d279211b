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
    }
    std::printf(fails ? "FAILED (%d)\n" : "ok\n", fails);
    return fails ? 1 : 0;
}
