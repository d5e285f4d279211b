// The C++ AndersonMix mirror (include/mgmol_b200.hpp) on plain host vectors:
// reads a case on stdin (n m beta iters, then x0[n]) and prints the trial
// solution after every update, for the fixed-point iteration of
// tests/anderson_case.py.  tests/test_cpp_host.py compares the output with the
// trajectories of the reference's own AndersonMix<Solution>
// (tests/golden/reference_anderson.npz).  No GPU involved.
//
//   g++ -std=c++11 -O1 -I include tests/cpp/test_anderson_mirror.cc -L mgmol_b200 -lmgmol_b200
#include <cstdio>
#include <vector>

#include "mgmol_b200.hpp"

struct Vec
{
    std::vector<double> u;
    static double invs;
    explicit Vec(size_t n) : u(n, 0.) {}
    void assign(const Vec& y) { u = y.u; }
    Vec& operator-=(const Vec& y)
    {
        for (size_t i = 0; i < u.size(); i++)
            u[i] -= y.u[i];
        return *this;
    }
    double dotProduct(const Vec& y) const
    {
        double s = 0.;
        for (size_t i = 0; i < u.size(); i++)
            s += u[i] * y.u[i];
        return invs * s;
    }
    void axpy(const double a, const Vec& y)
    {
        for (size_t i = 0; i < u.size(); i++)
            u[i] += a * y.u[i];
    }
    void scal(const double a)
    {
        for (size_t i = 0; i < u.size(); i++)
            u[i] *= a;
    }
};
double Vec::invs = 1.;

int main()
{
    int n, m, iters;
    double beta;
    if (std::scanf("%d %d %lf %d", &n, &m, &beta, &iters) != 4) return 2;
    Vec x(n), work(n), r(n);
    for (int i = 0; i < n; i++)
        if (std::scanf("%lf", &x.u[i]) != 1) return 2;
    mgmol_b200::AndersonMix<Vec> mix(m, beta, x, [](const Vec& v) { return new Vec(v.u.size()); });
    for (int it = 0; it < iters; it++)
    {
        double xx = 0., xax = 0.;
        for (int i = 0; i < n; i++)
        {
            const double d = (double)i / (n - 1.0);
            xx += x.u[i] * x.u[i];
            xax += x.u[i] * d * x.u[i];
        }
        Vec::invs        = 1. / xx;
        const double lam = xax / xx;
        for (int i = 0; i < n; i++)
        {
            const double d = (double)i / (n - 1.0);
            r.u[i]         = (lam * x.u[i] - d * x.u[i]) / 1.0; // diag[n-1] = 1
        }
        mix.update(r, work);
        for (int i = 0; i < n; i++)
            std::printf("%.17g%c", x.u[i], i + 1 < n ? ' ' : '\n');
    }
    return 0;
}
