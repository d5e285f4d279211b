// TEST INFRASTRUCTURE ONLY.
//
// The vector type the reference's own AndersonMix test instantiates the mixer
// with (tests/Anderson/Solution.h needs Boost, which this image lacks): a plain
// array of doubles with the operations AndersonMix<T> calls on T -- assign,
// operator-=, dotProduct, axpy, scal (src/AndersonMix.cc:72-319).  Together
// with the two empty headers next to it, it lets src/AndersonMix.cc compile
// UNMODIFIED with -DTESTING (which instantiates AndersonMix<Solution>,
// src/AndersonMix.cc:321-326).
#ifndef MGB_ORACLE_SOLUTION_H
#define MGB_ORACLE_SOLUTION_H

#include "MGmol_blas1.h" // as tests/Anderson/Solution.h does (declares DCOPY for the mixer)

#include <cmath>
#include <string>
#include <vector>

class Solution
{
public:
    std::vector<double> u_;
    double invs_;

    Solution() : invs_(1.) {}
    explicit Solution(const int n) : u_(n, 0.), invs_(1.) {}
    Solution(const std::string& /*name*/, const Solution& y) : u_(y.u_), invs_(y.invs_) {}
    void assign(const Solution& y) { u_ = y.u_; }
    double dotProduct(const Solution& v)
    {
        double s = 0.;
        for (size_t i = 0; i < u_.size(); i++)
            s += u_[i] * v.u_[i];
        return invs_ * s;
    }
    Solution& operator-=(const Solution& y)
    {
        for (size_t i = 0; i < u_.size(); i++)
            u_[i] -= y.u_[i];
        return *this;
    }
    void axpy(const double alpha, const Solution& y)
    {
        for (size_t i = 0; i < u_.size(); i++)
            u_[i] += alpha * y.u_[i];
    }
    void scal(const double alpha)
    {
        for (size_t i = 0; i < u_.size(); i++)
            u_[i] *= alpha;
    }
};

#endif
