// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// The reference's AndersonMix<T> (src/AndersonMix.cc, compiled UNMODIFIED with
// -DTESTING by oracle/Makefile, which instantiates it for the plain vector type
// Solution) behind extern "C" entry points, so that the host-side mirror of the
// mixer (mgmol_b200/host.py, include/mgmol_b200.hpp) can be pinned to it.
#include <iostream>
#include <sstream>
#include <vector>

#include "Solution.h"

#include "AndersonMix.h"

namespace
{
struct RefAnderson
{
    Solution x, work;
    AndersonMix<Solution>* mix;
    RefAnderson(int m, double beta, int n, const double* x0) : x(n), work(n)
    {
        for (int i = 0; i < n; i++)
            x.u_[i] = x0[i];
        mix = new AndersonMix<Solution>(m, beta, x);
    }
    ~RefAnderson() { delete mix; }
};
}

extern "C"
{
void* ref_anderson_create(int m, double beta, int n, const double* x0)
{
    return new RefAnderson(m, beta, n, x0);
}
void ref_anderson_destroy(void* h) { delete (RefAnderson*)h; }
void ref_anderson_restart(void* h) { ((RefAnderson*)h)->mix->restart(); }
// one AndersonMix::update with residual f (dotProduct weight invs); the trial
// solution after the update is returned in x_out, the mixed residual in f
void ref_anderson_update(void* h, double* f, int n, double invs, double* x_out)
{
    RefAnderson* a = (RefAnderson*)h;
    Solution fs(n);
    for (int i = 0; i < n; i++)
        fs.u_[i] = f[i];
    fs.invs_ = a->x.invs_ = a->work.invs_ = invs;
    std::ostringstream os;
    a->mix->update(fs, a->work, os, false);
    for (int i = 0; i < n; i++)
    {
        f[i]     = fs.u_[i];
        x_out[i] = a->x.u_[i];
    }
}
}
