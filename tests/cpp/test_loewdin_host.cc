// Host-only check of mgmol_b200::loewdinTransform (include/mgmol_b200.hpp): the
// n x n transform of GramMatrix::computeLoewdinTransform
// (src/GramMatrix.cc:267-299).  stdin: n, then S column-major; stdout: P, then
// sqrtS, one row per line.  No device needed.
#include <cstdio>
#include <vector>

#include "mgmol_b200.hpp"

int main()
{
    int n;
    if (std::scanf("%d", &n) != 1 || n < 1) return 2;
    std::vector<double> S((size_t)n * n), P(S.size()), R(S.size());
    for (size_t i = 0; i < S.size(); i++)
        if (std::scanf("%lf", &S[i]) != 1) return 2;
    if (!mgmol_b200::loewdinTransform(n, S.data(), P.data(), R.data()))
    {
        std::printf("not positive definite\n");
        return 3;
    }
    for (int pass = 0; pass < 2; pass++)
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++)
                std::printf("%.17g%c", (pass ? R : P)[i + (size_t)j * n], j + 1 < n ? ' ' : '\n');
    return 0;
}
