// Localization masks of LocGridOrbitals as the kernels see them.
#pragma once
#include "common.cuh"

namespace mgb
{

// One multigrid level of a mask set.  off[color * subdivx + iloc]:
//   -2 zero the slab, -1 keep, >= 0 index of the slab of values in the pool
// (slab = sub0 * ny * nz values, no ghosts).  off == nullptr: no masks.
struct MaskView
{
    const int* off;
    const float* pool;    // float copy (the multigrid blocks are float)
    const double* pool64; // lmasktype double values, or nullptr
    int sub0, subdivx, op;
    long long slab;
};

inline MaskView no_mask()
{
    MaskView m;
    m.off    = nullptr;
    m.pool   = nullptr;
    m.pool64 = nullptr;
    m.sub0 = m.subdivx = 1;
    m.op   = 0;
    m.slab = 0;
    return m;
}

// apply one mask value to one element: GridMaskMult (u *= (T)m,
// src/GridMaskMult.cc:79) or GridMaskMax (GridMask::limitAbsValue,
// src/GridMask.h:186-192)
template <typename T, typename M>
__device__ __forceinline__ T mask_apply(int op, T u, M m)
{
    if (op == 0) return u * (T)m;
    if (u > m) return (T)m;
    if (u < -m) return -(T)m;
    return u;
}

} // namespace mgb

struct mgb_masks;
namespace mgb
{
// view of `level`; error (off == nullptr and *rc != 0) if the set is not
// committed or the level does not exist
MaskView mask_view(const mgb_masks* m, int level, int* rc);
int masks_ncolors(const mgb_masks* m);
int masks_nlevels(const mgb_masks* m);
bool masks_match(const mgb_masks* m, const mgb_grid* level0);
} // namespace mgb
