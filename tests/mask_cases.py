"""The localization-mask case shared by the golden generator and the tests:
a 16^3 box in 2 x-slabs, 4 colors, 5 localized orbitals (one color slot empty
in slab 0, one orbital with a radius > 100 = mask 1 everywhere), 2 multigrid
levels."""
import numpy as np

from oracle.oracle import MaskTables

DIMS = (16, 16, 16)
LL = (8.0, 8.0, 8.0)
LEVELS = 2
SUBDIVX = 2
GIDS = [0, 1, 2, 3, 4]
GID_TABLE = np.array([[0, 1, -1, 3], [2, 1, 4, 3]])
RADII = np.array([2.5, 3.0, 1.2, 200.0, 2.0])
GAMMA = 0.3


def centers():
    return np.random.default_rng(3).uniform(0, 8, (5, 3))


def tables_to_arrays(mt, prefix):
    st, vo, pool = mt.flat()
    return {prefix + "_state": st, prefix + "_voff": vo, prefix + "_pool": pool}


def tables_from_arrays(d, prefix, op):
    st, vo, pool = d[prefix + "_state"], d[prefix + "_voff"], d[prefix + "_pool"]
    L, subdivx, ncolors = st.shape
    mt = MaskTables(DIMS, L - 1, subdivx, ncolors, op)
    for l in range(L):
        n = int(np.prod(mt.slab_shape(l)))
        for iloc in range(subdivx):
            for c in range(ncolors):
                s = int(st[l, iloc, c])
                if s == 2:
                    o = int(vo[l, iloc, c])
                    mt.set(l, iloc, c, 2, pool[o:o + n])
                else:
                    mt.set(l, iloc, c, s)
    return mt
