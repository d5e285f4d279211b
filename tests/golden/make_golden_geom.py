"""Golden table of pb::PEenv::geom (src/pb/PEenv.cc:335-611) from the compiled reference
(oracle/ref_shim_geom.cc): for a sweep of meshes and task counts the px x py x pz it picks
and the number of tasks it places (0: mesh refused; < ntasks: the reference would shrink the
communicator).  Run here (needs /root/reference built into oracle/_ref):

    python tests/golden/make_golden_geom.py
"""
import ctypes
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import REF_SO  # noqa: E402

NX = (8, 16, 24, 32, 40, 48, 64, 96, 128, 256, 30, 36, 60, 72, 100, 120)
NY = (8, 16, 32, 48, 64, 128, 20, 36)
NZ = (8, 16, 32, 64, 128, 256, 12, 28)
NT = (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16, 24, 32, 64)

lib = ctypes.CDLL(REF_SO)
out = (ctypes.c_int * 3)()
rows = []
for nx, ny, nz, nt in itertools.product(NX, NY, NZ, NT):
    placed = lib.ref_geom(nx, ny, nz, nt, 1, out)
    rows.append((nx, ny, nz, nt, placed, out[0], out[1], out[2]))
path = os.path.join(ROOT, "tests", "golden", "reference_geom.npz")
np.savez_compressed(path, table=np.asarray(rows, np.int16))
print("wrote", path, os.path.getsize(path), "bytes,", len(rows), "rows")
