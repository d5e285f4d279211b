"""Golden vectors of the non-local Kleinman-Bylander projectors from
the reference's OWN code: KBprojectorSparse.cc / Species.cc / Mesh.cc / radial/*.cc compiled
unmodified for ORBDTYPE double and float (oracle/Makefile, oracle/ref_shim_kb.cc) on real
pseudopotential files of /root/reference/potentials.  Per species and type: every ion's
sparse projector as KBprojectorSparse::setup builds it (node list, value arrays, kbcoeff *
sign), kbpsi = vel <beta|psi> as computeLocalElement forms it, and H phi += V_nl phi as
get_vnlpsi + MPaxpy leave it (function 0).  Run here (needs /root/reference):

    python tests/golden/make_golden_kb.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from kb_cases import CENTERS, DIMS, LAP, LL, SPECIES, fields, key  # noqa: E402
from oracle.oracle import RefKB  # noqa: E402

out = {}
for tag, pseudo, flag in SPECIES:
    for dt in (np.float64, np.float32):
        ref = RefKB(dt)
        info = ref.setup(DIMS, LL, LAP, pseudo, flag)
        ions = [ref.add_ion(c) for c in CENTERS]
        psi, h0 = fields(dt)
        kbpsi = ref.kb_psi(psi)
        hphi = ref.kb_vnlpsi(kbpsi[:, :1], h0[:1], True)
        out[key(tag, dt, "info")] = np.array([info["nproj"], info["dim_nl"], int(ions[0]["single"])])
        for j, ion in enumerate(ions):
            out[key(tag, dt, "nlindex%d" % j)] = ion["nlindex"]
            out[key(tag, dt, "proj%d" % j)] = ion["proj"]
            out[key(tag, dt, "coeff%d" % j)] = ion["coeff"]
        out[key(tag, dt, "kbpsi")] = kbpsi
        out[key(tag, dt, "hphi0")] = hphi[0]
path = os.path.join(ROOT, "tests", "golden", "reference_kb.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")
