"""Cases shared by the Kleinman-Bylander pin (tests/test_kb_cpu.py) and its golden-vector
generator: real pseudopotentials of the reference's potentials/ directory on a 24^3 box of
h = 0.6 bohr -- Si ONCV (two s + two p projector sets: 8 projectors per ion, axpyKet) and the
Si Troullier-Martins file (one s projector, p local: axpySKet) -- with two overlapping ions
and one whose projector ball wraps around the box."""
import numpy as np

DIMS, LL, LAP = (24, 24, 24), (14.4, 14.4, 14.4), 0
SPECIES = [("si_oncv", "pseudo.Si_ONCV_PBE_SG15", "n"), ("si_tm", "pseudo.Si", "n"),
           ("o_oncv_filtered", "pseudo.O_ONCV_PBE_SG15", "f")]
CENTERS = [(6.1, 5.7, 6.3), (7.3, 6.2, 6.0), (0.4, 13.9, 14.1)]
NFUNC = 3


def fields(dt):
    rng = np.random.default_rng(21)
    psi = rng.standard_normal((NFUNC,) + DIMS).astype(dt)
    h0 = rng.standard_normal((NFUNC,) + DIMS).astype(dt)
    return psi, h0


def vel():
    return float(np.prod([l / n for l, n in zip(LL, DIMS)]))


def key(tag, dt, what):
    return "%s_%s_%s" % (tag, "f64" if np.dtype(dt) == np.float64 else "f32", what)
