"""Row f3 on the CPU: the plain-C restatement of the Kleinman-Bylander projector arithmetic
(oracle/mgmol_oracle_kb.inc) pinned against the reference's OWN code -- KBprojectorSparse.cc /
Species.cc / Mesh.cc / radial/*.cc compiled unmodified for both ORBDTYPEs (oracle/
ref_shim_kb.cc), real pseudopotential files: live where /root/reference exists, and through
the golden vectors generated from it (tests/golden/reference_kb.npz) everywhere -- and
against a dense numpy formulation of the same operator, V_nl = sum_rows |beta_row> coeff_row
vel <beta_row|."""
import os

import numpy as np
import pytest

from kb_cases import CENTERS, DIMS, LAP, LL, SPECIES, fields, key, vel
from oracle.oracle import RefKB, synthetic_kb_projectors, synthetic_orbitals

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gkb():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_kb.npz")))


def _golden_ions(gkb, tag, dt):
    return [{"nlindex": gkb[key(tag, dt, "nlindex%d" % j)], "proj": gkb[key(tag, dt, "proj%d" % j)],
             "coeff": gkb[key(tag, dt, "coeff%d" % j)]} for j in range(len(CENTERS))]


@pytest.mark.parametrize("tag", [s[0] for s in SPECIES])
@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_port_kb_against_reference_golden(port, gkb, tag, dt):
    """The restatement on the reference's own projectors (node lists and values as
    KBprojectorSparse::setup built them) against the reference's results: float bit for bit
    (MPdot and MPaxpy are plain loops with double accumulation and (T) casts there); double:
    the scatter bit for bit, the projections within the reordering of DDOT's sum."""
    ions = _golden_ions(gkb, tag, dt)
    nproj, dim_nl, single = gkb[key(tag, dt, "info")]
    assert all(len(i["coeff"]) == nproj and len(i["nlindex"]) == dim_nl ** 3 for i in ions)
    # overlapping ions and a ball that wraps around the box
    cnt = np.zeros(int(np.prod(DIMS)), int)
    for ion in ions:
        assert len(np.unique(ion["nlindex"])) == len(ion["nlindex"])
        cnt[ion["nlindex"]] += 1
    assert cnt.max() >= 2
    psi, h0 = fields(dt)
    kref = gkb[key(tag, dt, "kbpsi")]
    kb = port.kb_psi(ions, psi, vel())
    if dt == np.float32:
        assert np.array_equal(kb, kref)
    else:
        assert np.abs(kb - kref).max() <= 1e-13 * np.abs(kref).max()
    h = port.kb_vnlpsi(ions, kref[:, :1], h0[:1], add=True)
    assert np.array_equal(h[0], gkb[key(tag, dt, "hphi0")])
    assert (single == 1) == (nproj == 1)      # axpySKet for pseudo.Si, axpyKet for ONCV


@pytest.mark.skipif(not RefKB.available(), reason="compiled reference / potentials not present")
@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_kb_golden_is_what_the_compiled_reference_returns(port, gkb, dt):
    tag, pseudo, flag = SPECIES[0]
    ref = RefKB(dt)
    ref.setup(DIMS, LL, LAP, pseudo, flag)
    ions = [ref.add_ion(c) for c in CENTERS]
    for j, ion in enumerate(ions):
        assert np.array_equal(ion["nlindex"], gkb[key(tag, dt, "nlindex%d" % j)])
        assert np.array_equal(ion["proj"], gkb[key(tag, dt, "proj%d" % j)])
        assert np.array_equal(ion["coeff"], gkb[key(tag, dt, "coeff%d" % j)])
    psi, h0 = fields(dt)
    kb = ref.kb_psi(psi)
    assert np.array_equal(kb, gkb[key(tag, dt, "kbpsi")])
    # every function, both flavours of the scatter, live against the restatement
    assert np.array_equal(ref.kb_vnlpsi(kb, h0, True), port.kb_vnlpsi(ions, kb, h0, add=True))
    z = np.zeros_like(h0)
    assert np.array_equal(ref.kb_vnlpsi(kb, z, False), port.kb_vnlpsi(ions, kb, z, add=False))


@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-13), (np.float32, 2e-6)])
def test_port_kb_against_dense_operator(port, dt, tol):
    dims, ll, N = (12, 10, 16), (3.0, 2.5, 4.0), 5
    vel = float(np.prod([l / n for l, n in zip(ll, dims)]))
    ions = synthetic_kb_projectors(dims, ll, 7, 0.9, dt)
    assert any(len(i["coeff"]) == 4 for i in ions) and any(len(i["coeff"]) == 1 for i in ions)
    npt = int(np.prod(dims))
    phi = synthetic_orbitals(N, dims, dt)
    # dense rows
    rows, coeff = [], []
    for ion in ions:
        for p in range(ion["proj"].shape[0]):
            r = np.zeros(npt)
            r[ion["nlindex"]] = ion["proj"][p].astype(np.float64)
            rows.append(r)
            coeff.append(ion["coeff"][p])
    Bm, cf = np.array(rows), np.array(coeff)
    f = phi.reshape(N, -1).astype(np.float64)
    kb_exact = vel * Bm @ f.T
    kb = port.kb_psi(ions, phi, vel)
    scale = vel * np.abs(Bm) @ np.abs(f).T
    assert (np.abs(kb - kb_exact) / scale).max() <= tol
    v_exact = ((kb * cf[:, None]).T @ Bm).reshape(phi.shape)
    v = port.kb_vnlpsi(ions, kb, np.zeros_like(phi), add=False)
    assert np.abs(v - v_exact).max() <= (1e-13 if dt == np.float64 else 3e-6) * np.abs(v_exact).max()
    # add: H phi += V_nl phi
    h0 = synthetic_orbitals(N, dims, dt, first=20)
    h = port.kb_vnlpsi(ions, kb, h0, add=True)
    assert np.abs(h - (h0.astype(np.float64) + v_exact)).max() <= \
        (1e-13 if dt == np.float64 else 3e-6) * max(np.abs(v_exact).max(), np.abs(h0).max())
    # overlapping balls exist (the order of the ions matters for the float roundings)
    cnt = np.zeros(npt, int)
    for ion in ions:
        cnt[ion["nlindex"]] += 1
    assert cnt.max() >= 2
