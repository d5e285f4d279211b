"""Row f3 on the CPU: the plain-C restatement of the Kleinman-Bylander projector arithmetic
(oracle/mgmol_oracle_kb.inc, parity UNPINNED: the reference classes do not compile apart
from the driver) against a dense numpy formulation of the same operator, V_nl = sum_rows
|beta_row> coeff_row vel <beta_row|."""
import numpy as np
import pytest

from oracle.oracle import synthetic_kb_projectors, synthetic_orbitals


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
