"""CPU tests of the C-ABI library itself: it loads, exports every symbol the
header declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from mgmol_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "mgmol_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mgb_[A-Za-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _header_symbols() == _lib.exported_symbols()


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in _header_symbols():
        assert hasattr(L, name), name


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib.lib()
    g = _lib.MgbGrid()
    for d in range(3):
        g.dim[d] = g.gdim[d] = 8
        g.h[d] = 0.1
        g.bc[d] = 1
        g.nproc[d] = 1
    g.ghosts = 1
    buf = (ctypes.c_double * 8)()
    rc = L.mgb_hpsi(0, 1, ctypes.byref(g), buf, 512, buf, buf, 512, 1, None, None, None)
    assert rc == -4  # MGB_ENODEVICE
    assert b"no CPU fallback" in L.mgb_last_error()
    p = ctypes.c_void_p()
    assert L.mgb_precond_create(ctypes.byref(p), 0, 2, ctypes.byref(g), 4) == -4
    # every compute entry point refuses the same way
    n = ctypes.c_size_t(512)
    assert L.mgb_syrk_t(1, 1, n, 1.0, buf, n, buf, 1, None) == -4
    assert L.mgb_gemm_tn(1, 1, 1, n, 1.0, buf, n, buf, n, 0.0, buf, 1, None) == -4
    assert L.mgb_gemm_nn(1, n, 1, 1, 1.0, buf, n, buf, 1, 0.0, buf, n, None) == -4
    assert L.mgb_apply_b(0, 1, ctypes.byref(g), buf, n, buf, n, 1, None, None) == -4
    assert L.mgb_residual(0, 1, ctypes.byref(g), buf, n, buf, n, buf, 1, buf, n, 1, None,
                          None) == -4
    assert L.mgb_dot_cols(1, n, 1, 1.0, buf, n, buf, n, buf, None) == -4
    assert L.mgb_rho_blas3(1, n, 1, buf, n, buf, 1, buf, n, buf, None) == -4
    assert L.mgb_masks_create(ctypes.byref(p), ctypes.byref(g), 1, 2, 4, 0) == -4
    assert L.mgb_hpsi_host(0, 1, ctypes.byref(g), buf, n, buf, buf, n, 1, 0) == -4
    assert L.mgb_poisson_solve(0, 0, 1, ctypes.byref(g), buf, buf, 2, 2, 10, 1e-16, 10, None) == -4


def test_product_never_imports_oracle():
    """The product path must not route through the oracle."""
    pkg = os.path.join(ROOT, "mgmol_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/", "").lower() or \
                    f in ("_lib.py",), (f, "mentions oracle")


def test_host_constants_without_gpu():
    """Pure host arithmetic of the boundary (Lap constants, gamma)."""
    L = _lib.lib()
    h = (ctypes.c_double * 3)(0.25, 0.275, 0.25625)
    out = (ctypes.c_double * 3)()
    from oracle.oracle import Port
    P = Port()
    for lt in (0, 1, 2, 3, 4, 10):
        assert L.mgb_lap_constants(lt, h, out) == 0
        ref = (ctypes.c_double * 3)()
        assert P.lib.orc_lap_constants(lt, h, ref) == 0
        assert tuple(out) == tuple(ref)
    assert L.mgb_lap_constants(7, h, out) != 0
    assert L.mgb_gamma(0.03, 2, 0.4, -0.3) == P.gamma(0.03, 2, 0.4, -0.3)


def test_streamk_plan_covers_every_iteration_once():
    """Host logic of the contraction scheduler (no device): for ragged shapes,
    slab batches and CTA counts the segments partition the (tile, k-iteration)
    space exactly, every CTA's cost differs from the mean by at most one
    iteration, and no CTA touches more tiles than its partial slots allow."""
    import ctypes
    import numpy as np
    from mgmol_b200._lib import lib, check
    L = lib()
    rng = np.random.default_rng(0)
    cases = [(1, 256, 256, 2097152, 1, 32, 148, 10), (0, 256, 256, 2097152, 1, 32, 148, 10),
             (1, 4096, 4096, 262144, 1, 32, 148, 10), (0, 130, 70, 9216, 1, 16, 37, 10),
             (1, 5, 5, 4096, 4, 32, 16, 12), (0, 300, 520, 32768, 3, 32, 148, 10)]
    for _ in range(12):
        syrk = int(rng.integers(0, 2))
        m = int(rng.integers(1, 700))
        n = m if syrk else int(rng.integers(1, 700))
        cases.append((syrk, m, n, int(rng.integers(64, 200000)), int(rng.integers(1, 5)),
                      int(rng.choice([16, 32])), int(rng.integers(1, 149)), int(rng.integers(8, 17))))
    for syrk, m, n, k, nb, kc, ncta, cd in cases:
        cap = 200000
        segs = np.zeros((cap, 4), np.int64)
        ns, nkt, nt, nd = ctypes.c_int(), ctypes.c_longlong(), ctypes.c_int(), ctypes.c_int()
        check(L.mgb_debug_tn_plan(syrk, m, n, k, nb, kc, ncta, cd,
                                  segs.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(ns),
                                  ctypes.byref(nkt), ctypes.byref(nt), ctypes.byref(nd)))
        assert ns.value <= cap
        s = segs[:ns.value]
        nkt, nt, nd = nkt.value, nt.value, nd.value
        assert nkt == -(-k // kc)
        tm = -(-m // 128)
        assert nt == (tm * (tm + 1) // 2 if syrk else tm * -(-n // 128)) * nb
        cover = np.zeros((nt, nkt), np.int32) if nt * nkt < 5_000_000 else None
        per_tile = {}
        for g, u, a, b in s:
            assert 0 <= u < nt and 0 <= a < b <= nkt
            per_tile.setdefault(int(u), []).append((int(a), int(b), int(g)))
            if cover is not None:
                cover[u, a:b] += 1
        if cover is not None:
            assert (cover == 1).all()
        else:
            for u in range(nt):
                iv = sorted(per_tile[u])
                assert iv[0][0] == 0 and iv[-1][1] == nkt
                assert all(iv[i][1] == iv[i + 1][0] for i in range(len(iv) - 1))
        # segments of a tile come from consecutive CTAs in ascending order (the
        # fix-up's summation order)
        for iv in per_tile.values():
            gs = [g for _, _, g in sorted(iv)]
            assert gs == sorted(gs)
        # balance: cost per CTA within one iteration of the mean
        cost = np.zeros(ncta)
        for g, u, a, b in s:
            cost[g] += (b - a) * (cd if u < nd else 16)
        tot = cost.sum()
        assert np.abs(cost - tot / ncta).max() <= 16 + 1e-9


def test_tn_plan_with_tile_aligned_cta_groups():
    """The tcgen05 kernels' plan: with G = NT * ng CTAs and tiles of equal cost, CTA g works on
    tile g / ng only, and the CTAs of one K range (same g % ng) hold exactly the same range of
    k-iterations -- what the L2 reuse of the operand boxes and the 2 x 2-tile multicast clusters
    (k_gemm_tn_umma<MC>) rely on."""
    import ctypes
    import numpy as np
    from mgmol_b200._lib import lib, check
    L = lib()
    for m, n, k, kc, ng in ((256, 256, 2097152, 64, 37), (512, 512, 2097152, 64, 9),
                            (384, 256, 100000, 64, 7), (1024, 1024, 884736, 64, 2),
                            (256, 512, 12345 * 64 + 17, 64, 18)):
        nt = -(-m // 128) * -(-n // 128)
        ncta = nt * ng
        segs = np.zeros((4 * ncta, 4), np.int64)
        ns, nkt, ntile, nd = ctypes.c_int(), ctypes.c_longlong(), ctypes.c_int(), ctypes.c_int()
        check(L.mgb_debug_tn_plan(0, m, n, k, 1, kc, ncta, 16,
                                  segs.ctypes.data_as(ctypes.c_void_p), len(segs), ctypes.byref(ns),
                                  ctypes.byref(nkt), ctypes.byref(ntile), ctypes.byref(nd)))
        assert ntile.value == nt and ns.value == ncta, "one segment per CTA"
        s = segs[:ns.value]
        rng = {}
        for g, u, a, b in s:
            assert u == g // ng
            rng.setdefault(int(g % ng), set()).add((int(a), int(b)))
        assert all(len(v) == 1 for v in rng.values()), "same k-iterations for every tile of a K range"
        ends = sorted(next(iter(v)) for v in rng.values())
        assert ends[0][0] == 0 and ends[-1][1] == nkt.value
        assert all(ends[i][1] == ends[i + 1][0] for i in range(len(ends) - 1))
