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
