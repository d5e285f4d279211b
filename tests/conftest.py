import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import Port
    return Port()


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
    return dict(np.load(path))


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and \
        np.array_equal(a.view(np.uint8), b.view(np.uint8))


def rel_inf(a, b):
    """max |a-b| relative to the per-function max norm of the reference b."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    ax = tuple(range(1, b.ndim))
    scale = np.abs(b).max(axis=ax, keepdims=True) if b.ndim > 1 else np.abs(b).max()
    scale = np.where(scale > 0, scale, 1.0)
    return float((np.abs(a - b) / scale).max())


def rel_elem(a, b, floor=1e-3):
    """max elementwise |a-b| / |b| over the entries with |b| >= floor x the per-function
    max norm of the reference b -- north_star's "elementwise relative" figure on the
    entries that are not cancellation residue.  A max-norm error e (rel_inf) bounds it by
    e / floor."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    ax = tuple(range(1, b.ndim))
    scale = np.abs(b).max(axis=ax, keepdims=True) if b.ndim > 1 else np.abs(b).max()
    big = np.abs(b) >= floor * np.where(scale > 0, scale, 1.0)
    if not big.any():
        return 0.0
    return float((np.abs(a - b)[big] / np.abs(b)[big]).max())
