"""Multi-GPU parity (NCCL over NVLink), one rank per GPU: runs
tests/mgpu_worker.py under torchrun when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_parity(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multi-gpu parity ok" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_poisson_parity(world):
    """Decomposed PoissonMG / PoissonPCG / Hartree against the golden solutions of the
    compiled reference (tests/mgpu_poisson_worker.py)."""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29520 + world),
           os.path.join(ROOT, "tests", "mgpu_poisson_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multi-gpu poisson parity ok" in r.stdout
