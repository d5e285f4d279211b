"""The bench.py contract that can be checked without a GPU: the reference arm
(the reference's own compiled Hamiltonian path on the host cores) prints one
JSON line with every key the driver reads, ranks other than 0 stay silent, and
the GPU arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args,
                          capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
              "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["metric"] == "hpsi_gridpt_orbital_updates_per_s" and d["unit"] == "updates/s"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["higher_is_better"] is True
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"]
    assert cb["value"] == d["value"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_are_silent():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _run(["--steps", "1", "--warmup", "1"])
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)


def test_layout_is_shared_by_both_arms_and_follows_the_decomposition_rule():
    """bench.layout: the `config` dict both arms print (identical by construction), the
    default workload = the largest single-GPU configuration (256^3 x 512 doubles per GPU),
    the rank count's factors placed on x and y (z, the contiguous direction, last)."""
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    want = {1: "1x1x1", 2: "2x1x1", 4: "2x2x1", 8: "4x2x1"}
    for world, dec in want.items():
        args = argparse.Namespace(workload=bench.DEFAULT_WORKLOAD, decomp="auto", strong=False,
                                  orbitals=0, lap=None, dtype="f64")
        L = bench.layout(args, world)
        cfg = L["config"]
        assert cfg["decomposition"] == dec
        assert cfg["global_grid"] == [256, 256, 256] and cfg["orbitals"] == 512 * world
        assert np_prod(cfg["grid_per_gpu"]) * world == 256 ** 3
        # fixed work per GPU: weak scaling
        assert np_prod(cfg["grid_per_gpu"]) * cfg["orbitals"] == 256 ** 3 * 512
        assert "workload" in cfg and "tolerance" in cfg and "l2" in cfg
        assert json.dumps(bench.layout(args, world)["config"]) == json.dumps(cfg)
    args = argparse.Namespace(workload=bench.DEFAULT_WORKLOAD, decomp="2x2x2", strong=False,
                              orbitals=0, lap=None, dtype="f64")
    assert bench.layout(args, 8)["config"]["decomposition"] == "2x2x2"
    args.decomp = "auto"
    args.workload = "h2o64"
    L = bench.layout(args, 2)
    assert L["config"]["grid_per_gpu"] == [128, 128, 128] and L["config"]["orbitals"] == 256
    assert L["lap_type"] == 2


def np_prod(v):
    out = 1
    for x in v:
        out *= int(x)
    return out
