#!/usr/bin/env python
"""Benchmark of the orbital-update hot path (BASELINE.json metric:
"H psi grid-pt*orbital updates/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload h2o64|synth256|sih4] [--dtype f64|f32] [--lap 0|2]

A step is one Hamiltonian::applyLocal over the whole orbital block (the fused
H psi kernel).  Default workload = BASELINE configs[1], examples/H2O_64:
128^3 grid, 256 orbitals, ORBDTYPE double, FDtype=4th (lap_type 2); at N > 1
GPUs every rank owns a 128^3 box of a domain split along x (weak scaling) and
exchanges x-halo planes over NCCL every step.  Prints ONE JSON line.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (local dims, lattice of the local box, orbitals, lap_type, description)
    "h2o64": ((128, 128, 128), 23.4884, 256, 2,
              "examples/H2O_64 H psi: 128^3 grid x 256 orbitals, FDtype=4th"),
    "synth256": ((256, 256, 256), 46.9768, 64, 0,
                 "synthetic sweep: 256^3 grid, Mehrstellen"),
    "sih4": ((40, 40, 40), 14.0, 4, 0, "examples/SiH4: 40^3 grid x 4 orbitals, Mehrstellen"),
}


def measured_traffic(kernel_prefix, workload, dtype, lap_type):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant
    kernel, from the committed ncu --set full capture (profiles/); None when no
    capture matches this configuration."""
    p = os.path.join(ROOT, "profiles", "hpsi_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    if d.get("workload") == "%s %s lap%d" % (workload, dtype, 4 if lap_type == 2 else 0) \
            and d.get("kernel", "").startswith(kernel_prefix):
        return d.get("traffic_bytes_per_launch")
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region: NVML in-process
    (a query every 5 ms), nvidia-smi as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReason* bits (nvml.h)
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40),
            ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []      # (sm_mhz, reason bitmask)
        self.max_mhz = None
        self.stop_evt = threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:  # noqa: BLE001
            self.nvml = None

    def _reasons(self):
        n = self.nvml
        for name in ("nvmlDeviceGetCurrentClocksEventReasons",
                     "nvmlDeviceGetCurrentClocksThrottleReasons"):
            fn = getattr(n, name, None)
            if fn is not None:
                return int(fn(self.h))
        return 0

    def run(self):
        while not self.stop_evt.is_set():
            try:
                if self.nvml is not None:
                    mhz = float(self.nvml.nvmlDeviceGetClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
                    self.samples.append((mhz, self._reasons()))
                    self.stop_evt.wait(0.005)
                    continue
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                     "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    f = [x.strip() for x in out.split(",")]
                    mask = 0
                    for i, (_, bit) in enumerate(self.BITS):
                        if len(f) > 2 + i and f[2 + i].lower().startswith("active"):
                            mask |= bit
                    self.samples.append((float(f[0]), mask))
                    if self.max_mhz is None and f[1].replace(".", "").isdigit():
                        self.max_mhz = float(f[1])
            except Exception:  # noqa: BLE001
                pass
            self.stop_evt.wait(0.2)

    def summary(self):
        self.stop_evt.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(s[0] for s in self.samples)
        mask = 0
        for s_ in self.samples:
            mask |= s_[1]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": [n for n, bit in self.BITS if mask & bit],
                "samples": len(self.samples),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------
# CPU reference arm: the reference's own compiled sources (oracle/_ref) or,
# if they were not built, our restatement.  P worker processes each own a
# 1/P sub-box of the sample (the reference's MPI decomposition, halo = local
# wrap) and run its serial H psi.
# ---------------------------------------------------------------------------
def _cpu_worker(args):
    kind, lap_type, dims, ll, nfunc, dt, reps = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle.oracle import Port, Ref, synthetic_potential
    impl = Ref() if kind == "reference" else Port()
    rng = np.random.default_rng(11)
    phi = rng.standard_normal((nfunc,) + tuple(dims)).astype(dt)
    v = synthetic_potential(dims)
    impl.hpsi(lap_type, phi[:1], v, ll)  # warm
    t0 = time.perf_counter()
    for _ in range(reps):
        impl.hpsi(lap_type, phi, v, ll)
    return time.perf_counter() - t0


def cpu_hpsi_rate(lap_type, dims, ll, dt, budget_s=12.0):
    """grid-pt*orbital updates/s of the CPU reference on all host cores."""
    import multiprocessing as mp
    from oracle.oracle import Ref
    kind = "reference" if Ref.available() else "port"
    cores = os.cpu_count() or 1
    # split the box along x over P workers like a px x 1 x 1 PEenv
    P = 1
    while P * 2 <= cores and dims[0] % (P * 2) == 0 and dims[0] // (P * 2) >= 4:
        P *= 2
    sub = (dims[0] // P, dims[1], dims[2])
    subll = (ll[0] / P, ll[1], ll[2])
    ctx = mp.get_context("spawn")
    with ctx.Pool(P) as pool:
        t1 = max(pool.map(_cpu_worker, [(kind, lap_type, sub, subll, 1, dt, 1)] * P))
        nfunc = int(max(1, min(64, budget_s / max(t1, 1e-4) / 2)))
        t = max(pool.map(_cpu_worker, [(kind, lap_type, sub, subll, nfunc, dt, 1)] * P))
    updates = float(np.prod(dims)) * nfunc
    sample = ("%s H psi (Hamiltonian::applyLocal sequence), %dx%dx%d box split over %d "
              "single-thread ranks, %d orbitals" % (kind, dims[0], dims[1], dims[2], P, nfunc))
    return updates / t, P, kind, sample


def _cpu_iter_worker(args):
    """One rank of the CPU reference doing the in-scope work of one orbital-update
    iteration on its sub-box: H psi, Phi^T (H Phi), multigrid-preconditioned
    residual, Gram, Phi M (the reference's own kernels / mputils + BLAS)."""
    kind, lap_type, dims, ll, nfunc, dt = args
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from oracle.oracle import Port, Ref, synthetic_potential
    impl = Ref() if kind == "reference" else Port()
    rng = np.random.default_rng(11)
    phi = rng.standard_normal((nfunc,) + tuple(dims)).astype(dt)
    v = synthetic_potential(dims)
    M = rng.standard_normal((nfunc, nfunc)) / np.sqrt(nfunc)
    t = {}
    t0 = time.perf_counter()
    hphi = impl.hpsi(lap_type, phi, v, ll)
    t["hpsi"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    impl.gemm_tn(phi, hphi, 1.0)
    t["phiT_H_phi"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    impl.precond_mg(lap_type, 2, hphi, ll, 0.3)
    t["precond_mg"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    (impl.syrk if hasattr(impl, "syrk") else (lambda a, al: impl.gemm_tn(a, a, al)))(phi, 1.0)
    t["gram"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    impl.gemm_nn(phi, M)
    t["phi_M"] = time.perf_counter() - t0
    return t


def cpu_iteration(lap_type, dims, ll, dt, norb):
    """Seconds the CPU reference needs for the same pieces on the same box with
    all host cores: P single-thread ranks, each on a 1/P x-slab with all
    orbitals (no communication counted -- favourable to the CPU)."""
    import multiprocessing as mp
    from oracle.oracle import Ref
    kind = "reference" if Ref.available() else "port"
    cores = os.cpu_count() or 1
    P = 1
    # x slabs of >= 8 planes, divisible by 4 (two multigrid levels)
    while (P * 2 <= cores and dims[0] % (P * 2) == 0 and (dims[0] // (P * 2)) % 4 == 0
           and dims[0] // (P * 2) >= 8):
        P *= 2
    try:
        import psutil
        avail = psutil.virtual_memory().available
        per = 12.0 * np.prod(dims) / P * norb * np.dtype(dt).itemsize
        while P > 1 and per * P > 0.5 * avail:
            P //= 2
    except Exception:
        pass
    sub = (dims[0] // P, dims[1], dims[2])
    subll = (ll[0] / P, ll[1], ll[2])
    ctx = mp.get_context("spawn")
    with ctx.Pool(P) as pool:
        res = pool.map(_cpu_iter_worker, [(kind, lap_type, sub, subll, norb, dt)] * P)
    pieces = {k: max(r[k] for r in res) for k in res[0]}
    return {"seconds": sum(pieces.values()), "pieces_s": pieces, "cores": P, "kind": kind,
            "sample": "%s kernels, %dx%dx%d box as %d single-thread ranks (x slabs), all %d "
                      "orbitals, one pass, no communication counted"
                      % (kind, dims[0], dims[1], dims[2], P, norb)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dims, cell, norb, lap_default, desc = WORKLOADS[args.workload]
    lap_type = lap_default if args.lap is None else args.lap
    dt = np.float64 if args.dtype == "f64" else np.float32
    ll = (cell,) * 3
    vals = []
    info = None
    for _ in range(args.warmup + args.steps):
        rate, P, kind, sample = cpu_hpsi_rate(lap_type, dims, ll, dt,
                                              budget_s=60.0 / max(1, args.steps + args.warmup))
        vals.append(rate)
        info = (P, kind, sample)
    vals = vals[args.warmup:]
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "hpsi_gridpt_orbital_updates_per_s", "value": value,
        "unit": "updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        # a step is a bounded sample: report the time one full step of the
        # workload takes at the sampled rate
        "ms_per_step": float(np.prod(dims)) * norb / value * 1e3,
        "ms_per_step_basis": "whole workload at the sampled rate",
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": desc, "lap_type": lap_type, "grid": list(dims),
                   "orbitals": norb},
        "cpu_baseline": {"value": value, "unit": "updates/s", "cores": info[0],
                         "kind": info[1], "sample": info[2]},
        "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def _time_cuda(torch, fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def measure_pieces(H, grid, phi, ham, lap_type, tdt, S, norb, npt, comm=None, hstep=None):
    """The other three pieces of the path on the same orbital block (outside the
    timed region of the headline metric): multigrid-preconditioned residual,
    Gram, projected Hamiltonian, orbital mixing; each with the roofline that
    bounds it.  FP64 tensor peak = cuBLAS DGEMM measured here (MEASURED_PEAKS
    has no FP64 entry)."""
    import torch
    hbm, _ = measured_peaks()
    out = {}
    # FP64 tensor (DMMA) peak: large square cuBLAS DGEMM, best of 5
    m = 4096
    x = torch.rand((m, m), device="cuda", dtype=torch.float64)
    y = torch.rand((m, m), device="cuda", dtype=torch.float64)
    best = min(_time_cuda(torch, lambda: torch.matmul(x, y), reps=3, warm=1) for _ in range(2))
    fp64_peak = 2.0 * m ** 3 / (best * 1e-3) / 1e12
    del x, y
    out["fp64_tensor_peak_tflops"] = {"value": fp64_peak,
                                      "how": "cuBLAS DGEMM 4096^3 via torch.matmul, measured in this run"}
    upd = float(npt) * norb
    hphi = hstep() if hstep else ham.applyLocal(phi)
    # multigrid-preconditioned residual (OrbitalsPreconditioning::precond_mg)
    res = H.Orbitals(grid, norb, tdt)
    res.psi().copy_(hphi.psi())
    pc = H.OrbitalsPreconditioning()
    pc.setup(res, 2, lap_type)
    if comm is not None:
        pc.set_comm(comm)
    pc.gamma_ = 0.3
    ms = _time_cuda(torch, lambda: pc.precond_mg(res))
    model = (74.0 + 2 * S) * upd  # streaming model of the V-cycle
    out["precond_mg"] = {"ms": ms, "updates_per_s": upd / (ms * 1e-3), "mg_levels": 2,
                         "mode": {1: "literal", 2: "fused"}.get(pc.last_mode()),
                         "roofline": {"bound": "hbm", "achieved": model / (ms * 1e-3) / 1e9,
                                      "peak": hbm, "unit": "GB/s",
                                      "frac": model / (ms * 1e-3) / 1e9 / hbm,
                                      "model_bytes_per_update": 74.0 + 2 * S}}
    fl = float(norb) * norb * npt
    ms = _time_cuda(torch, lambda: phi.computeGram(comm))
    out["gram"] = {"ms": ms, "roofline": {"bound": "tensor", "achieved": fl / (ms * 1e-3) / 1e12,
                                          "peak": fp64_peak, "unit": "TFLOP/s",
                                          "frac": fl / (ms * 1e-3) / 1e12 / fp64_peak,
                                          "flops": "N^2 K (syrk)"}}
    ms = _time_cuda(torch, lambda: phi.computeLocalProduct(hphi, comm))
    out["phiT_H_phi"] = {"ms": ms, "roofline": {"bound": "tensor", "achieved": 2 * fl / (ms * 1e-3) / 1e12,
                                                "peak": fp64_peak, "unit": "TFLOP/s",
                                                "frac": 2 * fl / (ms * 1e-3) / 1e12 / fp64_peak,
                                                "flops": "2 N^2 K"}}
    M = torch.rand((norb, norb), device="cuda", dtype=torch.float64) - 0.5
    prod = H.Orbitals(grid, norb, tdt)
    ms = _time_cuda(torch, lambda: phi.multiplyByMatrix(M, prod))
    out["phi_M"] = {"ms": ms, "roofline": {"bound": "tensor", "achieved": 2 * fl / (ms * 1e-3) / 1e12,
                                           "peak": fp64_peak, "unit": "TFLOP/s",
                                           "frac": 2 * fl / (ms * 1e-3) / 1e12 / fp64_peak,
                                           "flops": "2 N^2 K"}}

    if comm is None:
        # "next" rows: residual assembly res = (B Phi) theta - H Phi in one
        # contraction pass with a fused epilogue, and the density rho += sum_j (Phi X)_j phi_j
        ms = _time_cuda(torch, lambda: H.computeResidualUsingHPhi(ham.lapOper(), phi, hphi, M, prod))
        out["residual"] = {"ms": ms, "roofline": {"bound": "tensor",
                                                  "achieved": 2 * fl / (ms * 1e-3) / 1e12,
                                                  "peak": fp64_peak, "unit": "TFLOP/s",
                                                  "frac": 2 * fl / (ms * 1e-3) / 1e12 / fp64_peak,
                                                  "flops": "2 N^2 K"}}
        rho = torch.zeros(grid.shape(), dtype=torch.float64, device="cuda")
        ms = _time_cuda(torch, lambda: H.computeRhoUsingBlas3(phi, M, rho))
        out["density"] = {"ms": ms, "roofline": {"bound": "tensor",
                                                 "achieved": 2 * fl / (ms * 1e-3) / 1e12,
                                                 "peak": fp64_peak, "unit": "TFLOP/s",
                                                 "frac": 2 * fl / (ms * 1e-3) / 1e12 / fp64_peak,
                                                 "flops": "2 N^2 K (+ 2 N K)"}}
        del rho

    # one orbital-update iteration's worth of the in-scope path, back to back on
    # one stream the way an SCF step orders it: H psi (with the
    # halo), Phi^T H Phi (+ all-reduce), preconditioned residual, Gram
    # (+ all-reduce), Phi M
    def iteration():
        h = hstep() if hstep else ham.applyLocal(phi, True)
        phi.computeLocalProduct(h, comm)
        pc.precond_mg(res)
        phi.computeGram(comm)
        phi.multiplyByMatrix(M, prod)
    ms = _time_cuda(torch, iteration, reps=3, warm=1)
    out["orbital_update_iteration"] = {
        "ms": ms, "sequence": "H psi, Phi^T H Phi, precond_mg (2 levels), Gram, Phi M"}
    pc.close()
    del res
    if comm is None:
        out.update(measure_poisson(H, grid, lap_type))
    return out


def measure_poisson(H, grid, lap_type):
    """SURVEY 8f row f4, reported next to the path: the Hartree Poisson multigrid
    (SolverLap / Mgm / Vcycle, double) on the workload's grid, ten V(2,2) sweeps on
    one scalar field.  Host control flow over the C-ABI grid operations; the
    cycle is launch-bound at this size, which is what the number shows.  A failure
    here never affects the headline line."""
    import torch
    try:
        from mgmol_b200._lib import lib
        from mgmol_b200.poisson import PoissonMG
        lt = lap_type if lap_type in (0, 1, 2) else 0
        rho = torch.rand(grid.shape(), dtype=torch.float64, device="cuda")
        rho -= rho.mean()
        solver = PoissonMG(grid, lt)
        vh = torch.zeros_like(rho)

        def solve():
            vh.zero_()
            solver.solve(vh, rho)
        n0 = lib().mgb_launch_count()
        ms = _time_cuda(torch, solve, reps=2, warm=1)
        launches = (lib().mgb_launch_count() - n0) // 3
        sweeps = max(1, solver.getNbSweeps())
        return {"poisson_mg": {"ms": ms, "ms_per_sweep": ms / sweeps, "sweeps": sweeps,
                               "relative_residual": solver.getFinalRelativeResidual(),
                               "launches_per_solve": int(launches), "lap_type": lt,
                               "bound": "launch latency (one field, %d kernels)" % launches}}
    except Exception as e:  # noqa: BLE001
        return {"poisson_mg": {"unavailable": "%s: %s" % (type(e).__name__, e)}}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mgmol_b200 import host as H
    from mgmol_b200._lib import lib, check

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        # keep this rank's host buffers (and its copy threads) on the NUMA node of
        # its GPU: the pinned host blocks of the end-to-end leg are first-touched
        # by this process
        try:
            if os.environ.get("MGB_BENCH_NO_BIND"):
                raise RuntimeError("binding disabled")
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                if os.environ.get("MGB_BENCH_VERBOSE"):
                    sys.stderr.write("rank %d bound to %d cpus\n" % (rank, len(cpus)))
        except Exception:  # noqa: BLE001
            pass
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from mgmol_b200.parallel import Communicator
        comm = Communicator(rank, world)

    dims, cell, norb, lap_default, desc = WORKLOADS[args.workload]
    if args.orbitals:
        norb = args.orbitals
    lap_type = lap_default if args.lap is None else args.lap
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    S = 8 if args.dtype == "f64" else 4
    g = H.ghosts_for(lap_type)
    if args.strong:
        # strong scaling: the workload's own grid split along x over the ranks
        assert dims[0] % world == 0 and (dims[0] // world) % 4 == 0, "x planes per rank"
        gdims = dims
        grid = H.Grid(gdims, (cell, cell, cell), g, (1, 1, 1), (world, 1, 1), (rank, 0, 0))
        dims = (dims[0] // world, dims[1], dims[2])
    else:
        gdims = (dims[0] * world, dims[1], dims[2])
        grid = H.Grid(gdims, (cell * world, cell, cell), g, (1, 1, 1), (world, 1, 1),
                      (rank, 0, 0))
    npt = grid.size()

    # synthetic orbitals: plane wave along z with orbital-dependent wavevector
    # + 0.1 U(-1,1) noise; potential a noisy constant.  A handful of large
    # launches (chunks of <= 64 orbitals) so the ncu launch list stays short.
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    phi = H.Orbitals(grid, norb, tdt)
    x = torch.arange(dims[2], device="cuda", dtype=torch.float64) / dims[2]
    for j0 in range(0, norb, 64):
        j1 = min(norb, j0 + 64)
        k = (torch.arange(j0, j1, device="cuda") % 5 + 1).to(torch.float64)
        wave = torch.cos(2 * np.pi * k[:, None] * x[None, :])[:, None, None, :]
        noise = torch.rand((j1 - j0,) + dims, generator=gen, device="cuda", dtype=tdt)
        phi.psi()[j0:j1] = (noise * 0.2 - 0.1) + wave.to(tdt)
        del noise
    vtot = (torch.rand(dims, generator=gen, device="cuda", dtype=torch.float64) * 0.1 - 0.75)
    ham = H.Hamiltonian()
    ham.setup(grid, lap_type)
    ham.potential(H.Potentials(vtot))

    # N > 1: V's halo is exchanged once (it is fixed over the steps); the
    # neighbours' boundary planes of the orbitals are read in place over NVLink
    # (peer mapping of the orbital block) -- or, if the block cannot be mapped,
    # packed and exchanged with NCCL send/recv every step.
    xh_phi = xh_v = None
    halo_mode = None
    if world > 1:
        from mgmol_b200._lib import MgbError
        xh_v = torch.zeros((1, 2 * g) + dims[1:], dtype=torch.float64, device="cuda")
        comm.halo_exchange_x(grid, g, vtot[None], xh_v)
        halo_mode = "peer_nvlink"
        if os.environ.get("MGB_BENCH_HALO") == "nccl":
            halo_mode = "nccl_packed"
        else:
            try:
                comm.register(phi.psi())
                ham.applyLocal(phi, True, None, xh_v, comm)
            except MgbError as e:
                if rank == 0:
                    sys.stderr.write("bench: peer halo unavailable (%s)\n" % e)
                halo_mode = "nccl_packed"
        ok = torch.tensor([1 if halo_mode == "peer_nvlink" else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok) == 0:
            halo_mode = "nccl_packed"
            xh_phi = torch.zeros((norb, 2 * g) + dims[1:], dtype=tdt, device="cuda")

    def step():
        if world > 1 and halo_mode == "peer_nvlink":
            return ham.applyLocal(phi, True, None, xh_v, comm)
        if world > 1:
            comm.halo_exchange_x(grid, g, phi.psi(), xh_phi)
        return ham.applyLocal(phi, True, xh_phi, xh_v)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = lib().mgb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    launches = lib().mgb_launch_count() - n0
    ms = ev0.elapsed_time(ev1)
    path = lib().mgb_hpsi_last_path()

    # kernel-only duration (no halo exchange) for the roofline
    evs = []
    for _ in range(min(args.steps, 10)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if halo_mode == "peer_nvlink":
            ham.applyLocal(phi, True, None, xh_v, comm)
        else:
            ham.applyLocal(phi, True, xh_phi, xh_v)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    clocks = sampler.summary() if sampler else None

    # end-to-end through the reference-facing call with HOST buffers: pinned
    # host orbitals in, pinned host H psi out.  N = 1: one mgb_hpsi_host call per
    # step (H2D, kernel and D2H pipelined over orbital blocks inside the
    # library).  N > 1: copy in, halo exchange + kernel, copy out.
    h_phi = torch.empty(phi.psi().shape, dtype=tdt).pin_memory()
    h_phi.copy_(phi.psi())
    h_out = torch.empty_like(h_phi).pin_memory()
    h_v = torch.empty(vtot.shape, dtype=torch.float64).pin_memory()
    h_v.copy_(vtot)
    e2e_steps = max(2, min(args.steps, 5))

    e2e_mode = "pipelined host call (mgb_hpsi_host)"
    if world > 1:
        e2e_mode = "pipelined host call per rank, halos in place (mgb_hpsi_host_peer)"
        try:
            ham.lapOper().applyWithPotHostPeer(comm, h_phi, h_v, h_out)
            okp = 1
        except Exception as e:  # noqa: BLE001
            okp = 0
            if rank == 0:
                sys.stderr.write("bench: host peer pipeline unavailable (%s)\n" % e)
        okt = torch.tensor([okp], device="cuda")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if int(okt) == 0:
            e2e_mode = "copy in, halo + kernel, copy out"

    def e2e_step():
        if world == 1:
            ham.lapOper().applyWithPotHost(h_phi, h_v, h_out)
        elif e2e_mode.startswith("pipelined"):
            ham.lapOper().applyWithPotHostPeer(comm, h_phi, h_v, h_out)
        else:
            phi.psi().copy_(h_phi, non_blocking=True)
            out = step()
            h_out.copy_(out.psi(), non_blocking=True)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(e2e_steps):
        e2e_step()
    ev1.record()
    barrier()
    e2e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3)

    t = torch.tensor([ms, e2e_ms, kern_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, kern_ms = (float(v) for v in t.cpu())

    pieces = None
    if not args.no_pieces:
        pieces = measure_pieces(H, grid, phi, ham, lap_type, tdt, S, norb, npt, comm,
                                step if world > 1 else None)
        if world > 1:
            # max over ranks of every timing
            keys = ["precond_mg", "gram", "phiT_H_phi", "phi_M", "orbital_update_iteration"]
            tt = torch.tensor([pieces[k]["ms"] for k in keys], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            for k, v in zip(keys, tt.cpu().tolist()):
                pieces[k]["ms_max_over_ranks"] = v

    if rank == 0:
        updates_per_step = float(npt) * norb * world
        value = updates_per_step * args.steps / (ms * 1e-3)
        e2e = updates_per_step * e2e_steps / (e2e_ms * 1e-3)
        peak, peak_src = measured_peaks()
        achieved = 2.0 * S * npt * norb / (kern_ms * 1e-3) / 1e9
        cpu_dims = gdims if args.strong else dims
        cpu_rate, cores, kind, sample = cpu_hpsi_rate(
            lap_type, cpu_dims, (cell,) * 3, np.float64 if args.dtype == "f64" else np.float32)
        line = {
            "metric": "hpsi_gridpt_orbital_updates_per_s", "value": value, "unit": "updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": desc, "lap_type": lap_type, "grid_per_gpu": list(dims),
                       "orbitals": norb, "decomposition": "%dx1x1" % world,
                       "halo": halo_mode,
                       "hpsi_path": {1: "tma_fused", 2: "generic_fused", 3: "ghosted"}.get(path),
                       "l2": "inputs larger than L2 (%.1f GB per step)" %
                             (2.0 * S * npt * norb / 1e9)},
            "e2e": {"value": e2e, "unit": "updates/s",
                    "h2d_bytes_per_step": int(S * npt * norb) * world,
                    "d2h_bytes_per_step": int(S * npt * norb) * world,
                    "how": e2e_mode},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": measured_traffic("k_hpsi_tma" if path == 1 else "k_hpsi_generic",
                                                     args.workload, args.dtype, lap_type)
                         if world == 1 and not args.orbitals else None,
                         "kernel": "k_hpsi_tma" if path == 1 else "k_hpsi_generic",
                         "kernel_ms": kern_ms, "peak_source": peak_src,
                         "algorithmic_bytes_per_update": 2 * S},
            "cpu_baseline": {"value": cpu_rate, "unit": "updates/s", "cores": cores,
                             "kind": kind, "sample": sample},
        }
        if pieces:
            if world == 1 and not args.no_cpu_iteration:
                ci = cpu_iteration(lap_type, dims, (cell,) * 3,
                                   np.float64 if args.dtype == "f64" else np.float32, norb)
                it = pieces["orbital_update_iteration"]
                it["cpu_reference"] = ci
                it["speedup_vs_cpu_reference"] = ci["seconds"] * 1e3 / it["ms"]
            line["pieces"] = pieces
        print(json.dumps(line))
    if world > 1:
        comm.check()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--strong", action="store_true",
                    help="split the workload's own grid over the GPUs (default: weak scaling, "
                         "one full grid per GPU)")
    ap.add_argument("--no-cpu-iteration", action="store_true",
                    help="skip the CPU reference timing of the orbital-update iteration")
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="h2o64", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--lap", type=int, default=None)
    ap.add_argument("--orbitals", type=int, default=0)
    ap.add_argument("--no-pieces", action="store_true",
                    help="skip the per-piece measurements (V-cycle, contractions)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
