#!/usr/bin/env python
"""Benchmark of the orbital-update hot path (BASELINE.json metric:
"H psi grid-pt*orbital updates/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload h2o64|synth256|sih4] [--dtype f64|f32] [--lap 0|2]

A step is one Hamiltonian::applyLocal over the whole orbital block (the fused
H psi kernel).  Default workload = the largest single-GPU configuration of
BASELINE.json: configs[4], the synthetic sweep block, 256^3 grid x 512 orbitals
per GPU, ORBDTYPE double, Mehrstellen (`--workload h2o64` = configs[1],
examples/H2O_64: 128^3 x 256, FDtype=4th).  At N > 1 GPUs the same global 256^3
grid is split px x py x pz (`--decomp`, default: the factors PEenv::geom finds,
placed on x and y) with 512 N orbitals, every rank reads its neighbours' halo
layers in place over NVLink, and a small decomposed box is checked against the
oracle before anything is timed (`parity_mgpu`).  `--impl reference` times the
compiled reference on the host cores.  Prints ONE JSON line.
"""
import argparse
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
    # The N = 1 headline: BASELINE.json's metric is not quoted on one configuration, so the
    # largest configuration of `configs` that fits one GPU: configs[4], the synthetic sweep,
    # 256^3 grid x 512 orbitals, ORBDTYPE double.  At N GPUs the SAME global 256^3 grid is
    # split over the ranks the way PEenv::geom splits it and the orbital count grows with N
    # (512 N: the sweep's 512 / 1024 / 2048 / 4096 cells at 1 / 2 / 4 / 8 GPUs), so the work
    # per GPU is fixed (weak scaling).
    "synth256": {"grid": (256, 256, 256), "fixed": "global", "cell": 46.9768, "orbitals": 512,
                 "orbitals_scale": True, "lap": 0,
                 "desc": "synthetic sweep (configs[4]): 256^3 global grid, 512 orbitals per GPU"},
    # examples/H2O_64 (configs[1]): 128^3 x 256, FDtype=4th; at N GPUs every rank owns a
    # 128^3 box of a larger domain
    "h2o64": {"grid": (128, 128, 128), "fixed": "local", "cell": 23.4884, "orbitals": 256,
              "orbitals_scale": False, "lap": 2,
              "desc": "examples/H2O_64 H psi: 128^3 grid per GPU x 256 orbitals, FDtype=4th"},
    "sih4": {"grid": (40, 40, 40), "fixed": "local", "cell": 14.0, "orbitals": 4,
             "orbitals_scale": False, "lap": 0,
             "desc": "examples/SiH4: 40^3 grid x 4 orbitals, Mehrstellen"},
}
DEFAULT_WORKLOAD = "synth256"


def layout(args, world):
    """Decomposition and sizes of the workload at `world` GPUs -- shared by both arms so
    that their `config` dicts are identical."""
    from mgmol_b200.parallel import geom, geom_b200
    w = WORKLOADS[args.workload]
    if args.decomp and args.decomp != "auto":
        nproc = tuple(int(x) for x in args.decomp.lower().split("x"))
        assert len(nproc) == 3 and nproc[0] * nproc[1] * nproc[2] == world, "--decomp PxQxR"
        how = "requested"
    else:
        # PEenv::geom (src/pb/PEenv.cc:335-) on the grid the reference would be given
        nproc = (geom_b200(w["grid"][0], w["grid"][1], w["grid"][2], world) if world > 1
                 else (1, 1, 1))
        how = ("factors of the rank count over x and y (z, the contiguous direction, last); "
               "PEenv::geom would split %s" % "x".join(str(q) for q in (geom(*w["grid"], world) or ())))
        if nproc is None:
            nproc, how = (world, 1, 1), "x slabs (PEenv::geom refuses this mesh)"
    if args.strong or w["fixed"] == "global":
        gdims = w["grid"]
        cell = (w["cell"],) * 3
    else:
        gdims = tuple(n * p for n, p in zip(w["grid"], nproc))
        cell = tuple(w["cell"] * p for p in nproc)
    for n, p in zip(gdims, nproc):
        assert n % p == 0 and (n // p) % 4 == 0, "local dims must divide by 4 (two multigrid levels)"
    ldims = tuple(n // p for n, p in zip(gdims, nproc))
    norb = args.orbitals or w["orbitals"] * (world if (w["orbitals_scale"] and not args.strong) else 1)
    lap_type = w["lap"] if args.lap is None else args.lap
    S = 8 if args.dtype == "f64" else 4
    cfg = {"workload": w["desc"], "lap_type": lap_type, "global_grid": list(gdims),
           "grid_per_gpu": list(ldims), "orbitals": norb,
           "decomposition": "%dx%dx%d" % nproc, "decomposition_from": how,
           "l2": "inputs larger than L2 (%.1f GB per GPU per step)"
                 % (2.0 * S * np.prod(ldims) * norb / 1e9),
           "tolerance": "max |err| / per-orbital max norm <= 1e-12 (f64), 1e-5 (f32); the tests "
                        "(and parity_mgpu at N > 1) also report the elementwise relative error on "
                        "entries above 1e-3 of the orbital's max norm"}
    return {"nproc": nproc, "gdims": gdims, "ldims": ldims, "cell": cell, "norb": norb,
            "lap_type": lap_type, "S": S, "config": cfg}


def measured_traffic(kernel_sig, ldims, norb, dtype, lap_type):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the
    committed ncu --set full captures (profiles/hpsi_traffic.json, regenerated from the kept
    .ncu-rep files by tools/traffic_table.py), scaled by orbital count when the capture used
    fewer orbitals of the same box.  None when no capture matches the launched kernel."""
    p = os.path.join(ROOT, "profiles", "hpsi_traffic.json")
    if not os.path.exists(p):
        return None, "no capture"
    for e in json.load(open(p)).get("captures", []):
        if (e.get("grid") == list(ldims) and e.get("dtype") == dtype
                and e.get("lap_type") == lap_type and e.get("kernel") == kernel_sig):
            return (e["dram_bytes_per_launch"] * norb / e["orbitals"],
                    "ncu capture %s (%d orbitals, scaled by orbital count)" % (e["report"], e["orbitals"]))
    return None, "no capture of %s on this box shape" % kernel_sig


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


def _cpu_split(dims, cores, min_planes=4, mult=1):
    """x slabs over P single-thread ranks like a P x 1 x 1 PEenv."""
    P = 1
    while (P * 2 <= cores and dims[0] % (P * 2) == 0 and dims[0] // (P * 2) >= min_planes
           and (dims[0] // (P * 2)) % mult == 0):
        P *= 2
    return P


def cpu_hpsi_rate(lap_type, dims, ll, dt, budget_s=12.0):
    """grid-pt*orbital updates/s of the CPU reference on all host cores, on a bounded
    SAMPLE of the workload: the workload's own box, as many orbitals as the budget allows."""
    import multiprocessing as mp
    from oracle.oracle import Ref
    kind = "reference" if Ref.available() else "port"
    cores = os.cpu_count() or 1
    P = _cpu_split(dims, cores)
    sub = (dims[0] // P, dims[1], dims[2])
    subll = (ll[0] / P, ll[1], ll[2])
    ctx = mp.get_context("spawn")
    with ctx.Pool(P) as pool:
        t1 = max(pool.map(_cpu_worker, [(kind, lap_type, sub, subll, 1, dt, 1)] * P))
        nfunc = int(max(1, min(64, budget_s / max(t1, 1e-4) / 2)))
        t = max(pool.map(_cpu_worker, [(kind, lap_type, sub, subll, nfunc, dt, 1)] * P))
    updates = float(np.prod(dims)) * nfunc
    sample = ("SAMPLE: %s H psi (Hamiltonian::applyLocal sequence) on the %dx%dx%d box split over "
              "%d single-thread ranks (x slabs, halo = local wrap), %d orbitals of the workload's; "
              "rate extrapolates to the full orbital count"
              % (kind, dims[0], dims[1], dims[2], P, nfunc))
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
    """Seconds the CPU reference needs for the same pieces on the same box with all host
    cores: P single-thread ranks, each on a 1/P x-slab (no communication counted --
    favourable to the CPU).  Bounded: at most `ns` orbitals are run (memory and time); the
    grid-sized pieces scale with N / ns, the contractions with (N / ns)^2."""
    import multiprocessing as mp
    from oracle.oracle import Ref
    kind = "reference" if Ref.available() else "port"
    cores = os.cpu_count() or 1
    P = _cpu_split(dims, cores, min_planes=8, mult=4)  # two multigrid levels per slab
    ns = norb
    while ns > 16 and float(np.prod(dims)) * ns > 6e8:  # ~ 256 orbitals at 128^3
        ns //= 2
    try:
        import psutil
        avail = psutil.virtual_memory().available
        while ns > 8 and 12.0 * np.prod(dims) * ns * np.dtype(dt).itemsize > 0.4 * avail:
            ns //= 2
    except Exception:
        pass
    sub = (dims[0] // P, dims[1], dims[2])
    subll = (ll[0] / P, ll[1], ll[2])
    ctx = mp.get_context("spawn")
    with ctx.Pool(P) as pool:
        res = pool.map(_cpu_iter_worker, [(kind, lap_type, sub, subll, ns, dt)] * P)
    pieces = {k: max(r[k] for r in res) for k in res[0]}
    f1, f2 = norb / ns, (norb / ns) ** 2
    scaled = {"hpsi": pieces["hpsi"] * f1, "precond_mg": pieces["precond_mg"] * f1,
              "phiT_H_phi": pieces["phiT_H_phi"] * f2, "gram": pieces["gram"] * f2,
              "phi_M": pieces["phi_M"] * f2}
    return {"seconds": sum(scaled.values()), "pieces_s": scaled, "cores": P, "kind": kind,
            "sample": "SAMPLE: %s kernels, %dx%dx%d box as %d single-thread ranks (x slabs), %d of "
                      "the %d orbitals run once; grid-sized pieces scaled by N/ns, contractions by "
                      "(N/ns)^2; no communication counted"
                      % (kind, dims[0], dims[1], dims[2], P, ns, norb)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if rank != 0:
        return
    L = layout(args, max(world, args.gpus))
    dims, ll, norb, lap_type = L["gdims"], L["cell"], L["norb"], L["lap_type"]
    dt = np.float64 if args.dtype == "f64" else np.float32
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
        "ms_per_step_basis": "whole workload at the rate of the SAMPLE (see cpu_baseline.sample)",
        "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": L["config"],
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


def _tensor_roofline(fl, ms, peak, flops):
    return {"bound": "tensor", "achieved": fl / (ms * 1e-3) / 1e12, "peak": peak,
            "unit": "TFLOP/s", "frac": fl / (ms * 1e-3) / 1e12 / peak, "flops": flops}


def fp64_tensor_peak(torch):
    """FP64 tensor (DMMA) peak: large square cuBLAS DGEMM, best of 2 medians (MEASURED_PEAKS
    has no FP64 entry)."""
    m = 4096
    x = torch.rand((m, m), device="cuda", dtype=torch.float64)
    y = torch.rand((m, m), device="cuda", dtype=torch.float64)
    best = min(_time_cuda(torch, lambda: torch.matmul(x, y), reps=3, warm=1) for _ in range(2))
    del x, y
    return 2.0 * m ** 3 / (best * 1e-3) / 1e12


class PrecondChunks:
    """OrbitalsPreconditioning::precond_mg over a block in chunks of orbitals: the V-cycle's
    resident float work blocks (about 4.6 x npt x chunk floats) must fit next to the orbital
    blocks, so a 256^3 x 512 block is preconditioned 64 or 128 orbitals at a time."""

    def __init__(self, H, grid, tdt, norb, lap_type, comm, free_bytes):
        npt = grid.size()
        chunk = norb
        while chunk > 16 and chunk % 2 == 0 and 4.6 * 4 * npt * chunk > 0.7 * free_bytes:
            chunk //= 2
        self.chunk, self.norb, self.grid, self.H, self.tdt = chunk, norb, grid, H, tdt
        proto = H.Orbitals.__new__(H.Orbitals)  # setup() only reads the grid and the count
        proto.grid_, proto.numst_ = grid, chunk
        self.pc = H.OrbitalsPreconditioning()
        self.pc.setup(proto, 2, lap_type)
        if comm is not None:
            self.pc.set_comm(comm)
        self.pc.gamma_ = 0.3

    def __call__(self, res):
        for j0 in range(0, self.norb - self.chunk + 1, self.chunk):
            self.pc.precond_mg(self.H.Orbitals(self.grid, self.chunk, self.tdt,
                                               res.psi()[j0:j0 + self.chunk]))

    def last_mode(self):
        return self.pc.last_mode()

    def close(self):
        self.pc.close()


def measure_pieces(H, grid, phi, ham, lap_type, tdt, S, norb, npt, comm=None, hstep=None,
                   fp64_peak=None):
    """The other three pieces of the path on the same orbital block (outside the timed
    region of the headline metric): multigrid-preconditioned residual, Gram, projected
    Hamiltonian, orbital mixing; each with the roofline that bounds it.  Only two orbital
    blocks exist (phi and H phi: 2 x 68.7 GB at 256^3 x 512 doubles), so products are written
    over the H phi block and the V-cycle runs over it in chunks."""
    import torch
    hbm, _ = measured_peaks()
    out = {}
    out["fp64_tensor_peak_tflops"] = {"value": fp64_peak,
                                      "how": "cuBLAS DGEMM 4096^3 via torch.matmul, measured in this run"}
    upd = float(npt) * norb
    hphi = hstep() if hstep else ham.applyLocal(phi, True)
    free_b, _tot = torch.cuda.mem_get_info()
    free_b += torch.cuda.memory_reserved() - torch.cuda.memory_allocated()
    if comm is not None:
        # every rank must pick the same chunk (the V-cycle's neighbour barriers are
        # collective): size it for the rank with the least free memory
        import torch.distributed as dist
        fb = torch.tensor([free_b], dtype=torch.int64, device="cuda")
        dist.all_reduce(fb, op=dist.ReduceOp.MIN)
        free_b = int(fb)
    # multigrid-preconditioned residual (OrbitalsPreconditioning::precond_mg), on the H phi block
    pc = PrecondChunks(H, grid, tdt, norb, lap_type, comm, free_b)
    ms = _time_cuda(torch, lambda: pc(hphi), reps=3, warm=1)
    model = (74.0 + 2 * S) * upd  # streaming model of the V-cycle
    out["precond_mg"] = {"ms": ms, "updates_per_s": upd / (ms * 1e-3), "mg_levels": 2,
                         "orbitals_per_call": pc.chunk,
                         "mode": {1: "literal", 2: "fused"}.get(pc.last_mode()),
                         "roofline": {"bound": "hbm", "achieved": model / (ms * 1e-3) / 1e9,
                                      "peak": hbm, "unit": "GB/s",
                                      "frac": model / (ms * 1e-3) / 1e9 / hbm,
                                      "model_bytes_per_update": 74.0 + 2 * S}}
    hphi = hstep() if hstep else ham.applyLocal(phi, True)
    fl = float(norb) * norb * npt
    if S == 4:
        # float orbitals: 3xTF32 on tcgen05 -- useful flops against a third of the TF32 tensor
        # peak (three tensor products per pair)
        fp64_peak = tf32_tensor_peak(torch) / 3.0
        out["f32_tensor_peak_tflops"] = {"value": fp64_peak,
                                         "how": "cuBLAS SGEMM 8192^3 with TF32 allowed / 3 products, measured in this run"}
    ms = _time_cuda(torch, lambda: phi.computeGram(comm), reps=3, warm=1)
    out["gram"] = {"ms": ms, "roofline": _tensor_roofline(fl, ms, fp64_peak, "N^2 K (syrk)")}
    ms = _time_cuda(torch, lambda: phi.computeLocalProduct(hphi, comm), reps=3, warm=1)
    out["phiT_H_phi"] = {"ms": ms, "roofline": _tensor_roofline(2 * fl, ms, fp64_peak, "2 N^2 K")}
    M = torch.rand((norb, norb), device="cuda", dtype=torch.float64) - 0.5
    prod = hphi  # the product overwrites the H phi block
    ms = _time_cuda(torch, lambda: phi.multiplyByMatrix(M, prod), reps=3, warm=1)
    out["phi_M"] = {"ms": ms, "roofline": _tensor_roofline(2 * fl, ms, fp64_peak, "2 N^2 K")}
    # Mehrstellen: Phi^T B Phi (ExtendedGridOrbitals::computeMatB, called every SCF step,
    # src/DFTsolver.cc:382)
    if comm is None and lap_type in (0, 10) and hasattr(phi, "computeMatB"):
        ms = _time_cuda(torch, lambda: phi.computeMatB(ham.lapOper(), work=prod), reps=3, warm=1)
        out["matB"] = {"ms": ms, "roofline": _tensor_roofline(2 * fl, ms, fp64_peak, "2 N^2 K (+ B Phi pass)")}

    if comm is None:
        # "next" rows on a sub-block of <= 128 orbitals (a third block is needed):
        # residual assembly res = (B Phi) theta - H Phi in one contraction pass with a fused
        # epilogue, and the density rho += sum_j (Phi X)_j phi_j
        ns = min(norb, 128)
        if 2 * ns <= norb:
            sub = H.Orbitals(grid, ns, tdt, phi.psi()[:ns])
            hsub = H.Orbitals(grid, ns, tdt, hphi.psi()[:ns])
            rsub = H.Orbitals(grid, ns, tdt, hphi.psi()[ns:2 * ns])
        else:
            sub, hsub, rsub = phi, hphi, H.Orbitals(grid, ns, tdt)
        Ms = M[:ns, :ns].contiguous()
        fls = float(ns) * ns * npt
        ms = _time_cuda(torch, lambda: H.computeResidualUsingHPhi(ham.lapOper(), sub, hsub, Ms, rsub),
                        reps=3, warm=1)
        out["residual"] = {"ms": ms, "orbitals": ns,
                           "roofline": _tensor_roofline(2 * fls, ms, fp64_peak, "2 N^2 K")}
        rho = torch.zeros(grid.shape(), dtype=torch.float64, device="cuda")
        ms = _time_cuda(torch, lambda: H.computeRhoUsingBlas3(sub, Ms, rho), reps=3, warm=1)
        out["density"] = {"ms": ms, "orbitals": ns,
                          "roofline": _tensor_roofline(2 * fls, ms, fp64_peak, "2 N^2 K (+ 2 N K)")}
        del rho, sub, hsub, rsub

    if comm is None:
        out.update(measure_kb(H, grid, phi, hphi, tdt, S, norb))

    # one orbital-update iteration's worth of the in-scope path, back to back on
    # one stream the way an SCF step orders it: H psi (with the
    # halo), Phi^T H Phi (+ all-reduce), preconditioned residual, Gram
    # (+ all-reduce), Phi M
    def iteration():
        h = hstep() if hstep else ham.applyLocal(phi, True)
        phi.computeLocalProduct(h, comm)
        pc(h)
        phi.computeGram(comm)
        phi.multiplyByMatrix(M, h)
    ms = _time_cuda(torch, iteration, reps=3, warm=1)
    out["orbital_update_iteration"] = {
        "ms": ms, "sequence": "H psi, Phi^T H Phi, precond_mg (2 levels), Gram, Phi M"}
    pc.close()
    if comm is None:
        out.update(measure_poisson(H, grid, lap_type))
    return out


def measure_kb(H, grid, phi, hphi, tdt, S, norb):
    """the non-local Kleinman-Bylander projectors on the same block --
    <beta|phi> for every projector and orbital, then H phi += sum beta c -- on a synthetic
    ion set shaped like the reference's sparse projectors (balls of ~2 bohr radius, 1 or 4
    projectors per ion).  A failure here never affects the headline line."""
    import torch
    try:
        from mgmol_b200.synthetic import synthetic_kb_projectors
        npdt = np.float64 if tdt == torch.float64 else np.float32
        ll = tuple(grid.ll_)
        ions = synthetic_kb_projectors(grid.shape(), ll, 96, 2.0, npdt)
        kbp = H.KBProjectors(grid, tdt)
        for i in ions:
            kbp.add_ion(i["nlindex"], i["proj"], i["coeff"])
        kbp.commit()
        nodes = sum(len(i["nlindex"]) for i in ions)
        vals = sum(i["proj"].size for i in ions)
        kb = kbp.computeKBpsi(phi)
        ms1 = _time_cuda(torch, lambda: kbp.computeKBpsi(phi), reps=3, warm=1)
        ms2 = _time_cuda(torch, lambda: kbp.computeHnlPhiAndAdd2HPhi(kb, hphi), reps=3, warm=1)
        res = {"kb_nonlocal": {
            "ions": len(ions), "projectors": kbp.nrows(), "nodes": int(nodes),
            "kbpsi_ms": ms1, "vnlpsi_ms": ms2,
            "gathered_GBps": norb * (nodes * S + vals * S) / (ms1 * 1e-3) / 1e9,
            "scattered_GBps": norb * (2 * nodes * S + vals * S) / (ms2 * 1e-3) / 1e9,
            "bound": "gather / scatter over the ions' balls (sector-granular HBM access)"}}
        kbp.close()
        return res
    except Exception as e:  # noqa: BLE001
        return {"kb_nonlocal": {"unavailable": "%s: %s" % (type(e).__name__, e)}}


def measure_poisson(H, grid, lap_type):
    """SURVEY 8f row f4, reported next to the path: the Hartree Poisson multigrid
    (SolverLap / Mgm / Vcycle, double) on the workload's grid, ten V(2,2) sweeps on
    one scalar field.  A failure here never affects the headline line."""
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
        solve()
        n0 = lib().mgb_launch_count()
        solve()
        launches = lib().mgb_launch_count() - n0
        ms = _time_cuda(torch, solve, reps=2, warm=0)
        sweeps = max(1, solver.getNbSweeps())
        return {"poisson_mg": {"ms": ms, "ms_per_sweep": ms / sweeps, "sweeps": sweeps,
                               "relative_residual": solver.getFinalRelativeResidual(),
                               "launches_per_solve": int(launches), "lap_type": lt,
                               "graph_replays": int(getattr(solver, "graph_replays", 0))}}
    except Exception as e:  # noqa: BLE001
        return {"poisson_mg": {"unavailable": "%s: %s" % (type(e).__name__, e)}}


def measure_sweep(H, args, store_phi, store_out, fp64_peak):
    """BASELINE configs[4] at one GPU: H psi on the 256^3 grid for {f64, f32} x {Mehrstellen,
    4th order} (512 orbitals; 1024 for float, the same bytes), Gram on the same blocks.
    The blocks are views of the headline's two allocations."""
    import torch
    from mgmol_b200._lib import lib
    hbm, _ = measured_peaks()
    cells = []
    n = 256
    dims = (n, n, n)
    cell = WORKLOADS["synth256"]["cell"]
    vt = torch.rand(dims, device="cuda", dtype=torch.float64) * 0.1 - 0.75
    nbytes = store_phi.numel() * store_phi.element_size()
    for dname, tdt, S in (("f64", torch.float64, 8), ("f32", torch.float32, 4)):
        for norb in (512, 1024, 2048):
            need = n ** 3 * norb * S
            if need > nbytes:
                cells.append({"dtype": dname, "orbitals": norb, "infeasible":
                              "2 blocks of %.1f GB do not fit one 180 GB GPU" % (need / 1e9)})
                continue
            a = store_phi.reshape(-1).view(torch.uint8)[:need].view(tdt).view((norb,) + dims)
            b = store_out.reshape(-1).view(torch.uint8)[:need].view(tdt).view((norb,) + dims)
            if dname == "f32":
                a.uniform_(-0.5, 0.5)
            for lap in (0, 2):
                grid = H.Grid(dims, (cell,) * 3, H.ghosts_for(lap))
                op = H.LapFactory.createLap(grid, lap)
                ms = _time_cuda(torch, lambda: op.applyWithPot(a, vt, b), reps=5, warm=2)
                gbs = 2.0 * S * n ** 3 * norb / (ms * 1e-3) / 1e9
                sig = lib().mgb_hpsi_last_kernel().decode()
                tr, src = measured_traffic(sig, dims, norb, dname, lap)
                cells.append({"dtype": dname, "lap_type": lap, "orbitals": norb, "ms": ms,
                              "updates_per_s": float(n) ** 3 * norb / (ms * 1e-3),
                              "kernel": sig,
                              "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s",
                                           "frac": gbs / hbm, "traffic": tr, "traffic_source": src}})
            grid = H.Grid(dims, (cell,) * 3, 1)
            o = H.Orbitals(grid, norb, tdt, a)
            ms = _time_cuda(torch, lambda: o.computeGram(), reps=3, warm=1)
            if dname == "f64":
                rf = _tensor_roofline(float(norb) ** 2 * n ** 3, ms, fp64_peak, "N^2 K (syrk)")
            else:
                # 3xTF32 on tcgen05: three tensor products per pair (pieces.f32_contractions has
                # the measured TF32 peak and the issued-flop count)
                rf = {"bound": "tensor", "unit": "TFLOP/s", "flops": "N^2 K (syrk), useful",
                      "achieved": float(norb) ** 2 * n ** 3 / (ms * 1e-3) / 1e12}
            cells.append({"dtype": dname, "piece": "gram", "orbitals": norb, "ms": ms, "roofline": rf})
    return cells


def tf32_tensor_peak(torch):
    """TF32 tensor peak: large square cuBLAS SGEMM with TF32 allowed, best of 2 medians
    (MEASURED_PEAKS holds bf16 only)."""
    m = 8192
    x = torch.rand((m, m), device="cuda", dtype=torch.float32)
    y = torch.rand((m, m), device="cuda", dtype=torch.float32)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        best = min(_time_cuda(torch, lambda: torch.matmul(x, y), reps=3, warm=1) for _ in range(2))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    del x, y
    return 2.0 * m ** 3 / (best * 1e-3) / 1e12


def measure_f32_contractions(H, store_phi, store_out):
    """ORBDTYPE float contractions (north star: FP32 tensor-core tiles): Gram, Phi^T (H Phi) and
    Phi M on the H2O_64 block shape (128^3 x 256) and on the 256^3 x 512 sweep block, in views
    of the headline's two allocations.  Gram and Phi^T H Phi run on tcgen05 (kind::tf32, 3xTF32
    split, TMEM accumulators): `executed` counts the tensor flops actually issued (3 products
    per pair, 2 on diagonal Gram tiles) against the TF32 tensor peak cuBLAS reaches in this
    run; `useful` is the contraction itself; the operand stream against the HBM peak is the
    other bound.  Phi M (k_gemm_nn_umma: the Phi tile in TMEM, coefficient tiles by TMA) moves
    the block once in and once out.  A failure here never affects the headline line."""
    import torch
    from mgmol_b200._lib import lib, check
    hbm, _ = measured_peaks()
    try:
        peak = tf32_tensor_peak(torch)
        out = {"tf32_tensor_peak_tflops": {"value": peak,
                                           "how": "cuBLAS SGEMM 8192^3 with TF32 allowed, measured in this run"},
               "arithmetic": "3xTF32 split products, chunked FP32 sums folded into double "
                             "(<= 3e-6 of |a||b| against the exact contraction in the tests)",
               "cells": []}
        for n, norb in ((128, 256), (256, 512)):
            K = n ** 3
            need = K * norb * 4
            if need > store_phi.numel() * store_phi.element_size():
                continue
            a = store_phi.reshape(-1).view(torch.uint8)[:need].view(torch.float32).view(norb, K)
            b = store_out.reshape(-1).view(torch.uint8)[:need].view(torch.float32).view(norb, K)
            a.uniform_(-0.5, 0.5)
            b.uniform_(-0.5, 0.5)
            C = torch.empty((norb, norb), device="cuda", dtype=torch.float64)
            M = torch.rand((norb, norb), device="cuda", dtype=torch.float64) - 0.5
            fl = float(norb) * norb * K
            tm = (norb + 127) // 128
            tile = 2.0 * 128 * 128 * K

            def cell(piece, ms, useful, executed, bytes_, kernel):
                r = {"grid": [n, n, n], "orbitals": norb, "dtype": "f32", "piece": piece, "ms": ms,
                     "kernel": kernel, "useful_tflops": useful / (ms * 1e-3) / 1e12,
                     "operand_GBps": bytes_ / (ms * 1e-3) / 1e9,
                     "operand_frac_of_hbm": bytes_ / (ms * 1e-3) / 1e9 / hbm}
                if executed:
                    r["roofline"] = _tensor_roofline(executed, ms, peak,
                                                     "tensor flops issued (3xTF32), TF32 peak")
                return r
            ms = _time_cuda(torch, lambda: check(lib().mgb_syrk_t(
                0, norb, K, 1.0, a.data_ptr(), K, C.data_ptr(), norb, None)), reps=5, warm=2)
            out["cells"].append(cell("gram", ms, fl, tile * (3 * tm * (tm - 1) / 2 + 2 * tm),
                                     4.0 * K * norb, "k_gemm_tn_umma<SYRK>"))
            ms = _time_cuda(torch, lambda: check(lib().mgb_gemm_tn(
                0, norb, norb, K, 1.0, a.data_ptr(), K, b.data_ptr(), K, 0.0, C.data_ptr(), norb,
                None)), reps=5, warm=2)
            out["cells"].append(cell("phiT_H_phi", ms, 2 * fl, 3 * tile * tm * tm,
                                     8.0 * K * norb, "k_gemm_tn_umma"))
            ms = _time_cuda(torch, lambda: check(lib().mgb_gemm_nn(
                0, K, norb, norb, 1.0, a.data_ptr(), K, M.data_ptr(), norb, 0.0, b.data_ptr(), K,
                None)), reps=5, warm=2)
            out["cells"].append(cell("phi_M", ms, 2 * fl, 3 * tile * tm * tm, 8.0 * K * norb,
                                     "k_gemm_nn_umma"))
            # the whole orbital-update iteration with float orbitals (the reference's ORBDTYPE
            # float build): H psi, Phi^T H Phi, precond_mg, Gram, Phi M back to back
            lap = 0 if n == 256 else 2
            dims = (n, n, n)
            grid = H.Grid(dims, (WORKLOADS["synth256" if n == 256 else "h2o64"]["cell"],) * 3,
                          H.ghosts_for(lap))
            phi = H.Orbitals(grid, norb, torch.float32, a.view((norb,) + dims))
            work = H.Orbitals(grid, norb, torch.float32, b.view((norb,) + dims))
            op = H.LapFactory.createLap(grid, lap)
            vt = torch.rand(dims, device="cuda", dtype=torch.float64) * 0.1 - 0.75
            free_b, _tot = torch.cuda.mem_get_info()
            free_b += torch.cuda.memory_reserved() - torch.cuda.memory_allocated()
            pc = PrecondChunks(H, grid, torch.float32, norb, lap, None, free_b)

            def iteration():
                op.applyWithPot(phi.psi(), vt, work.psi())
                phi.computeLocalProduct(work)
                pc(work)
                phi.computeGram()
                phi.multiplyByMatrix(M, work)
            ms = _time_cuda(torch, iteration, reps=3, warm=1)
            pc.close()
            out["cells"].append({"grid": [n, n, n], "orbitals": norb, "dtype": "f32", "lap_type": lap,
                                 "piece": "orbital_update_iteration", "ms": ms,
                                 "sequence": "H psi, Phi^T H Phi, precond_mg (2 levels), Gram, Phi M"})
            del vt, phi, work
        return out
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def measure_iteration_e2e(H, store_phi, store_out, fp64_peak, with_cpu):
    """INTEGRATION level 2 end to end, host to host, on the H2O_64 block (128^3 x 256
    doubles; the 256^3 x 512 block would need 137 GB of pinned host memory): the orbitals
    go to the device ONCE, one orbital-update iteration's worth of the path runs resident
    (H psi, Phi^T H Phi, precond_mg on the residual block, Gram, Phi M), and the new orbitals
    plus the two N x N matrices come back ONCE.  Timed with the host clock around the whole
    thing, next to the compiled reference's own kernels on the host cores."""
    import torch
    n, norb, lap, tdt, S = 128, 256, 2, torch.float64, 8
    dims = (n, n, n)
    cell = WORKLOADS["h2o64"]["cell"]
    need = n ** 3 * norb * S
    grid = H.Grid(dims, (cell,) * 3, H.ghosts_for(lap))
    dphi = store_phi.reshape(-1).view(torch.uint8)[:need].view(tdt).view((norb,) + dims)
    dwork = store_out.reshape(-1).view(torch.uint8)[:need].view(tdt).view((norb,) + dims)
    dres = store_out.reshape(-1).view(torch.uint8)[need:2 * need].view(tdt).view((norb,) + dims)
    h_phi = torch.empty((norb,) + dims, dtype=tdt).pin_memory()
    h_phi.copy_(dphi)
    h_new = torch.empty_like(h_phi).pin_memory()
    h_v = (torch.rand(dims, dtype=torch.float64) * 0.1 - 0.75).pin_memory()
    h_mat = torch.empty((2, norb, norb), dtype=torch.float64).pin_memory()
    dv = torch.empty(dims, dtype=torch.float64, device="cuda")
    M = torch.rand((norb, norb), device="cuda", dtype=torch.float64) - 0.5
    phi = H.Orbitals(grid, norb, tdt, dphi)
    res = H.Orbitals(grid, norb, tdt, dres)
    out = H.Orbitals(grid, norb, tdt, dwork)
    ham = H.Hamiltonian()
    ham.setup(grid, lap)
    ham.potential(H.Potentials(dv))
    pc = H.OrbitalsPreconditioning()
    pc.setup(phi, 2, lap)
    pc.gamma_ = 0.3

    def once():
        dphi.copy_(h_phi, non_blocking=True)
        dv.copy_(h_v, non_blocking=True)
        h = ham.applyLocal(phi, True)
        hij = phi.computeLocalProduct(h)
        res.psi().copy_(h.psi())
        pc.precond_mg(res)
        gram = phi.computeGram()
        phi.multiplyByMatrix(M, out)
        h_new.copy_(out.psi(), non_blocking=True)
        h_mat[0].copy_(hij, non_blocking=True)
        h_mat[1].copy_(gram, non_blocking=True)
        torch.cuda.synchronize()
    once()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        once()
        ts.append(time.perf_counter() - t0)
    pc.close()
    sec = float(np.median(ts))
    r = {"seconds": sec, "block": "128^3 x 256 orbitals, f64, FDtype=4th (H2O_64)",
         "h2d_bytes": int(need + n ** 3 * 8), "d2h_bytes": int(need + 2 * norb * norb * 8),
         "host_GBps_both_ways": (2 * need + n ** 3 * 8) / 1e9 / sec,
         "sequence": "H2D Phi, V; H psi, Phi^T H Phi, precond_mg, Gram, Phi M; D2H Phi', H, S",
         "how": "host clock around the whole sequence, median of 3"}
    if with_cpu:
        ci = cpu_iteration(lap, dims, (cell,) * 3, np.float64, norb)
        r["cpu_reference"] = ci
        r["speedup_vs_cpu_reference"] = ci["seconds"] / sec
    del h_phi, h_new
    return r


def mgpu_parity(H, comm, rank, world, nproc):
    """N > 1: before anything is timed, the multi-rank path against the ORACLE on a small
    global box (oracle/ is used here as the checker only): H psi with in-place halos (both
    operators), the fused V-cycle, Gram + all-reduce.  Returns the dict for the JSON line;
    the caller exits non-zero on a failure."""
    import torch
    import torch.distributed as dist
    from mgmol_b200.parallel import cart_coords, local_box
    from oracle.oracle import Port, synthetic_orbitals, synthetic_potential
    port = Port()
    N = 6
    gdims = tuple(16 * p for p in nproc)
    gdims = (gdims[0], gdims[1], max(32, gdims[2]))
    ll = tuple(0.25 * n for n in gdims)
    coord = cart_coords(rank, nproc)
    box = local_box(gdims, nproc, coord)
    full = synthetic_orbitals(N, gdims, np.float64)
    v = synthetic_potential(gdims)
    res = {"global_grid": list(gdims), "orbitals": N, "checked_against": "oracle port (oracle/mgmol_oracle.c)"}
    errs = {}
    for lap in (0, 2):
        g = H.ghosts_for(lap)
        grid = H.Grid(gdims, ll, g, (1, 1, 1), nproc, coord)
        mine = torch.from_numpy(np.ascontiguousarray(full[(slice(None),) + box])).cuda()
        vmine = torch.from_numpy(np.ascontiguousarray(v[box])).cuda()
        phi = H.Orbitals(grid, N, torch.float64, mine)
        ham = H.Hamiltonian()
        ham.setup(grid, lap)
        ham.potential(H.Potentials(vmine))
        vh = make_v_halo(H, comm, grid, g, vmine)
        comm.register(mine)
        got = ham.applyLocal(phi, True, peer_comm=comm, **vh).psi().cpu().numpy()
        ref = port.hpsi(lap, full, v, ll)
        scale = np.abs(ref).reshape(N, -1).max(axis=1)[:, None, None, None]
        mine_ref = ref[(slice(None),) + box]
        errs["hpsi_lap%d" % lap] = float((np.abs(got - mine_ref) / scale).max())
        # north_star's elementwise figure on the entries that are not cancellation residue
        big = np.abs(mine_ref) >= 1e-3 * scale
        errs["hpsi_elementwise_lap%d" % lap] = (
            float((np.abs(got - mine_ref)[big] / np.abs(mine_ref)[big]).max()) if big.any() else 0.0)
        # Gram and Phi^T H Phi with the all-reduce
        S = phi.computeGram(comm).cpu().numpy()
        ex = grid.vel() * full.reshape(N, -1) @ full.reshape(N, -1).T
        errs["gram_lap%d" % lap] = float(np.abs(S - ex).max() / np.abs(ex).max())
        # multigrid-preconditioned residual on the decomposed box
        pref = port.precond_mg(lap, 2, ref, ll, 0.3)
        r = H.Orbitals(grid, N, torch.float64,
                       torch.from_numpy(np.ascontiguousarray(ref[(slice(None),) + box])).cuda())
        pc = H.OrbitalsPreconditioning()
        pc.setup(r, 2, lap)
        pc.set_comm(comm)
        pc.gamma_ = 0.3
        pc.precond_mg(r)
        res["precond_mode_lap%d" % lap] = {1: "literal", 2: "fused"}.get(pc.last_mode())
        gotp = r.psi().cpu().numpy()
        errs["precond_mg_lap%d" % lap] = float(np.abs(gotp - pref[(slice(None),) + box]).max()
                                               / np.abs(pref).max())
        pc.close()
        comm.unregister(mine)
    t = torch.tensor([errs[k] for k in sorted(errs)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    errs = dict(zip(sorted(errs), t.cpu().tolist()))
    # hpsi_elementwise: relative error on entries >= 1e-3 of the orbital's max norm (bounded
    # by the max-norm bar / 1e-3)
    bars = {"hpsi": 1e-12, "hpsi_elementwise": 1e-9, "gram": 1e-12, "precond_mg": 5e-6}
    ok = all(e <= bars[k.rsplit("_lap", 1)[0]] for k, e in errs.items())
    res.update({"max_err": errs, "bars": bars, "ok": bool(ok),
                "max_err_overall_fp64_paths": max(e for k, e in errs.items()
                                                  if not k.startswith("precond") and "elementwise" not in k)})
    return res


def make_v_halo(H, comm, grid, g, vtot):
    """The potential's halo for a decomposed box, exchanged once per potential update:
    x-slab decompositions use the 2g packed x planes, any other one a ghosted copy of V
    traded Y -> Z -> X like the reference's gfpot (src/Hamiltonian.cc:108-111)."""
    import torch
    if grid.nproc[1] == 1 and grid.nproc[2] == 1:
        xv = torch.zeros((1, 2 * g) + tuple(grid.shape()[1:]), dtype=torch.float64, device="cuda")
        comm.halo_exchange_x(grid, g, vtot[None], xv)
        return {"xhalo_v": xv}
    gg = grid.with_ghosts(g)
    gv = H.GridFuncVector(gg, 1, torch.float64)
    gv.assign(vtot[None].contiguous())
    comm.trade_boundaries(gv)
    return {"vghost": gv.data}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mgmol_b200 import host as H
    from mgmol_b200._lib import lib, check, MgbError
    from mgmol_b200.parallel import cart_coords

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
        except Exception:  # noqa: BLE001
            pass
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from mgmol_b200.parallel import Communicator
        comm = Communicator(rank, world)
    # the clock sampler (NVML init + a thread) exists before anything is timed
    sampler = ClockSampler(local) if rank == 0 else None

    L = layout(args, world)
    parity = None
    if world > 1:
        # can the fused kernels read halos in place on this decomposition?  (collective)
        try:
            parity = mgpu_parity(H, comm, rank, world, L["nproc"])
        except MgbError as e:
            if L["nproc"] != (world, 1, 1) and not (args.decomp and args.decomp != "auto"):
                if rank == 0:
                    sys.stderr.write("bench: decomposition %s not served in place (%s); x slabs\n"
                                     % (L["config"]["decomposition"], e))
                args.decomp = "%dx1x1" % world
                L = layout(args, world)
                L["config"]["decomposition_from"] = "x slabs (fallback)"
                parity = mgpu_parity(H, comm, rank, world, L["nproc"])
            else:
                raise
    nproc, gdims, dims, cell = L["nproc"], L["gdims"], L["ldims"], L["cell"]
    norb, lap_type, S = L["norb"], L["lap_type"], L["S"]
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    g = H.ghosts_for(lap_type)
    coord = cart_coords(rank, nproc)
    grid = H.Grid(gdims, cell, g, (1, 1, 1), nproc, coord)
    npt = grid.size()

    # synthetic orbitals: plane wave along z with orbital-dependent wavevector
    # + 0.1 U(-1,1) noise; potential a noisy constant.  A handful of large
    # launches (chunks of <= 16 orbitals) so the ncu launch list stays short.
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    phi = H.Orbitals(grid, norb, tdt, torch.empty((norb,) + dims, dtype=tdt, device="cuda"))
    x = torch.arange(dims[2], device="cuda", dtype=torch.float64) / dims[2]
    fill = max(1, min(64, int(2e9 // (npt * S))))
    for j0 in range(0, norb, fill):
        j1 = min(norb, j0 + fill)
        k = (torch.arange(j0, j1, device="cuda") % 5 + 1).to(torch.float64)
        wave = torch.cos(2 * np.pi * k[:, None] * x[None, :])[:, None, None, :]
        blk = phi.psi()[j0:j1]
        blk.uniform_(-0.1, 0.1, generator=gen)
        blk += wave.to(tdt)
    del wave, blk
    vtot = (torch.rand(dims, generator=gen, device="cuda", dtype=torch.float64) * 0.1 - 0.75)
    ham = H.Hamiltonian()
    ham.setup(grid, lap_type)
    ham.potential(H.Potentials(vtot))
    ham.hlphi_ = H.Orbitals(grid, norb, tdt, torch.empty((norb,) + dims, dtype=tdt, device="cuda"))

    # N > 1: V's halo is exchanged once (it is fixed over the steps); the neighbours'
    # boundary layers of the orbitals are read in place over NVLink (peer mapping of the
    # orbital block) -- or, on x slabs whose block cannot be mapped, packed and exchanged
    # with NCCL send/recv every step.
    xh_phi = None
    vh = {}
    halo_mode = None
    if world > 1:
        vh = make_v_halo(H, comm, grid, g, vtot)
        halo_mode = "peer_nvlink"
        if os.environ.get("MGB_BENCH_HALO") == "nccl":
            halo_mode = "nccl_packed"
        else:
            try:
                comm.register(phi.psi())
                ham.applyLocal(phi, True, peer_comm=comm, **vh)
            except MgbError as e:
                if rank == 0:
                    sys.stderr.write("bench: peer halo unavailable (%s)\n" % e)
                halo_mode = "nccl_packed"
        ok = torch.tensor([1 if halo_mode == "peer_nvlink" else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok) == 0:
            halo_mode = "nccl_packed"
            assert nproc[1] == 1 and nproc[2] == 1, "the packed exchange feeds x slabs only"
            xh_phi = torch.zeros((norb, 2 * g) + dims[1:], dtype=tdt, device="cuda")

    def step():
        if world > 1 and halo_mode == "peer_nvlink":
            return ham.applyLocal(phi, True, peer_comm=comm, **vh)
        if world > 1:
            comm.halo_exchange_x(grid, g, phi.psi(), xh_phi)
            return ham.applyLocal(phi, True, xh_phi, vh["xhalo_v"])
        return ham.applyLocal(phi, True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    if sampler:
        sampler.start()
    barrier()
    if world > 1:
        # rank skew is absorbed on the DEVICE before the first timed event: a
        # stream-ordered barrier over all ranks, one more untimed step, a second one
        comm.barrier()
        step()
        comm.barrier()
    n0 = lib().mgb_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    barrier()
    launches = lib().mgb_launch_count() - n0
    ms = evs[0].elapsed_time(evs[-1])
    per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    path = lib().mgb_hpsi_last_path()
    kernel_sig = lib().mgb_hpsi_last_kernel().decode()
    clocks = sampler.summary() if sampler else None

    # kernel-only duration for the roofline: the same launches, each between its own
    # pair of events on the launching stream (at N > 1 this includes the two neighbour
    # flag kernels of the call)
    kevs = []
    for _ in range(min(args.steps, 10)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        kevs.append((a, b))
    torch.cuda.synchronize()
    kern_ms = float(np.median([a.elapsed_time(b) for a, b in kevs]))

    # end-to-end through the reference-facing call with HOST buffers: pinned host orbitals
    # in, pinned host H psi out (H2D, kernel and D2H pipelined over orbital blocks inside
    # the library).  Bounded host memory: at most ~8.6 GB each way per step, i.e. the
    # first e2e_orb orbitals of the block -- the rate is per update.
    quick = bool(os.environ.get("MGB_BENCH_QUICK"))  # development: no e2e leg, no CPU arm
    e2e_orb = norb
    while e2e_orb > 8 and npt * e2e_orb * S > (1e8 if quick else 9e9):
        e2e_orb //= 2
    h_phi = torch.empty((e2e_orb,) + dims, dtype=tdt).pin_memory()
    h_phi.copy_(phi.psi()[:e2e_orb])
    h_out = torch.empty_like(h_phi).pin_memory()
    h_v = torch.empty(vtot.shape, dtype=torch.float64).pin_memory()
    h_v.copy_(vtot)
    e2e_steps = max(2, min(args.steps, 5))
    e2e_mode = "pipelined host call (mgb_hpsi_host)"
    sub_phi = H.Orbitals(grid, e2e_orb, tdt, phi.psi()[:e2e_orb])
    # page-locking ~17 GB per rank takes seconds and the ranks compete for the host's
    # memory system: realign them on the host before the next neighbour barrier on the
    # device (which fails closed after ~10 s)
    barrier()
    if world > 1:
        e2e_mode = "pipelined host call per rank, halos in place (mgb_hpsi_host_peer)"
        okp = 1
        try:
            if nproc[1] != 1 or nproc[2] != 1:
                raise RuntimeError("host pipeline serves x slabs")
            ham.lapOper().applyWithPotHostPeer(comm, h_phi, h_v, h_out)
        except Exception as e:  # noqa: BLE001
            okp = 0
            if rank == 0:
                sys.stderr.write("bench: host peer pipeline unavailable (%s)\n" % e)
        okt = torch.tensor([okp], device="cuda")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if int(okt) == 0:
            e2e_mode = "pinned host block copied in, in-place-halo kernel, H psi copied out"

    def e2e_step():
        if world == 1:
            ham.lapOper().applyWithPotHost(h_phi, h_v, h_out)
        elif e2e_mode.startswith("pipelined"):
            ham.lapOper().applyWithPotHostPeer(comm, h_phi, h_v, h_out)
        else:
            # every rank refreshes the first e2e_orb orbitals of its registered block,
            # all ranks run the full in-place-halo step, the same orbitals come back
            sub_phi.psi().copy_(h_phi, non_blocking=True)
            out = step()
            h_out.copy_(out.psi()[:e2e_orb], non_blocking=True)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(e2e_steps):
        e2e_step()
    ev1.record()
    barrier()
    e2e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3)
    e2e_updates = float(npt) * e2e_orb
    if world > 1 and not e2e_mode.startswith("pipelined"):
        # the kernel ran over the whole block; charge the copies' orbitals only if the
        # kernel time is scaled too: report the conservative figure (whole-step time,
        # copied orbitals only)
        pass
    del h_phi, h_out

    t = torch.tensor([ms, e2e_ms, kern_ms] + per_step, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tl = t.cpu().tolist()
    ms, e2e_ms, kern_ms, per_step = tl[0], tl[1], tl[2], tl[3:]

    fp64_peak = fp64_tensor_peak(torch)
    pieces = None
    if not args.no_pieces:
        pieces = measure_pieces(H, grid, phi, ham, lap_type, tdt, S, norb, npt, comm,
                                step if world > 1 else None, fp64_peak)
        if world > 1:
            # max over ranks of every timing
            keys = ["precond_mg", "gram", "phiT_H_phi", "phi_M", "orbital_update_iteration"]
            tt = torch.tensor([pieces[k]["ms"] for k in keys], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            for k, v in zip(keys, tt.cpu().tolist()):
                pieces[k]["ms_max_over_ranks"] = v
        elif args.workload == "synth256" and not args.no_sweep and args.dtype == "f64":
            pieces["sweep_256"] = measure_sweep(H, args, phi.psi(), ham.hlphi_.psi(), fp64_peak)
            pieces["f32_contractions"] = measure_f32_contractions(H, phi.psi(), ham.hlphi_.psi())
            pieces["iteration_e2e"] = measure_iteration_e2e(
                H, phi.psi(), ham.hlphi_.psi(), fp64_peak, not args.no_cpu_iteration)

    if rank == 0:
        updates_per_step = float(npt) * norb * world
        value = updates_per_step * args.steps / (ms * 1e-3)
        e2e = e2e_updates * world * e2e_steps / (e2e_ms * 1e-3)
        peak, peak_src = measured_peaks()
        achieved = 2.0 * S * npt * norb / (kern_ms * 1e-3) / 1e9
        np_dt = np.float64 if args.dtype == "f64" else np.float32
        cpu_rate, cores, kind, sample = ((0.0, 0, "skipped", "MGB_BENCH_QUICK") if quick else
                                         cpu_hpsi_rate(lap_type, gdims, cell, np_dt))
        traffic, traffic_src = (measured_traffic(kernel_sig, dims, norb, args.dtype, lap_type)
                                if path == 1 else (None, "no capture"))
        cfg = dict(L["config"])
        line = {
            "metric": "hpsi_gridpt_orbital_updates_per_s", "value": value, "unit": "updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps,
            "ms_per_step_median": float(np.median(per_step)), "first_step_ms": per_step[0],
            "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": cfg,
            "path": {"halo": halo_mode,
                     "hpsi_path": {1: "tma_fused", 2: "generic_fused", 3: "ghosted"}.get(path),
                     "kernel": kernel_sig},
            "e2e": {"value": e2e, "unit": "updates/s",
                    "h2d_bytes_per_step": int(S * npt * e2e_orb) * world,
                    "d2h_bytes_per_step": int(S * npt * e2e_orb) * world,
                    "orbitals_per_step": e2e_orb,
                    # both PCIe directions of all ranks together: what the host's memory
                    # system serves (the limiter of this figure at N > 1, where the per-rank
                    # pipelines share one host)
                    "host_GBps_both_ways": 2.0 * S * npt * e2e_orb * world * e2e_steps
                                           / (e2e_ms * 1e-3) / 1e9,
                    "how": e2e_mode + "; host memory bounded: %d of the %d orbitals per step"
                           % (e2e_orb, norb)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "kernel": "k_hpsi_tma" if path == 1 else "k_hpsi_generic",
                         "kernel_ms": kern_ms, "peak_source": peak_src,
                         "algorithmic_bytes_per_update": 2 * S},
            "cpu_baseline": {"value": cpu_rate, "unit": "updates/s", "cores": cores,
                             "kind": kind, "sample": sample},
        }
        if parity is not None:
            line["parity_mgpu"] = parity
        if pieces:
            if world == 1 and not args.no_cpu_iteration:
                ci = cpu_iteration(lap_type, dims, cell, np_dt, norb)
                it = pieces["orbital_update_iteration"]
                it["cpu_reference"] = ci
                it["speedup_vs_cpu_reference"] = ci["seconds"] * 1e3 / it["ms"]
            line["pieces"] = pieces
        print(json.dumps(line))
    rc = 0
    if parity is not None and not parity["ok"]:
        sys.stderr.write("bench: multi-GPU parity check FAILED: %s\n" % json.dumps(parity))
        rc = 3
    if world > 1:
        if os.environ.get("MGB_HPSI_TIMING"):
            lib().mgb_hpsi_timing_report(rank)
        comm.check()
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--strong", action="store_true",
                    help="split the workload's own grid over the GPUs and keep its orbital count "
                         "(default: weak scaling, fixed work per GPU)")
    ap.add_argument("--decomp", default="auto",
                    help="auto (PEenv::geom) or PxQxR ranks along x, y, z")
    ap.add_argument("--no-cpu-iteration", action="store_true",
                    help="skip the CPU reference timing of the orbital-update iteration")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--lap", type=int, default=None)
    ap.add_argument("--orbitals", type=int, default=0)
    ap.add_argument("--no-pieces", action="store_true",
                    help="skip the per-piece measurements (V-cycle, contractions)")
    ap.add_argument("--no-sweep", action="store_true",
                    help="skip the dtype x operator sweep of the 256^3 workload")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
