"""Decomposed Poisson solvers on the device, launched by torchrun (one rank per GPU)
from tests/test_multi_gpu.py or by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_poisson_worker.py

Every rank keeps its block of the reference's 3-D decomposition of the golden cases
(tests/golden/reference_poisson.npz: solutions of the compiled reference solvers on the
whole box), runs PoissonMG / PoissonPCG / Hartree with the ranks' communicator --
boundary trades over NCCL, all-reduced norms, the gathered coarse level of
src/pb/Vcycle.h:66-143 -- and compares its block of the solution with the golden one.
Exit code 0 on every rank = parity."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mgmol_b200 import host as H  # noqa: E402
from mgmol_b200.parallel import Communicator, cart_coords, local_box  # noqa: E402
from mgmol_b200.poisson import Hartree, PoissonMG, PoissonPCG  # noqa: E402
from poisson_cases import CASES, DEFAULTS, PCG_CASES, charge, guess, key  # noqa: E402


def decompositions(world):
    return {2: [(2, 1, 1), (1, 2, 1), (1, 1, 2)],
            4: [(2, 2, 1), (1, 2, 2), (2, 1, 2)],
            8: [(2, 2, 2)]}[world]


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Communicator(rank, world)
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_poisson.npz")))
    fails = []

    def check(name, ok, info=""):
        if not ok:
            fails.append(name)
        if rank == 0:
            print("%-66s %s %s" % (name, "ok" if ok else "FAIL", info), flush=True)

    case = {c[0]: c for c in CASES}
    tdt = {np.float64: torch.float64, np.float32: torch.float32}
    for nproc in decompositions(world):
        coord = cart_coords(rank, nproc)
        for tag, lts, dts in (("per", (0, 2), (np.float64, np.float32)),
                              ("dir", (0, 1), (np.float64,)),
                              ("mix", (2,), (np.float64,)),
                              ("deep_per", (0,), (np.float64,)),
                              ("deep_dir", (2,), (np.float64, np.float32)),
                              ("conv", (0,), (np.float64,))):
            _, dims, ll, bc, kw = case[tag]
            par = dict(DEFAULTS, **kw)
            box = local_box(dims, nproc, coord)
            for lt in lts:
                for dt in dts:
                    ref, st = gold[key(tag, lt, dt)], gold[key(tag, lt, dt) + "_stats"]
                    grid = H.Grid(dims, ll, 1, bc, nproc, coord)
                    s = PoissonMG(grid, lt, tdt[dt], comm=comm)
                    s.setup(par["nu1"], par["nu2"], par["max_sweeps"], par["tol"],
                            par["max_nlevels"])
                    vh = torch.from_numpy(np.ascontiguousarray(guess(dims, dt)[box])).cuda()
                    rho = torch.from_numpy(np.ascontiguousarray(charge(dims, bc, dt)[box])).cuda()
                    conv = s.solve(vh, rho)
                    err = np.abs(vh.cpu().numpy().astype(np.float64) - ref[box]).max() / np.abs(ref).max()
                    eps = 1e-13 if dt == np.float64 else 2e-6
                    ok = (err <= eps and conv == bool(st[0]) and s.getNbSweeps() == int(st[1])
                          and abs(s.getFinalResidual() - st[2]) <= 1e-5 * abs(st[2]) + 1e-300)
                    check("Mgm %dx%dx%d %s lap%d %s" % (nproc + (tag, lt, np.dtype(dt).name)), ok,
                          "err %.1e sweeps %d" % (err, s.getNbSweeps()))
        # PCGSolver: cases whose local boxes reach the single-rank depth
        pcase = {c[0]: c for c in PCG_CASES}
        for tag, lt, ok_nproc in (("pcg_per", 0, ((2, 1, 1), (2, 2, 1))),
                                  ("pcg_shallow", 2, ((2, 1, 1), (1, 2, 1), (1, 1, 2), (2, 2, 1),
                                                      (1, 2, 2), (2, 1, 2), (2, 2, 2)))):
            if nproc not in ok_nproc:
                continue
            _, dims, ll, bc, kw = pcase[tag]
            par = dict(DEFAULTS, **kw)
            box = local_box(dims, nproc, coord)
            dt = np.float64
            ref, st = gold[key(tag, lt, dt)], gold[key(tag, lt, dt) + "_stats"]
            grid = H.Grid(dims, ll, 1, bc, nproc, coord)
            s = PoissonPCG(grid, lt, torch.float64, comm=comm)
            s.setup(par["nu1"], par["nu2"], par["max_sweeps"], par["tol"], par["max_nlevels"])
            vh = torch.from_numpy(np.ascontiguousarray(guess(dims, dt)[box])).cuda()
            rho = torch.from_numpy(np.ascontiguousarray(charge(dims, bc, dt)[box])).cuda()
            conv = s.solve(vh, rho)
            err = np.abs(vh.cpu().numpy() - ref[box]).max() / np.abs(ref).max()
            check("PCG %dx%dx%d %s lap%d" % (nproc + (tag, lt)),
                  err <= 1e-9 and conv == bool(st[0]), "err %.1e" % err)
        # Hartree on the decomposed box against Hartree on one rank (same library; the
        # single-rank class is pinned to the reference in tests/test_gpu_poisson.py)
        _, dims, ll, bc, kw = case["per"]
        box = local_box(dims, nproc, coord)
        rho_g = charge(dims, bc, np.float64) + 0.3
        rhoc_g = np.full(dims, rho_g.mean())
        one = Hartree(H.Grid(dims, ll, 1, bc), 0)
        one.solve(torch.from_numpy(rho_g).cuda(), torch.from_numpy(rhoc_g).cuda())
        many = Hartree(H.Grid(dims, ll, 1, bc, nproc, coord), 0, comm=comm)
        many.solve(torch.from_numpy(np.ascontiguousarray(rho_g[box])).cuda(),
                   torch.from_numpy(np.ascontiguousarray(rhoc_g[box])).cuda())
        v1 = one.vh().cpu().numpy()[0]
        vm = many.vh().cpu().numpy()[0]
        err = np.abs(vm - v1[box]).max() / np.abs(v1).max()
        ei = abs(many.IntVhRho() - one.IntVhRho()) / abs(one.IntVhRho())
        check("Hartree %dx%dx%d vh and Int vh rho" % nproc, err <= 1e-13 and ei <= 1e-12,
              "err %.1e int %.1e" % (err, ei))

    comm.check()
    torch.cuda.synchronize()
    flag = torch.tensor([len(fails)], device="cuda")
    dist.all_reduce(flag)
    comm.close()
    dist.destroy_process_group()
    if int(flag) != 0:
        print("rank %d failures: %s" % (rank, fails), flush=True)
        sys.exit(1)
    if rank == 0:
        print("multi-gpu poisson parity ok on %d ranks" % world, flush=True)


if __name__ == "__main__":
    main()
