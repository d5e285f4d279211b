"""Multi-GPU parity worker, launched by torchrun (one rank per GPU) from
tests/test_multi_gpu.py or by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_worker.py

Every rank builds the same global fields, keeps its block of the reference's
3-D decomposition, runs the multi-rank path, and compares with the single-rank
result of the same library on the global grid (itself pinned to the oracle by
tests/test_gpu_parity.py).  Exit code 0 on every rank = parity."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mgmol_b200 import host as H  # noqa: E402
from mgmol_b200.parallel import (Communicator, cart_coords, color_maps,  # noqa: E402
                                 ghosted_color_maps, local_box)


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Communicator(rank, world)
    fails = []

    def check(name, ok):
        if not ok:
            fails.append(name)
        if rank == 0:
            print("%-60s %s" % (name, "ok" if ok else "FAIL"), flush=True)

    gen = torch.Generator(device="cuda").manual_seed(77)  # same on every rank
    N = 6
    for dt in (torch.float64, torch.float32):
        for lap in (0, 2):
            g = H.ghosts_for(lap)
            for bc in ((1, 1, 1), (0, 0, 0)):
                # ---- fused H psi on an x-split ------------------------------------
                gdims = (8 * world, 16, 32)
                ll = (0.25 * gdims[0], 4.0, 8.0)
                full = (torch.rand((N,) + gdims, generator=gen, device="cuda",
                                   dtype=torch.float64) - 0.5).to(dt)
                v = torch.rand(gdims, generator=gen, device="cuda", dtype=torch.float64) - 0.7
                ggrid = H.Grid(gdims, ll, g, bc)
                ref = torch.empty_like(full)
                H.LapFactory.createLap(ggrid, lap).applyWithPot(full, v, ref)
                nproc = (world, 1, 1)
                coord = cart_coords(rank, nproc)
                box = local_box(gdims, nproc, coord)
                grid = H.Grid(gdims, ll, g, bc, nproc, coord)
                mine = full[(slice(None),) + box].contiguous()
                vmine = v[box].contiguous()
                xh = torch.zeros((N, 2 * g) + gdims[1:], dtype=dt, device="cuda")
                xv = torch.zeros((1, 2 * g) + gdims[1:], dtype=torch.float64, device="cuda")
                comm.halo_exchange_x(grid, g, mine, xh)
                comm.halo_exchange_x(grid, g, vmine[None].contiguous(), xv)
                out = torch.empty_like(mine)
                H.LapFactory.createLap(grid, lap).applyWithPot(mine, vmine, out, xh, xv)
                check("hpsi x-split %s lap%d bc%s" % (dt, lap, bc),
                      torch.equal(out, ref[(slice(None),) + box]))
                # ---- the same with the neighbours' planes read in place (peer
                # mapping over NVLink): no packed exchange of phi
                try:
                    comm.register(mine)
                    out2 = torch.full_like(mine, float("nan"))
                    H.LapFactory.createLap(grid, lap).applyWithPotPeer(
                        comm, mine, vmine, out2, xv)
                    check("hpsi x-split peer reads %s lap%d bc%s" % (dt, lap, bc),
                          torch.equal(out2, ref[(slice(None),) + box]))
                    comm.unregister(mine)
                    # ---- and from HOST buffers: per-rank pipeline, halos in
                    # place from the neighbours' input rings
                    if bc == (1, 1, 1):
                        hp = mine.cpu().pin_memory()
                        hv = vmine.cpu().pin_memory()
                        ho = torch.empty_like(hp).pin_memory()
                        H.LapFactory.createLap(grid, lap).applyWithPotHostPeer(
                            comm, hp, hv, ho, chunk=4)
                        check("hpsi x-split host pipeline %s lap%d" % (dt, lap),
                              torch.equal(ho.cuda(), ref[(slice(None),) + box]))
                except H.MgbError as e:
                    check("hpsi x-split peer reads %s lap%d bc%s: %s" % (dt, lap, bc, e), False)

    # ---- fused H psi on every decomposition of the ranks over x, y, z: all halos (faces,
    # and the edge lines the Mehrstellen stencil reaches) read in place from the
    # neighbours' blocks; bit-identical to the single-rank kernel on the global box
    decomps = {2: [(1, 1, 2), (1, 2, 1), (2, 1, 1)],
               4: [(1, 2, 2), (2, 2, 1), (2, 1, 2), (1, 1, 4), (1, 4, 1)],
               8: [(2, 2, 2), (1, 2, 4), (4, 2, 1), (2, 4, 1)]}.get(world, [(world, 1, 1)])
    for nproc in decomps:
        for dt in (torch.float64, torch.float32):
            for lap in (0, 2):
                g = H.ghosts_for(lap)
                for bc in ((1, 1, 1), (0, 0, 0)):
                    gdims = (8 * nproc[0], 16 * nproc[1], 32 * nproc[2])
                    ll = (0.25 * gdims[0], 0.3 * gdims[1], 0.2 * gdims[2])
                    full = (torch.rand((N,) + gdims, generator=gen, device="cuda",
                                       dtype=torch.float64) - 0.5).to(dt)
                    v = torch.rand(gdims, generator=gen, device="cuda", dtype=torch.float64) - 0.7
                    ggrid = H.Grid(gdims, ll, g, bc)
                    ref = torch.empty_like(full)
                    H.LapFactory.createLap(ggrid, lap).applyWithPot(full, v, ref)
                    coord = cart_coords(rank, nproc)
                    box = local_box(gdims, nproc, coord)
                    grid = H.Grid(gdims, ll, g, bc, nproc, coord)
                    mine = full[(slice(None),) + box].contiguous()
                    vmine = v[box].contiguous()
                    gv = H.GridFuncVector(grid, 1, torch.float64)
                    gv.assign(vmine[None].contiguous())
                    comm.trade_boundaries(gv)
                    try:
                        comm.register(mine)
                        out = torch.full_like(mine, float("nan"))
                        op = H.LapFactory.createLap(grid, lap)
                        op.applyWithPotPeer(comm, mine, vmine, out, vghost=gv.data)
                        op.applyWithPotPeer(comm, mine, vmine, out, vghost=gv.data)
                        check("hpsi %dx%dx%d in-place halos %s lap%d bc%s" % (nproc + (dt, lap, bc)),
                              torch.equal(out, ref[(slice(None),) + box]))
                        comm.unregister(mine)
                    except H.MgbError as e:
                        check("hpsi %dx%dx%d in-place halos %s lap%d bc%s: %s"
                              % (nproc + (dt, lap, bc, e)), False)

    # ---- ghosted Y -> Z -> X exchange on every 2-way / n-way split ---------------
    for dt in (torch.float64, torch.float32):
        for gw in (1, 2):
            for axis in range(3):
                nproc = [1, 1, 1]
                nproc[axis] = world
                nproc = tuple(nproc)
                gdims = [8, 12, 16]
                gdims[axis] *= world
                gdims = tuple(gdims)
                for bc in ((1, 1, 1), (0, 0, 0)):
                    full = (torch.rand((3,) + gdims, generator=gen, device="cuda",
                                       dtype=torch.float64) - 0.5).to(dt)
                    # single-rank reference: ghosted global block, traded locally
                    ggrid = H.Grid(gdims, (1.0, 1.0, 1.0), gw, bc)
                    gg = H.GridFuncVector(ggrid, 3, dt)
                    gg.assign(full)
                    gg.trade_boundaries()
                    coord = cart_coords(rank, nproc)
                    box = local_box(gdims, nproc, coord)
                    grid = H.Grid(gdims, (1.0, 1.0, 1.0), gw, bc, nproc, coord)
                    lg = H.GridFuncVector(grid, 3, dt)
                    lg.assign(full[(slice(None),) + box].contiguous())
                    comm.trade_boundaries(lg)
                    # my ghosted block = the window of the global ghosted block
                    win = tuple(slice(b.start, b.stop + 2 * gw) for b in box)
                    expect = gg.data[(slice(None),) + win]
                    # Dirichlet: the global block's own first-layer zeroing only
                    # applies to the rank that owns the low face; compare where
                    # both definitions agree (everything except that layer on
                    # ranks that do not own it is identical by construction)
                    check("ghosted exchange %s g%d axis%d bc%s" % (dt, gw, axis, bc),
                          torch.equal(lg.data, expect))

    # ---- multigrid preconditioner on a decomposed box ------------------------------
    # x-split: fused kernels reading the neighbours' planes in place (mode 2) and
    # the literal sequence over the packed exchange (mode 1); y-split: literal.
    # Each must equal the single-rank V-cycle of the same mode on the global box
    # bit for bit (same arithmetic per point).
    for dt in (torch.float64, torch.float32):
        for lap in (0, 2):
            g = H.ghosts_for(lap)
            for bc in ((1, 1, 1), (0, 0, 0)):
                cases = [((world, 1, 1), (2, 1)), ((1, world, 1), (2, 1)), ((1, 1, world), (2,))]
                cases += [(d, (2,)) for d in decomps if sum(1 for q in d if q > 1) > 1]
                for nproc, modes in cases:
                    gdims = (16 * nproc[0], 16 * nproc[1], 32 * nproc[2])
                    ll = tuple(0.25 * d for d in gdims)
                    axis = "%dx%dx%d" % nproc
                    coord = cart_coords(rank, nproc)
                    box = local_box(gdims, nproc, coord)
                    full = (torch.rand((4,) + gdims, generator=gen, device="cuda",
                                       dtype=torch.float64) - 0.5).to(dt)
                    for mode in modes:
                        ggrid = H.Grid(gdims, ll, g, bc)
                        gorb = H.Orbitals(ggrid, 4, dt, full.clone())
                        pc = H.OrbitalsPreconditioning()
                        pc.setup(gorb, 2, lap)
                        pc.set_mode(mode)
                        pc.gamma_ = 0.3
                        pc.precond_mg(gorb)
                        pc.close()
                        grid = H.Grid(gdims, ll, g, bc, nproc, coord)
                        orb = H.Orbitals(grid, 4, dt, full[(slice(None),) + box].contiguous())
                        pcl = H.OrbitalsPreconditioning()
                        pcl.setup(orb, 2, lap)
                        pcl.set_comm(comm)
                        pcl.set_mode(mode)
                        pcl.gamma_ = 0.3
                        pcl.precond_mg(orb)
                        pcl.precond_mg(orb)  # second call: work blocks reused
                        pc2 = H.OrbitalsPreconditioning()
                        pc2.setup(gorb, 2, lap)
                        pc2.set_mode(mode)
                        pc2.gamma_ = 0.3
                        pc2.precond_mg(gorb)
                        pc2.close()
                        check("precond_mg %s mode%d %s lap%d bc%s" % (axis, mode, dt, lap, bc),
                              pcl.last_mode() == mode
                              and torch.equal(orb.psi(), gorb.psi()[(slice(None),) + box]))
                        pcl.close()

    # ---- LocGridOrbitals on an x-split: gid-addressed halo ----------------------
    # Every rank holds the same 5 global orbitals but in its own color order, one
    # of them only on rank 0 (absent elsewhere) and one empty slot.  Expected
    # result of color c on rank r = the single-rank result for that orbital,
    # whose global function is zero on the boxes of ranks that do not hold it.
    ncol, ngid = 6, 5
    for dt in (torch.float64, torch.float32):
        for lap in (0, 2):
            g = H.ghosts_for(lap)
            gdims = (16 * world, 16, 32)
            ll = tuple(0.25 * d for d in gdims)
            nproc = (world, 1, 1)
            coord = cart_coords(rank, nproc)
            box = local_box(gdims, nproc, coord)
            rs = np.random.RandomState(5)
            tables = []
            for r in range(world):
                perm = list(rs.permutation(ngid)) + [-1]
                if r != 0:
                    perm = [x if x != 3 else -1 for x in perm]  # gid 3 only on rank 0
                tables.append(np.array([perm]))                # subdivx = 1
            gfun = (torch.rand((ngid,) + gdims, generator=gen, device="cuda",
                               dtype=torch.float64) - 0.5).to(dt)
            for r in range(1, world):                          # gid 3 lives on rank 0 only
                gfun[(3,) + local_box(gdims, nproc, cart_coords(r, nproc))] = 0
            v = torch.rand(gdims, generator=gen, device="cuda", dtype=torch.float64) - 0.7
            ggrid = H.Grid(gdims, ll, g, (1, 1, 1))
            href = torch.empty_like(gfun)
            H.LapFactory.createLap(ggrid, lap).applyWithPot(gfun, v, href)
            gorb = H.Orbitals(ggrid, ngid, dt, gfun.clone())
            pcg = H.OrbitalsPreconditioning()
            pcg.setup(gorb, 2, lap)
            pcg.set_mode(2)
            pcg.gamma_ = 0.3
            pcg.precond_mg(gorb)
            pcg.close()
            mine_tab = tables[rank]
            mine = torch.zeros((ncol,) + tuple(b.stop - b.start for b in box), dtype=dt,
                               device="cuda")
            for c, gid in enumerate(mine_tab[0]):
                if gid >= 0:
                    mine[c] = gfun[(int(gid),) + box]
            mw, me = color_maps(mine_tab, tables[(rank - 1) % world], tables[(rank + 1) % world])
            comm.set_color_maps(mw, me)
            grid = H.Grid(gdims, ll, g, (1, 1, 1), nproc, coord)
            vmine = v[box].contiguous()
            xv = torch.zeros((1, 2 * g) + gdims[1:], dtype=torch.float64, device="cuda")
            comm.halo_exchange_x(grid, g, vmine[None].contiguous(), xv)
            comm.register(mine)
            out = torch.full_like(mine, float("nan"))
            H.LapFactory.createLap(grid, lap).applyWithPotPeer(comm, mine, vmine, out, xv)
            ok = True
            for c, gid in enumerate(mine_tab[0]):
                exp = href[(int(gid),) + box] if gid >= 0 else torch.zeros_like(out[c])
                ok = ok and torch.equal(out[c], exp)
            check("LocGridOrbitals gid-addressed peer halo: H psi %s lap%d" % (dt, lap), ok)
            comm.unregister(mine)
            orb = H.Orbitals(grid, ncol, dt, mine.clone())
            pcl = H.OrbitalsPreconditioning()
            pcl.setup(orb, 2, lap)
            pcl.set_comm(comm)
            pcl.set_mode(2)
            pcl.gamma_ = 0.3
            pcl.precond_mg(orb)
            ok = True
            for c, gid in enumerate(mine_tab[0]):
                if gid == 3:
                    # held by one rank only: its ghosts read zero on every sweep,
                    # which a global V-cycle (the function spreads) does not model
                    continue
                exp = gorb.psi()[(int(gid),) + box] if gid >= 0 else torch.zeros_like(out[c])
                ok = ok and torch.equal(orb.psi()[c], exp)
            check("LocGridOrbitals gid-addressed peer halo: V-cycle %s lap%d" % (dt, lap), ok)
            pcl.close()
            comm.set_color_maps(None, None)

    # ---- gid-addressed packed exchange on a y-split (per-slab gids) ----------------
    for dt in (torch.float64, torch.float32):
        for gw in (1, 2):
            subdivx, ncol, ngid = 2, 4, 5
            gdims = (8, 8 * world, 12)
            nproc = (1, world, 1)
            coord = cart_coords(rank, nproc)
            box = local_box(gdims, nproc, coord)
            rs = np.random.RandomState(11)
            tables = [np.stack([rs.permutation(ngid)[:ncol] for _ in range(subdivx)])
                      for _ in range(world)]
            for r in range(world):
                tables[r][rs.randint(subdivx), rs.randint(ncol)] = -1
            # data of rank r, color c: a field that encodes (r, c)
            blocks = [(torch.rand((ncol,) + gdims, generator=gen, device="cuda",
                                  dtype=torch.float64) - 0.5).to(dt) for _ in range(world)]
            grid = H.Grid(gdims, (1.0, 1.0, 1.0), gw, (1, 1, 1), nproc, coord)
            lg = H.GridFuncVector(grid, ncol, dt)
            lg.data.fill_(7.0)  # sentinel in the ghosts
            lg.assign(blocks[rank][(slice(None),) + box].contiguous())
            lg.data[:, :, :gw] = 7.0
            lg.data[:, :, -gw:] = 7.0
            south, north = (rank - 1) % world, (rank + 1) % world
            nb = [[None, None], [tables[south], tables[north]], [None, None]]
            maps = ghosted_color_maps(tables[rank], nb)
            # identity maps in x and z (single rank there: local wraps ignore them)
            comm.set_ghosted_color_maps(maps)
            comm.trade_boundaries(lg)
            ok = True
            s0 = gdims[0] // subdivx
            ny = gdims[1] // world
            for side, nbr in ((0, south), (1, north)):
                nbox = local_box(gdims, nproc, cart_coords(nbr, nproc))
                theirs = blocks[nbr][(slice(None),) + nbox]
                for iloc in range(subdivx):
                    xs = slice(gw + iloc * s0, gw + (iloc + 1) * s0)
                    for c in range(ncol):
                        src = int(maps[1, side, iloc, c])
                        mine_rows = (lg.data[c, xs, :gw, gw:-gw] if side == 0
                                     else lg.data[c, xs, gw + ny:, gw:-gw])
                        if src < 0:
                            exp = torch.full_like(mine_rows, 7.0)
                        else:
                            rows = (slice(ny - gw, ny) if side == 0 else slice(0, gw))
                            exp = theirs[src, iloc * s0:(iloc + 1) * s0, rows, :]
                        ok = ok and torch.equal(mine_rows, exp)
            check("gid-addressed packed exchange y-split %s g%d" % (dt, gw), ok)
            # identity tables reproduce the slot-for-slot exchange bit for bit
            ident = [np.stack([np.arange(ncol)] * subdivx)] * world
            comm.set_ghosted_color_maps(ghosted_color_maps(
                ident[rank], [[None, None], [ident[south], ident[north]], [None, None]]))
            a = H.GridFuncVector(grid, ncol, dt)
            a.assign(blocks[rank][(slice(None),) + box].contiguous())
            comm.trade_boundaries(a)
            comm.set_ghosted_color_maps(None)
            b2 = H.GridFuncVector(grid, ncol, dt)
            b2.assign(blocks[rank][(slice(None),) + box].contiguous())
            comm.trade_boundaries(b2)
            check("gid-addressed exchange with identity tables %s g%d" % (dt, gw),
                  torch.equal(a.data, b2.data))

    # ---- partial Gram / projected Hamiltonian + NCCL all-reduce -------------------
    for dt, tol in ((torch.float64, 1e-12), (torch.float32, 3e-6)):
        gdims = (8 * world, 16, 32)
        nproc = (world, 1, 1)
        coord = cart_coords(rank, nproc)
        box = local_box(gdims, nproc, coord)
        full = (torch.rand((37,) + gdims, generator=gen, device="cuda", dtype=torch.float64)
                - 0.5).to(dt)
        ggrid = H.Grid(gdims, (2.0 * world, 4.0, 8.0), 1)
        grid = H.Grid(gdims, (2.0 * world, 4.0, 8.0), 1, (1, 1, 1), nproc, coord)
        orb = H.Orbitals(grid, 37, dt, full[(slice(None),) + box].contiguous())
        s = orb.computeGram(comm)
        a = full.reshape(37, -1).double()
        exact = ggrid.vel() * (a @ a.t())
        check("gram + allreduce %s" % dt,
              float((s - exact).abs().max() / exact.abs().max()) <= tol)
        hl = orb.computeLocalProduct(orb, comm)
        check("phiT A + allreduce %s" % dt,
              float((hl - exact).abs().max() / exact.abs().max()) <= tol)

    comm.check()  # no neighbour barrier timed out
    torch.cuda.synchronize()
    flag = torch.tensor([len(fails)], device="cuda")
    dist.all_reduce(flag)
    comm.close()
    dist.destroy_process_group()
    if int(flag) != 0:
        print("rank %d failures: %s" % (rank, fails), flush=True)
        sys.exit(1)
    if rank == 0:
        print("multi-gpu parity ok on %d ranks" % world, flush=True)


if __name__ == "__main__":
    main()
