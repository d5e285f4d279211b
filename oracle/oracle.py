"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front ends for the two CPU checkers.

* ``Port``  : our C restatement (oracle/mgmol_oracle.c -> libmgmol_oracle.so).
* ``Ref``   : the reference's own sources compiled unmodified
              (oracle/_ref/libmgmol_ref.so, built by ``make -C oracle ref`` where
              /root/reference exists; the .so travels to the GPU box).

Both expose the same methods so a test can run against either.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under mgmol_b200/ does.

Array conventions (reference layout): an orbital block
without ghosts is a C-contiguous numpy array of shape (nfunc, nx, ny, nz)
(= column-major ``npt x nfunc`` with ``lda = npt``); a ghosted block has shape
(nfunc, nx+2g, ny+2g, nz+2g).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libmgmol_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmgmol_ref.so")

_c_int3 = ctypes.c_int * 3
_c_dbl3 = ctypes.c_double * 3


def build(ref=True, port=True):
    """Compile the checkers (building the checker is not using it)."""
    targets = []
    if port:
        targets.append("oracle")
    if ref and os.path.isdir("/root/reference/src"):
        targets.append("ref")
    if targets:
        subprocess.check_call(["make", "-s", "-C", HERE] + targets)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _dt(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return 1
    if dtype == np.float32:
        return 0
    raise TypeError("ORBDTYPE must be float32 or float64, got %s" % dtype)


def ghosts_for(lap_type):
    """src/GridFactory.h:23-51."""
    return {0: 1, 10: 1, 1: 1, 2: 2, 3: 3, 4: 4}[lap_type]


def _sfx(dtype):
    return "_f64" if _dt(dtype) else "_f32"


class MaskTables:
    """Localization masks as plain data (what the path consumes): for every
    multigrid level l, x-slab iloc and color c

      state[l][iloc, c]  <= 0 zero the slab (GridMask::mask_not_zero_ -1 / 0,
                         or gid == -1), 1 keep, 2 apply values
      values[l][(iloc, c)]  no-ghost slab (sub0_l, ny_l, nz_l) of mask values
                         for state 2 (lmasktype)

    op: 0 multiply (GridMaskMult), 1 cut |u| <= mask (GridMaskMax)."""

    def __init__(self, dims, mg_levels, subdivx, ncolors, op):
        self.dims = tuple(dims)
        self.mg_levels = mg_levels
        self.subdivx = subdivx
        self.ncolors = ncolors
        self.op = op
        self.state = [np.ones((subdivx, ncolors), np.int32) for _ in range(mg_levels + 1)]
        self.values = [dict() for _ in range(mg_levels + 1)]

    def slab_shape(self, level):
        nx, ny, nz = (d >> level for d in self.dims)
        return (nx // self.subdivx, ny, nz)

    def set(self, level, iloc, color, state, values=None):
        self.state[level][iloc, color] = state
        self.values[level].pop((iloc, color), None)
        if state == 2:
            v = np.ascontiguousarray(values, dtype=np.float64).reshape(self.slab_shape(level))
            self.values[level][(iloc, color)] = v

    def flat(self):
        """(state int32 [L, subdivx, ncolors], voff int64 same shape, values
        float64 pool) for the C entry points."""
        L = self.mg_levels + 1
        state = np.stack(self.state).astype(np.int32)
        voff = -np.ones((L, self.subdivx, self.ncolors), np.int64)
        chunks, pos = [], 0
        for l in range(L):
            for (iloc, c), v in sorted(self.values[l].items()):
                voff[l, iloc, c] = pos
                chunks.append(v.ravel())
                pos += v.size
        pool = np.concatenate(chunks) if chunks else np.zeros(1, np.float64)
        return np.ascontiguousarray(state), np.ascontiguousarray(voff), np.ascontiguousarray(pool)

    def dense(self, level, dtype=np.float64):
        """(ncolors, nx_l, ny_l, nz_l) multiplier view (only meaningful for
        op 0); zero slabs 0, kept slabs 1."""
        s0, ny, nz = self.slab_shape(level)
        out = np.ones((self.ncolors, s0 * self.subdivx, ny, nz), dtype)
        for iloc in range(self.subdivx):
            for c in range(self.ncolors):
                st = self.state[level][iloc, c]
                sl = out[c, iloc * s0:(iloc + 1) * s0]
                if st <= 0:
                    sl[...] = 0
                elif st == 2:
                    sl[...] = self.values[level][(iloc, c)]
        return out


def synthetic_masks(dims, ll, mg_levels, subdivx, gid_table, centers, radii, op, ghosts=1):
    """Mask tables built the way GridMask::init does (src/GridMask.cc:139-262:
    minimum-image distance to a centre, profile of r / rcut inside rcut, zero
    up to rcut + delta, slab states from what was found) -- a numpy
    restatement of the mask GENERATION, which is outside the hot path; it only
    has to produce realistic data (tests pin the path against the reference
    with the reference's own masks, see Ref.masks_create).  gid_table:
    (subdivx, ncolors) ints, -1 = empty slot."""
    gid_table = np.asarray(gid_table, np.int64)
    ncolors = gid_table.shape[1]
    mt = MaskTables(dims, mg_levels, subdivx, ncolors, op)
    h = [ll[d] / dims[d] for d in range(3)]
    delta0 = (np.sqrt(sum(x * x for x in h)) + 1e-8) * ghosts
    for l in range(mg_levels + 1):
        s0, ny, nz = mt.slab_shape(l)
        hl = [x * (1 << l) for x in h]
        delta = delta0 * (1 << l)
        for iloc in range(subdivx):
            xs = (np.arange(s0) + iloc * s0) * hl[0]
            ys = np.arange(ny) * hl[1]
            zs = np.arange(nz) * hl[2]
            for c in range(ncolors):
                gid = gid_table[iloc, c]
                if gid < 0:
                    mt.set(l, iloc, c, 0)
                    continue
                ctr, rc = centers[gid], radii[gid]
                d = []
                for a, cc, L in ((xs, ctr[0], ll[0]), (ys, ctr[1], ll[1]), (zs, ctr[2], ll[2])):
                    t = a - cc
                    t -= L * np.round(t / L)
                    d.append(t)
                r = np.sqrt(d[0][:, None, None] ** 2 + d[1][None, :, None] ** 2
                            + d[2][None, None, :] ** 2)
                inside = r <= rc
                if rc > 100.0 and (r < rc + delta).any():
                    mt.set(l, iloc, c, 1)
                elif not (r < rc + delta).any():
                    mt.set(l, iloc, c, -1)
                elif not inside.any():
                    mt.set(l, iloc, c, 0)
                else:
                    x = np.clip(r / rc, 0, 1)
                    if op == 0:
                        prof = np.where(x < 0.99999, 1.0, (1.0 - x) / (1.0 - 0.99999))
                    else:
                        fac = 2.0 / (0.25 * 0.25)
                        e_rc = np.exp(-fac * 0.25 * 0.25)
                        prof = np.where(x > 0.75, np.float32(
                            (np.exp(-fac * (x - 0.75) ** 2) - e_rc) / (1 - e_rc)), 1.0)
                    mt.set(l, iloc, c, 2, np.where(inside, prof, 0.0))
    return mt


class Port:
    """Our restatement (kind "port")."""

    kind = "port"

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build(ref=False)
        self.lib = ctypes.CDLL(PORT_SO)
        self.lib.orc_gamma.restype = ctypes.c_double

    # -- ghosts -----------------------------------------------------------
    def add_ghosts(self, a, g):
        a = np.ascontiguousarray(a)
        nf, nx, ny, nz = a.shape
        out = np.empty((nf, nx + 2 * g, ny + 2 * g, nz + 2 * g), a.dtype)
        getattr(self.lib, "orc_add_ghosts" + _sfx(a.dtype))(
            _c_int3(nx, ny, nz), g, _ptr(a), _ptr(out), nf)
        return out

    def strip_ghosts(self, a, g):
        a = np.ascontiguousarray(a)
        nf = a.shape[0]
        nx, ny, nz = (s - 2 * g for s in a.shape[1:])
        out = np.empty((nf, nx, ny, nz), a.dtype)
        getattr(self.lib, "orc_strip_ghosts" + _sfx(a.dtype))(
            _c_int3(nx, ny, nz), g, _ptr(a), _ptr(out), nf)
        return out

    def trade_boundaries(self, noghost, g, bc=(1, 1, 1)):
        """no-ghost block -> ghosted block with exchanged boundaries."""
        u = self.add_ghosts(noghost, g)
        nf, nx, ny, nz = noghost.shape
        getattr(self.lib, "orc_trade_boundaries" + _sfx(u.dtype))(
            _c_int3(nx, ny, nz), g, _c_int3(*bc), _ptr(u), nf)
        return u

    # -- FD kernels on ghosted blocks --------------------------------------
    def fdkernel(self, kind, v, g, h, rhs_ghosts=None):
        """kind: 0 Mehrstellen, 1 2nd, 2 4th, 100 Mehrstellen RHS (B).
        v ghosted (nfunc, ...). Returns ghosted block with interior written
        (kind 100 with rhs_ghosts=0: no-ghost block)."""
        v = np.ascontiguousarray(v)
        nf = v.shape[0]
        dims = _c_int3(*(s - 2 * g for s in v.shape[1:]))
        hh = _c_dbl3(*h)
        s = _sfx(v.dtype)
        if kind == 100:
            rg = g if rhs_ghosts is None else rhs_ghosts
            out = np.zeros((nf,) + tuple(d + 2 * rg for d in dims), v.dtype)
            getattr(self.lib, "orc_rhs_4th_Mehr1" + s)(
                dims, g, _ptr(v), _ptr(out), rg, nf)
            return out
        name = {0: "orc_del2_4th_Mehr", 1: "orc_del2_2nd", 2: "orc_del2_4th"}[kind]
        out = np.zeros_like(v)
        getattr(self.lib, name + s)(dims, g, hh, _ptr(v), _ptr(out), nf)
        return out

    def lap_rhs(self, lap_type, phi, ll, bc=(1, 1, 1)):
        """Lap::rhs per orbital on a no-ghost block: B (Laph4M), B2 (Laph4MP)
        or a copy."""
        phi = np.ascontiguousarray(phi)
        nf, nx, ny, nz = phi.shape
        if lap_type not in (0, 10):
            return phi.copy()
        g = 1
        gv = self.trade_boundaries(phi, g, bc)
        out = np.empty_like(phi)
        if lap_type == 0:
            getattr(self.lib, "orc_rhs_4th_Mehr1" + _sfx(phi.dtype))(
                _c_int3(nx, ny, nz), g, _ptr(gv), _ptr(out), 0, nf)
        else:
            getattr(self.lib, "orc_rhs_4th_Mehr2" + _sfx(phi.dtype))(
                _c_int3(nx, ny, nz), g, _ptr(gv), _ptr(out), nf)
        return out

    # -- fused H psi ---------------------------------------------------------
    def hpsi(self, lap_type, phi, vtot, ll, bc=(1, 1, 1)):
        phi = np.ascontiguousarray(phi)
        vtot = np.ascontiguousarray(vtot, dtype=np.float64)
        nf, nx, ny, nz = phi.shape
        out = np.empty_like(phi)
        rc = getattr(self.lib, "orc_hpsi" + _sfx(phi.dtype))(
            lap_type, _c_int3(nx, ny, nz), _c_dbl3(*ll), _c_int3(*bc),
            _ptr(phi), _ptr(vtot), _ptr(out), nf)
        if rc:
            raise ValueError("lap_type %d has no applyWithPot" % lap_type)
        return out

    # -- multigrid ------------------------------------------------------------
    def lap_constants(self, lap_type, dims, g, ll):
        out = _c_dbl3()
        h = _c_dbl3(*(l / d for l, d in zip(ll, dims)))
        if self.lib.orc_lap_constants(lap_type, h, out):
            raise ValueError("bad lap_type")
        return tuple(out)

    def jacobi(self, lap_type, v, f, w, g, ll, jacobi_factor, bc=(1, 1, 1)):
        v = np.array(v, dtype=np.float32, order="C")
        w = np.array(w, dtype=np.float32, order="C")
        f = np.ascontiguousarray(f, dtype=np.float32)
        nf = v.shape[0]
        dims = tuple(s - 2 * g for s in v.shape[1:])
        h = _c_dbl3(*(l / d for l, d in zip(ll, dims)))
        upd = ctypes.c_int(0)
        rc = self.lib.orc_jacobi_f32(
            lap_type, _c_int3(*dims), h, g, _c_int3(*bc), _ptr(v),
            ctypes.byref(upd), _ptr(f), _ptr(w), nf,
            ctypes.c_double(jacobi_factor))
        if rc:
            raise ValueError("bad lap_type for jacobi")
        return v, w

    def restrict3D(self, fine, g, ll=None, bc=(1, 1, 1)):
        """fine ghosted block; boundaries are traded first as the reference
        does (GridFuncVector::restrict3D)."""
        fine = np.array(fine, order="C")
        nf = fine.shape[0]
        dims = tuple(s - 2 * g for s in fine.shape[1:])
        s = _sfx(fine.dtype)
        getattr(self.lib, "orc_trade_boundaries" + s)(
            _c_int3(*dims), g, _c_int3(*bc), _ptr(fine), nf)
        out = np.zeros((nf,) + tuple(d // 2 + 2 * g for d in dims), fine.dtype)
        getattr(self.lib, "orc_restrict3D" + s)(
            _c_int3(*dims), g, _ptr(fine), _ptr(out), nf)
        return out

    def extend3D(self, coarse, fine, g, ll=None, bc=(1, 1, 1)):
        """coarse ghosted block (traded first), fine ghosted block (updated)."""
        coarse = np.array(coarse, order="C")
        fine = np.array(fine, order="C")
        nf = fine.shape[0]
        fdims = tuple(s - 2 * g for s in fine.shape[1:])
        cdims = tuple(d // 2 for d in fdims)
        s = _sfx(fine.dtype)
        getattr(self.lib, "orc_trade_boundaries" + s)(
            _c_int3(*cdims), g, _c_int3(*bc), _ptr(coarse), nf)
        getattr(self.lib, "orc_extend3D" + s)(
            _c_int3(*fdims), g, _ptr(coarse), _ptr(fine), nf)
        return fine

    def precond_mg(self, lap_type, mg_levels, res, ll, gamma, bc=(1, 1, 1)):
        res = np.array(res, order="C")
        nf, nx, ny, nz = res.shape
        rc = self.lib.orc_precond_mg(
            lap_type, mg_levels, _dt(res.dtype), _c_int3(nx, ny, nz),
            _c_dbl3(*ll), _c_int3(*bc), _ptr(res), nf, ctypes.c_double(gamma))
        if rc:
            raise ValueError("precond_mg: unsupported configuration (rc=%d)" % rc)
        return res


    def app_mask_noghost(self, u, masks, level=0):
        """LocGridOrbitals::applyMask on a no-ghost block (ncolors, nx, ny, nz)."""
        u = np.array(u, order="C")
        nf, nx, ny, nz = u.shape
        st, vo, pool = masks.flat()
        getattr(self.lib, "orc_app_mask_noghost" + _sfx(u.dtype))(
            _c_int3(nx, ny, nz), masks.subdivx, masks.ncolors, masks.op,
            _ptr(st[level]), _ptr(vo[level]), _ptr(pool), _ptr(u),
            ctypes.c_size_t(nx * ny * nz), nf)
        return u

    def app_mask_ghosted(self, u, g, masks, level=0):
        """GridFuncVector::app_mask(level) on a ghosted block."""
        u = np.array(u, order="C")
        nf = u.shape[0]
        dims = _c_int3(*(x - 2 * g for x in u.shape[1:]))
        st, vo, pool = masks.flat()
        getattr(self.lib, "orc_app_mask" + _sfx(u.dtype))(
            dims, g, masks.subdivx, masks.ncolors, masks.op, _ptr(st[level]),
            _ptr(vo[level]), _ptr(pool), _ptr(u), nf)
        return u

    def precond_mg_masked(self, lap_type, mg_levels, res, ll, gamma, masks, bc=(1, 1, 1)):
        res = np.array(res, order="C")
        nf, nx, ny, nz = res.shape
        st, vo, pool = masks.flat()
        assert masks.mg_levels >= mg_levels and masks.ncolors == nf
        rc = self.lib.orc_precond_mg_masked(
            lap_type, mg_levels, _dt(res.dtype), _c_int3(nx, ny, nz),
            _c_dbl3(*ll), _c_int3(*bc), _ptr(res), nf, ctypes.c_double(gamma),
            masks.subdivx, masks.ncolors, masks.op, _ptr(st), _ptr(vo), _ptr(pool))
        if rc:
            raise ValueError("precond_mg: unsupported configuration (rc=%d)" % rc)
        return res

    def gamma(self, inv_diag, mg_levels, vmax, small_eig):
        return self.lib.orc_gamma(
            ctypes.c_double(inv_diag), mg_levels, ctypes.c_double(vmax),
            ctypes.c_double(small_eig))

    # -- Poisson solvers of the Hartree potential ----------------
    def poisson_solve(self, lap_type, vh, rho, ll, bc=(1, 1, 1), nu1=2, nu2=2, max_sweeps=10,
                      tol=1e-16, max_nlevels=10):
        """orc_poisson_mg: SolverLap::solve = Mgm + average0.  Returns (solution,
        converged, (nb_sweeps, final_residual, final_relative_residual,
        residual_reduction))."""
        vh = np.array(vh, order="C")
        rho = np.ascontiguousarray(rho, dtype=vh.dtype)
        stats = (ctypes.c_double * 4)()
        conv = getattr(self.lib, "orc_poisson_mg" + _sfx(vh.dtype))(
            lap_type, _c_int3(*vh.shape), _c_dbl3(*ll), _c_int3(*bc), _ptr(vh), _ptr(rho), nu1,
            nu2, max_sweeps, ctypes.c_double(tol), max_nlevels, stats)
        if conv < 0:
            raise ValueError("operator %d is not restated" % lap_type)
        return vh, bool(conv), tuple(stats)

    def pcg_solve(self, lap_type, vh, rho, ll, bc=(1, 1, 1), nu1=2, nu2=2, max_sweeps=10,
                  tol=1e-16, max_nlevels=10):
        """orc_poisson_pcg: PCGSolver::solve.  Returns (solution, converged,
        (final_residual, residual_reduction))."""
        vh = np.array(vh, order="C")
        rho = np.ascontiguousarray(rho, dtype=vh.dtype)
        stats = (ctypes.c_double * 2)()
        conv = getattr(self.lib, "orc_poisson_pcg" + _sfx(vh.dtype))(
            lap_type, _c_int3(*vh.shape), _c_dbl3(*ll), _c_int3(*bc), _ptr(vh), _ptr(rho), nu1,
            nu2, max_sweeps, ctypes.c_double(tol), max_nlevels, stats)
        if conv < 0:
            raise ValueError("operator %d is not restated" % lap_type)
        return vh, bool(conv), tuple(stats)

    # -- contractions ---------------------------------------------------------
    # -- non-local Kleinman-Bylander projectors (row f3; pinned against RefKB below and
    #    tests/golden/reference_kb.npz) --------------------------------------------
    @staticmethod
    def _kb_pack(ions, dtype):
        node0, row0, idx, vals, coeff = [0], [0], [], [], []
        for ion in ions:
            n = len(ion["nlindex"])
            pr = np.ascontiguousarray(ion["proj"], dtype=dtype).reshape(len(ion["coeff"]), n)
            node0.append(node0[-1] + n)
            row0.append(row0[-1] + pr.shape[0])
            idx.append(np.asarray(ion["nlindex"], np.int32))
            vals.append(pr.reshape(-1))
            coeff.append(np.asarray(ion["coeff"], np.float64))
        cat = lambda a, dt: (np.ascontiguousarray(np.concatenate(a), dtype=dt) if a  # noqa: E731
                             else np.zeros(0, dt))
        return (np.asarray(node0, np.int64), np.asarray(row0, np.int32), cat(idx, np.int32),
                cat(vals, dtype), cat(coeff, np.float64))

    def kb_psi(self, ions, psi, vel):
        """kbpsi[row, f] = vel <beta_row | psi_f> (KBPsiMatrixSparse::computeKBpsi)."""
        psi = np.ascontiguousarray(psi)
        nf = psi.shape[0]
        npt = int(np.prod(psi.shape[1:]))
        node0, row0, idx, vals, _ = self._kb_pack(ions, psi.dtype)
        out = np.zeros((int(row0[-1]), nf), np.float64)
        getattr(self.lib, "orc_kb_psi" + _sfx(psi.dtype))(
            len(ions), _ptr(node0), _ptr(row0), _ptr(idx), _ptr(vals), ctypes.c_double(vel),
            _ptr(psi), ctypes.c_size_t(npt), nf, _ptr(out))
        return out

    def kb_vnlpsi(self, ions, kbpsi, out, add):
        """get_vnlpsi for every function; add: out += vnlpsi (computeHnlPhiAndAdd2HPhi)."""
        out = np.array(out, order="C")
        nf = out.shape[0]
        npt = int(np.prod(out.shape[1:]))
        node0, row0, idx, vals, coeff = self._kb_pack(ions, out.dtype)
        kbpsi = np.ascontiguousarray(kbpsi, np.float64)
        getattr(self.lib, "orc_kb_vnlpsi" + _sfx(out.dtype))(
            len(ions), _ptr(node0), _ptr(row0), _ptr(idx), _ptr(vals), _ptr(coeff), _ptr(kbpsi),
            ctypes.c_size_t(npt), _ptr(out), ctypes.c_size_t(npt), nf, int(bool(add)))
        return out

    def gemm_tn(self, a, b, alpha=1.0):
        """alpha * A^T B for blocks a (m, npt), b (n, npt) -> (m, n) double,
        returned as C[i, j] = alpha * <a_i, b_j>."""
        a2 = np.ascontiguousarray(a).reshape(a.shape[0], -1)
        b2 = np.ascontiguousarray(b).reshape(b.shape[0], -1)
        m, k = a2.shape
        n = b2.shape[0]
        c = np.zeros((n, m), np.float64)  # column-major m x n
        getattr(self.lib, "orc_gemm_tn" + _sfx(a2.dtype))(
            m, n, k, ctypes.c_double(alpha), _ptr(a2), k, _ptr(b2), k,
            ctypes.c_double(0.0), _ptr(c), m)
        return np.ascontiguousarray(c.T)

    def gemm_nn(self, phi, mat, alpha=1.0):
        """Phi * M: phi (N, npt...) , mat[l, j] (N x n) -> (n, npt...)."""
        shape = phi.shape
        p2 = np.ascontiguousarray(phi).reshape(shape[0], -1)
        k, m = p2.shape
        matf = np.asfortranarray(mat, dtype=np.float64)
        n = matf.shape[1]
        out = np.zeros((n, m), p2.dtype)
        getattr(self.lib, "orc_gemm_nn" + _sfx(p2.dtype))(
            m, n, k, ctypes.c_double(alpha), _ptr(p2), m,
            matf.ctypes.data_as(ctypes.c_void_p), k, ctypes.c_double(0.0),
            _ptr(out), m)
        return out.reshape((n,) + shape[1:])


class Ref:
    """The reference's own compiled sources (kind "reference")."""

    kind = "reference"

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(
                REF_SO + " missing: run `make -C oracle ref` where "
                "/root/reference exists")
        self.lib = ctypes.CDLL(REF_SO)
        self.lib.ref_mpdot.restype = ctypes.c_double
        self._port = None

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def trade_boundaries(self, noghost, g, bc=(1, 1, 1), ll=(1.0, 1.0, 1.0)):
        a = np.ascontiguousarray(noghost)
        nf, nx, ny, nz = a.shape
        out = np.empty((nf, nx + 2 * g, ny + 2 * g, nz + 2 * g), a.dtype)
        self.lib.ref_trade_boundaries(
            _dt(a.dtype), _c_int3(nx, ny, nz), g, _c_dbl3(*ll), _c_int3(*bc),
            _ptr(a), _ptr(out), nf)
        return out

    def fdkernel(self, kind, v, g, h, rhs_ghosts=None):
        v = np.ascontiguousarray(v)
        nf = v.shape[0]
        dims = tuple(s - 2 * g for s in v.shape[1:])
        ll = _c_dbl3(*(hh * d for hh, d in zip(h, dims)))
        rg = g if rhs_ghosts is None else rhs_ghosts
        if kind == 100:
            out = np.zeros((nf,) + tuple(d + 2 * rg for d in dims), v.dtype)
        else:
            out = np.zeros_like(v)
        self.lib.ref_fdkernel(
            kind, _dt(v.dtype), _c_int3(*dims), g, ll, _ptr(v), _ptr(out), nf, rg)
        return out

    def hpsi(self, lap_type, phi, vtot, ll, bc=(1, 1, 1)):
        phi = np.ascontiguousarray(phi)
        vtot = np.ascontiguousarray(vtot, dtype=np.float64)
        nf, nx, ny, nz = phi.shape
        out = np.empty_like(phi)
        self.lib.ref_hpsi(
            lap_type, _dt(phi.dtype), _c_int3(nx, ny, nz), _c_dbl3(*ll),
            _c_int3(*bc), _ptr(phi), _ptr(vtot), _ptr(out), nf)
        return out

    def lap_constants(self, lap_type, dims, g, ll):
        out = _c_dbl3()
        self.lib.ref_lap_constants(
            lap_type, _c_int3(*dims), g, _c_dbl3(*ll), out)
        return tuple(out)

    def jacobi(self, lap_type, v, f, w, g, ll, jacobi_factor, bc=(1, 1, 1)):
        v = np.array(v, dtype=np.float32, order="C")
        w = np.array(w, dtype=np.float32, order="C")
        f = np.ascontiguousarray(f, dtype=np.float32)
        nf = v.shape[0]
        dims = tuple(s - 2 * g for s in v.shape[1:])
        self.lib.ref_jacobi(
            lap_type, _c_int3(*dims), g, _c_dbl3(*ll), _c_int3(*bc), _ptr(v),
            _ptr(f), _ptr(w), nf, ctypes.c_double(jacobi_factor))
        return v, w

    def restrict3D(self, fine, g, ll=(1.0, 1.0, 1.0), bc=(1, 1, 1)):
        fine = np.ascontiguousarray(fine)
        nf = fine.shape[0]
        dims = tuple(s - 2 * g for s in fine.shape[1:])
        out = np.zeros((nf,) + tuple(d // 2 + 2 * g for d in dims), fine.dtype)
        self.lib.ref_restrict3D(
            _dt(fine.dtype), _c_int3(*dims), g, _c_dbl3(*ll), _c_int3(*bc),
            _ptr(fine), _ptr(out), nf)
        return out

    def extend3D(self, coarse, fine, g, ll=(1.0, 1.0, 1.0), bc=(1, 1, 1)):
        coarse = np.ascontiguousarray(coarse)
        fine = np.array(fine, order="C")
        nf = fine.shape[0]
        dims = tuple(s - 2 * g for s in fine.shape[1:])
        self.lib.ref_extend3D(
            _dt(fine.dtype), _c_int3(*dims), g, _c_dbl3(*ll), _c_int3(*bc),
            _ptr(coarse), _ptr(fine), nf)
        return fine

    def precond_mg(self, lap_type, mg_levels, res, ll, gamma, bc=(1, 1, 1)):
        res = np.array(res, order="C")
        nf, nx, ny, nz = res.shape
        self.lib.ref_precond_mg(
            lap_type, mg_levels, _dt(res.dtype), _c_int3(nx, ny, nz),
            _c_dbl3(*ll), _c_int3(*bc), _ptr(res), nf, ctypes.c_double(gamma))
        return res


    def lap_rhs(self, lap_type, phi, ll, bc=(1, 1, 1)):
        """pb::Lap::rhs(GridFunc&, T*) per orbital (Laph4MP asserts > 1 ghost)."""
        phi = np.ascontiguousarray(phi)
        nf, nx, ny, nz = phi.shape
        out = np.empty_like(phi)
        self.lib.ref_lap_rhs(lap_type, _dt(phi.dtype), _c_int3(nx, ny, nz), 2, _c_dbl3(*ll),
                             _c_int3(*bc), _ptr(phi), _ptr(out), nf)
        return out

    # -- AndersonMix<Solution> of the reference -------------------------------------
    def poisson_solve(self, lap_type, vh, rho, ll, bc=(1, 1, 1), nu1=2, nu2=2, max_sweeps=10,
                      tol=1e-16, max_nlevels=10):
        """pb::SolverLap<Lap<T>, T>::solve (src/pb/SolverLap.cc:62-72): Mgm +
        average0.  vh: initial guess (no ghosts); returns (solution, converged,
        (nb_sweeps, final_residual, final_relative_residual, residual_reduction))."""
        vh = np.array(vh, order="C")
        rho = np.ascontiguousarray(rho, dtype=vh.dtype)
        stats = (ctypes.c_double * 4)()
        conv = self.lib.ref_poisson_solve(
            lap_type, _dt(vh.dtype), _c_int3(*vh.shape), _c_dbl3(*ll), _c_int3(*bc), _ptr(vh),
            _ptr(rho), nu1, nu2, max_sweeps, ctypes.c_double(tol), max_nlevels, stats)
        if conv < 0:
            raise ValueError("operator %d not wired in the reference shim" % lap_type)
        return vh, bool(conv), tuple(stats)

    def pcg_solve(self, lap_type, vh, rho, ll, bc=(1, 1, 1), nu1=2, nu2=2, max_sweeps=10,
                  tol=1e-16, max_nlevels=10):
        """PCGSolver<Lap<T>, T>::solve (src/PCGSolver.cc:165-252), Control::lap_type
        = lap_type.  Returns (solution, converged, (final_residual, residual_reduction))."""
        vh = np.array(vh, order="C")
        rho = np.ascontiguousarray(rho, dtype=vh.dtype)
        stats = (ctypes.c_double * 2)()
        conv = self.lib.ref_pcg_solve(
            lap_type, _dt(vh.dtype), _c_int3(*vh.shape), _c_dbl3(*ll), _c_int3(*bc), _ptr(vh),
            _ptr(rho), nu1, nu2, max_sweeps, ctypes.c_double(tol), max_nlevels, stats)
        if conv < 0:
            raise ValueError("operator %d not wired in the reference shim" % lap_type)
        return vh, bool(conv), tuple(stats)

    def anderson_create(self, m, beta, x0):
        self.lib.ref_anderson_create.restype = ctypes.c_void_p
        x0 = np.ascontiguousarray(x0, np.float64)
        return ctypes.c_void_p(self.lib.ref_anderson_create(
            m, ctypes.c_double(beta), len(x0), _ptr(x0)))

    def anderson_update(self, h, f, invs=1.0):
        """One AndersonMix::update; returns (x after the update, mixed f)."""
        f = np.array(f, np.float64)
        x = np.empty_like(f)
        self.lib.ref_anderson_update(h, _ptr(f), len(f), ctypes.c_double(invs), _ptr(x))
        return x, f

    def anderson_destroy(self, h):
        self.lib.ref_anderson_destroy(h)

    # -- localization masks (GridMask / Map2Masks of the reference) ------------
    def masks_create(self, dims, ll, ghosts, mg_levels, subdivx, op, gids, centers, radii):
        """Reference mask objects for `gids` (GridMaskMult op 0 / GridMaskMax
        op 1), values generated by GridMask::init.  Returns an opaque handle."""
        self.lib.ref_masks_create.restype = ctypes.c_void_p
        gids = np.ascontiguousarray(gids, np.int32)
        centers = np.ascontiguousarray(centers, np.float64)
        radii = np.ascontiguousarray(radii, np.float64)
        h = self.lib.ref_masks_create(
            _c_int3(*dims), _c_dbl3(*ll), ghosts, mg_levels, subdivx, op,
            len(gids), _ptr(gids), _ptr(centers), _ptr(radii))
        return (ctypes.c_void_p(h), tuple(dims), mg_levels, subdivx, op)

    def masks_tables(self, handle, gid_table):
        """MaskTables (plain data) of a reference mask set for a
        (subdivx, ncolors) gid table."""
        h, dims, mg_levels, subdivx, op = handle
        gid_table = np.asarray(gid_table, np.int64)
        ncolors = gid_table.shape[1]
        mt = MaskTables(dims, mg_levels, subdivx, ncolors, op)
        for l in range(mg_levels + 1):
            n = self.lib.ref_masks_loc_numpt(h, l)
            buf = np.empty(n, np.float64)
            for iloc in range(subdivx):
                for c in range(ncolors):
                    gid = int(gid_table[iloc, c])
                    if gid < 0:
                        mt.set(l, iloc, c, 0)
                        continue
                    st = self.lib.ref_masks_state(h, gid, l, iloc)
                    if st == 2:
                        got = self.lib.ref_masks_values(h, gid, l, iloc, _ptr(buf))
                        assert got == n
                        mt.set(l, iloc, c, 2, buf.copy())
                    else:
                        mt.set(l, iloc, c, st)
        return mt

    def app_mask_noghost(self, u, handle, gid_table, level=0):
        u = np.array(u, order="C")
        nf, nx, ny, nz = u.shape
        gt = np.ascontiguousarray(gid_table, np.int32)
        self.lib.ref_app_mask_noghost(
            handle[0], _dt(u.dtype), level, nf, _ptr(gt), _ptr(u),
            ctypes.c_long(nx * ny * nz))
        return u

    def precond_mg_masked(self, lap_type, res, gamma, handle, gid_table, bc=(1, 1, 1)):
        """precond_mg with the reference's Map2Masks set; grid, levels and
        ghosts are those of the mask handle."""
        res = np.array(res, order="C")
        nf = res.shape[0]
        gt = np.ascontiguousarray(gid_table, np.int32)
        assert gt.shape == (handle[3], nf)
        self.lib.ref_precond_mg_masked(
            handle[0], lap_type, _dt(res.dtype), _c_int3(*bc), nf, _ptr(gt),
            _ptr(res), ctypes.c_double(gamma))
        return res

    def gemm_tn(self, a, b, alpha=1.0):
        """MPgemm('T','N') (mputils.cc): double in -> DGEMM, float in ->
        double-accumulating loops; result matrix in double."""
        a2 = np.ascontiguousarray(a).reshape(a.shape[0], -1)
        b2 = np.ascontiguousarray(b).reshape(b.shape[0], -1)
        m, k = a2.shape
        n = b2.shape[0]
        c = np.zeros((n, m), np.float64)
        self.lib.ref_mpgemm(
            _dt(a2.dtype), 1, ctypes.c_char(b"T"), ctypes.c_char(b"N"), m, n, k,
            ctypes.c_double(alpha), _ptr(a2), k, _ptr(b2), k,
            ctypes.c_double(0.0), _ptr(c), m)
        return np.ascontiguousarray(c.T)

    def syrk(self, a, alpha=1.0):
        """MPsyrk('l','t') as LocalMatrices::syrk calls it
        (local_matrices/LocalMatrices.cc:210-247); lower triangle mirrored."""
        a2 = np.ascontiguousarray(a).reshape(a.shape[0], -1)
        n, k = a2.shape
        c = np.zeros((n, n), np.float64)
        self.lib.ref_mpsyrk(
            _dt(a2.dtype), 1, ctypes.c_char(b"l"), ctypes.c_char(b"t"), n, k,
            ctypes.c_double(alpha), _ptr(a2), k, ctypes.c_double(0.0), _ptr(c), n)
        # column-major lower == row-major upper
        full = np.triu(c) + np.triu(c, 1).T
        return full

    def gemm_nn(self, phi, mat, alpha=1.0):
        shape = phi.shape
        p2 = np.ascontiguousarray(phi).reshape(shape[0], -1)
        k, m = p2.shape
        matf = np.asfortranarray(mat, dtype=np.float64)
        n = matf.shape[1]
        out = np.zeros((n, m), p2.dtype)
        self.lib.ref_mpgemmNN_phiM(
            _dt(p2.dtype), m, n, k, ctypes.c_double(alpha), _ptr(p2), m,
            matf.ctypes.data_as(ctypes.c_void_p), k, ctypes.c_double(0.0),
            _ptr(out), m)
        return out.reshape((n,) + shape[1:])


# ---------------------------------------------------------------------------
# Deterministic synthetic inputs
# ---------------------------------------------------------------------------
from mgmol_b200.synthetic import synthetic_kb_projectors  # noqa: E402,F401  (numpy-only input generator)


H2O512_CELL = 46.9768  # bohr, cubic cell of examples/H2O_512 (h = 0.1835 @256)


def synthetic_orbitals(nfunc, dims, dtype=np.float64, first=0, noise=0.1):
    """psi_j = sum of 3 plane waves with orbital-dependent integer wave
    vectors + noise * U(-1, 1); rng seed 1234 + j."""
    nx, ny, nz = dims
    x = np.arange(nx)[:, None, None] / nx
    y = np.arange(ny)[None, :, None] / ny
    z = np.arange(nz)[None, None, :] / nz
    out = np.empty((nfunc, nx, ny, nz), dtype)
    for j in range(nfunc):
        gj = first + j
        rng = np.random.default_rng(1234 + gj)
        k = rng.integers(-3, 4, size=(3, 3))
        amp = rng.uniform(0.3, 1.0, size=3)
        psi = np.zeros((nx, ny, nz))
        for w in range(3):
            psi += amp[w] * np.cos(
                2 * np.pi * (k[w, 0] * x + k[w, 1] * y + k[w, 2] * z) + 0.3 * w)
        psi += noise * rng.uniform(-1.0, 1.0, size=(nx, ny, nz))
        out[j] = psi.astype(dtype)
    return out


def synthetic_potential(dims, noise=0.05):
    """Smooth -2 .. +0.5 Ry field (cos mixture) + seeded noise."""
    nx, ny, nz = dims
    x = np.arange(nx)[:, None, None] / nx
    y = np.arange(ny)[None, :, None] / ny
    z = np.arange(nz)[None, None, :] / nz
    v = (-0.75 + 0.6 * np.cos(2 * np.pi * x) * np.cos(2 * np.pi * y)
         + 0.4 * np.cos(4 * np.pi * z) + 0.25 * np.cos(2 * np.pi * (x + y + z)))
    rng = np.random.default_rng(99)
    return np.ascontiguousarray(v + noise * rng.uniform(-1, 1, size=(nx, ny, nz)))


class RefKB:
    """The reference's own sparse Kleinman-Bylander projector code (KBprojectorSparse.cc,
    Species.cc, Mesh.cc, radial/*.cc compiled unmodified, oracle/ref_shim_kb.cc) for one
    ORBDTYPE: real pseudopotentials from the reference's potentials/ directory, the
    projector of an ion as KBprojectorSparse::setup builds it, <beta|psi> as
    computeLocalElement forms it and get_vnlpsi's loop.  Only where /root/reference
    exists (the pseudopotential files are read from there); the GPU box uses the golden
    vectors generated from it (tests/golden/make_golden_kb.py)."""

    POTENTIALS = "/root/reference/potentials"

    def __init__(self, dtype):
        self.dtype = np.dtype(dtype)
        so = os.path.join(HERE, "_ref", "libmgmol_refkb_f%d.so" % (8 * self.dtype.itemsize))
        if not os.path.exists(so):
            raise FileNotFoundError(so + " missing: run `make -C oracle ref`")
        self.lib = ctypes.CDLL(so, mode=getattr(os, "RTLD_LOCAL", 0))
        assert self.lib.refkb_orbdtype_bytes() == self.dtype.itemsize
        self.nions = 0

    @staticmethod
    def available(dtype=np.float64):
        so = os.path.join(HERE, "_ref", "libmgmol_refkb_f%d.so" % (8 * np.dtype(dtype).itemsize))
        return os.path.exists(so) and os.path.isdir(RefKB.POTENTIALS)

    def setup(self, dims, ll, lap_type, pseudo, filter_flag="n", origin=(0.0, 0.0, 0.0)):
        """Mesh + species.  Returns dict(nlradius, dim_nl, max_l, llocal, nproj)."""
        info = (ctypes.c_double * 5)()
        path = pseudo if os.path.isabs(pseudo) else os.path.join(self.POTENTIALS, pseudo)
        rc = self.lib.refkb_setup(_c_int3(*dims), _c_dbl3(*origin), _c_dbl3(*ll), lap_type,
                                  path.encode(), ctypes.c_char(filter_flag.encode()), info)
        if rc:
            raise FileNotFoundError(path)
        self.dims, self.nions = tuple(dims), 0
        return dict(zip(("nlradius", "dim_nl", "max_l", "llocal", "nproj"), list(info)))

    def add_ion(self, center):
        """KBprojectorSparse(species).setup(center); returns the ion as the dict the
        product and the port take (nlindex, proj (nproj, size_nl), coeff = kbcoeff * sign)."""
        j = self.lib.refkb_add_ion(_c_dbl3(*center))
        assert j == self.nions
        self.nions += 1
        n, npj = self.lib.refkb_ion_size(j), self.lib.refkb_ion_nproj(j)
        idx = np.zeros(n, np.int32)
        proj = np.zeros((npj, n), self.dtype)
        coeff = np.zeros(npj, np.float64)
        rc = self.lib.refkb_ion_data(j, _ptr(idx), _ptr(proj), _ptr(coeff))
        assert rc == 0, rc
        return {"nlindex": idx, "proj": proj, "coeff": coeff,
                "single": bool(self.lib.refkb_ion_single(j))}

    def kb_psi(self, psi):
        """kbpsi[row, f] for every ion added so far (rows ion after ion)."""
        psi = np.ascontiguousarray(psi, self.dtype)
        rows = []
        for j in range(self.nions):
            out = np.zeros((self.lib.refkb_ion_nproj(j), psi.shape[0]), np.float64)
            col = np.zeros(out.shape[0], np.float64)
            for f in range(psi.shape[0]):
                self.lib.refkb_psi(j, _ptr(psi[f]), _ptr(col))
                out[:, f] = col
            rows.append(out)
        return np.concatenate(rows, axis=0)

    def kb_vnlpsi(self, kbpsi, out, add):
        out = np.array(out, dtype=self.dtype, order="C")
        kbpsi = np.ascontiguousarray(kbpsi, np.float64)
        for f in range(out.shape[0]):
            rows = np.ascontiguousarray(kbpsi[:, f])
            n = self.lib.refkb_vnlpsi(_ptr(rows), _ptr(out[f]), int(bool(add)))
            assert n == kbpsi.shape[0]
        return out
