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

    def gamma(self, inv_diag, mg_levels, vmax, small_eig):
        return self.lib.orc_gamma(
            ctypes.c_double(inv_diag), mg_levels, ctypes.c_double(vmax),
            ctypes.c_double(small_eig))

    # -- contractions ---------------------------------------------------------
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
