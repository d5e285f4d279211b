"""A numpy stand-in for the device GridFuncVector (nfunc = 1), every operation
done by the oracle's C restatement of the same reference function the device
entry point is tested against in test_gpu_parity.py.  It lets the CPU suite run
the product's Poisson control flow (mgmol_b200/poisson.py) without a GPU and
compare it with the compiled reference solver.  Test infrastructure only."""
import numpy as np

_KIND = {0: 0, 1: 1, 2: 2}


class TwinField:
    def __init__(self, port, grid, dtype):
        self.port_ = port
        self.grid_ = grid
        self.data = np.zeros((1,) + tuple(grid.shapeg()), dtype)
        self.updated_boundaries_ = False

    # -- the GridFuncVector interface of mgmol_b200/host.py ---------------------
    def grid(self):
        return self.grid_

    def size(self):
        return 1

    def resetData(self):
        self.data[...] = 0
        self.updated_boundaries_ = True

    def set_updated_boundaries(self, flag):
        self.updated_boundaries_ = bool(flag)

    def _sl(self):
        g = self.grid_.ghost_pt()
        nx, ny, nz = self.grid_.shape()
        return (slice(None), slice(g, g + nx), slice(g, g + ny), slice(g, g + nz))

    def scal(self, alpha):
        dt = self.data.dtype
        if alpha == 0.0:
            self.data = np.zeros_like(self.data)
        elif alpha != 1.0:
            self.data = (float(alpha) * self.data.astype(np.float64)).astype(dt)

    def assign(self, noghost):
        self.data[...] = 0
        self.data[self._sl()] = noghost          # converts to this field's precision
        self.updated_boundaries_ = False

    def getValues(self, out):
        out[...] = self.data[self._sl()]
        return out

    def values(self):
        return np.ascontiguousarray(self.data[self._sl()])

    def trade_boundaries(self):
        if self.updated_boundaries_:
            return
        inner = np.ascontiguousarray(self.data[self._sl()])
        self.data = self.port_.trade_boundaries(inner, self.grid_.ghost_pt(), self.grid_.bc)
        self.updated_boundaries_ = True

    def _h(self):
        return tuple(self.grid_.hgrid(d) for d in range(3))

    def applyLap(self, lap_type, rhs):
        self.trade_boundaries()
        out = self.port_.fdkernel(_KIND[lap_type], self.data, self.grid_.ghost_pt(), self._h())
        rhs.data[self._sl()] = out[self._sl()]      # the kernel writes interior points only
        rhs.set_updated_boundaries(False)

    def applyRHS(self, lap_type, rhs):
        assert lap_type == 0
        self.trade_boundaries()
        g = self.grid_.ghost_pt()
        out = self.port_.fdkernel(100, self.data, g, self._h(), g)
        rhs.data[self._sl()] = out[self._sl()]
        rhs.set_updated_boundaries(False)

    def axpy(self, alpha, other):
        dt = self.data.dtype
        self.data = self.data + (float(alpha) * other.data.astype(np.float64)).astype(dt)
        self.updated_boundaries_ = other.updated_boundaries_ and self.updated_boundaries_

    def restrict3D(self, ucoarse):
        self.trade_boundaries()
        ucoarse.data = self.port_.restrict3D(self.data, self.grid_.ghost_pt(), bc=self.grid_.bc)

    def extend3D(self, ucoarse):
        ucoarse.trade_boundaries()
        self.data = self.port_.extend3D(ucoarse.data, self.data, self.grid_.ghost_pt(),
                                        bc=self.grid_.bc)
        self.updated_boundaries_ = False

    # -- the GridFunc operations ---------------------------------------------------
    def copy_from(self, other):
        self.data = other.data.copy()
        self.updated_boundaries_ = other.updated_boundaries_

    def _interior(self, drop):
        g = self.grid_.ghost_pt()
        nx, ny, nz = self.grid_.shape()
        lo = [g + (1 if (drop and self.grid_.bc[d] != 1) else 0) for d in range(3)]
        return self.data[0, lo[0]:g + nx, lo[1]:g + ny, lo[2]:g + nz].astype(np.float64)

    def gdot(self, other):
        return float((self._interior(True) * other._interior(True)).sum())

    def norm2(self):
        return (self.gdot(self) * self.grid_.vel()) ** 0.5

    def get_average(self):
        return float(self._interior(False).sum()) / float(self.grid_.size())

    def add_scalar(self, alpha):
        self.data = self.data + np.asarray(alpha, np.float64).astype(self.data.dtype)


def field_factory(port, dtype=None):
    """field(grid, dtype) as PoissonMG / PoissonPCG expect it; torch dtypes of
    the preconditioner are mapped to numpy."""
    import torch
    tmap = {torch.float32: np.float32, torch.float64: np.float64}
    return lambda grid, dt: TwinField(port, grid, tmap.get(dt, dt))


# ---------------------------------------------------------------------------
# Decomposed boxes (world_size > 1 over gloo): the stand-in for
# mgmol_b200.poisson.DecomposedField.  A decomposed boundary trade leaves on
# every rank the matching piece of the single-rank trade of the whole box (that
# is what the device exchange is tested for in tests/mgpu_worker.py), so the twin
# gathers the box, lets the oracle trade it, and cuts its piece out again.
# ---------------------------------------------------------------------------
class DecomposedTwinField(TwinField):
    def _box(self, ghosts=0):
        gr = self.grid_
        return tuple(slice(c * n, c * n + n + 2 * ghosts) for c, n in zip(gr.coord, gr.dim_))

    def _allreduce(self, a):
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(a, np.float64))
        dist.all_reduce(t)
        return t.numpy()

    def gather(self):
        glob = np.zeros((1,) + tuple(self.grid_.gdim_), np.float64)
        glob[(slice(None),) + self._box()] = self.data[self._sl()]
        return self._allreduce(glob).astype(self.data.dtype)

    def trade_boundaries(self):
        if self.updated_boundaries_:
            return
        g = self.grid_.ghost_pt()
        traded = self.port_.trade_boundaries(self.gather(), g, self.grid_.bc)
        self.data = np.ascontiguousarray(traded[(slice(None),) + self._box(g)])
        self.updated_boundaries_ = True

    def _raw(self, name, *args):
        from oracle.oracle import _sfx
        return getattr(self.port_.lib, name + _sfx(self.data.dtype))(*args)

    def restrict3D(self, ucoarse):
        from oracle.oracle import _c_int3, _ptr
        self.trade_boundaries()
        g = self.grid_.ghost_pt()
        fine = np.ascontiguousarray(self.data)
        out = np.zeros((1,) + tuple(d // 2 + 2 * g for d in self.grid_.shape()), fine.dtype)
        self._raw("orc_restrict3D", _c_int3(*self.grid_.shape()), g, _ptr(fine), _ptr(out), 1)
        ucoarse.data = out

    def extend3D(self, ucoarse):
        from oracle.oracle import _c_int3, _ptr
        ucoarse.trade_boundaries()
        g = self.grid_.ghost_pt()
        coarse = np.ascontiguousarray(ucoarse.data)
        fine = np.ascontiguousarray(self.data)
        self._raw("orc_extend3D", _c_int3(*self.grid_.shape()), g, _ptr(coarse), _ptr(fine), 1)
        self.data = fine
        self.updated_boundaries_ = False

    def _interior(self, drop):
        g = self.grid_.ghost_pt()
        nx, ny, nz = self.grid_.shape()
        lo = [g + (1 if (drop and self.grid_.bc[d] != 1 and self.grid_.coord[d] == 0) else 0)
              for d in range(3)]
        return self.data[0, lo[0]:g + nx, lo[1]:g + ny, lo[2]:g + nz].astype(np.float64)

    def gdot(self, other):
        return float(self._allreduce(np.array([super().gdot(other)]))[0])

    def get_average(self):
        gr = self.grid_
        n = gr.nproc[0] * gr.nproc[1] * gr.nproc[2]
        return float(self._allreduce(np.array([super().get_average()]))[0]) / n


def decomposed_field_factory(port):
    """Single-rank grids (the replicated coarse levels) get the plain twin."""
    import torch
    tmap = {torch.float32: np.float32, torch.float64: np.float64}

    def make(grid, dt):
        cls = DecomposedTwinField if tuple(grid.nproc) != (1, 1, 1) else TwinField
        return cls(port, grid, tmap.get(dt, dt))
    return make
