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
