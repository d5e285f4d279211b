"""Deterministic synthetic inputs shared by the tests, the CPU checkers and bench.py
(numpy only; no device code, no reference code)."""
import numpy as np


def synthetic_kb_projectors(dims, ll, nions, radius, dtype=np.float64, seed=5):
    """Sparse projector vectors shaped like KBprojectorSparse's: per ion a ball of nodes
    around a random centre (periodic wrap, so neighbouring balls overlap) and 1 (s only) or
    4 (s + three p) value arrays over it: a radial profile times 1, x, y, z.  coeff =
    kbcoeff * sign.  Ions: list of dicts nlindex / proj (nproj, size_nl) / coeff."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = dims
    h = [l / n for l, n in zip(ll, dims)]
    ions = []
    for j in range(nions):
        c = rng.uniform(0, 1, 3) * np.asarray(ll)
        r = [int(np.ceil(radius / h[d])) for d in range(3)]
        c0 = [int(round(c[d] / h[d])) for d in range(3)]
        ix = np.arange(c0[0] - r[0], c0[0] + r[0] + 1)
        iy = np.arange(c0[1] - r[1], c0[1] + r[1] + 1)
        iz = np.arange(c0[2] - r[2], c0[2] + r[2] + 1)
        X, Y, Z = np.meshgrid(ix * h[0] - c[0], iy * h[1] - c[1], iz * h[2] - c[2], indexing="ij")
        R2 = X * X + Y * Y + Z * Z
        inside = R2 <= radius * radius
        IX, IY, IZ = np.meshgrid(ix % nx, iy % ny, iz % nz, indexing="ij")
        idx = ((IX * ny + IY) * nz + IZ)[inside].astype(np.int32)
        idx, first = np.unique(idx, return_index=True)  # a ball wider than the box wraps onto itself
        prof = np.exp(-2.0 * R2[inside][first] / radius ** 2) * (1.0 - R2[inside][first] / radius ** 2)
        x, y, z = X[inside][first], Y[inside][first], Z[inside][first]
        if j % 3 == 0:
            proj = np.stack([prof, prof * x, prof * y, prof * z])
            coeff = np.array([1.7, -0.9, -0.9, -0.9]) * (1.0 + 0.1 * j)
        else:
            proj = prof[None, :]
            coeff = np.array([-1.3 if j % 2 else 2.1])
        ions.append({"nlindex": idx, "proj": proj.astype(dtype), "coeff": coeff})
    return ions
