"""Seeded, block-addressable synthetic affinities (test / bench harness, not the timed path).

Jittered-grid Voronoi cells, changed in one respect: the membrane profile is
the rational 1/(1+(Δ/2.5)^4) instead of exp(-(Δ/3)^2), so that every operation is an IEEE
+,-,*,/,sqrt in float64 and the numpy generator here and the CUDA generator
(csrc/synth.cu, `bs_synth_affs`) agree bit for bit.

  site(cell)  = pitch * (cell + 0.25 + 0.5 * h3(seed, cell))          pitch = (12, 48, 48)
  dist2(p, s) = (4*(pz-sz))^2 + (py-sy)^2 + (px-sx)^2                 over the 27 neighbouring cells
  lab(p)      = cell of the nearest site;   Δ = d2 - d1;   b(p) = 1/(1+(Δ/2.5)^4)
  a_c(p)      = link between p and q = p - e_c :  m = max(b(p), b(q))
                same cell ? 1-m : 0.05*(1-m);  + 0.08*(u-0.5);  clamp [0,1];  0 if q outside
  uint8 = rint(255*a)   float32 = float(a)
"""
import numpy as np

PITCH = (12, 48, 48)
_G = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = x + _G
        z = x
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def _hash(seed, *words):
    x = np.uint64(seed)
    with np.errstate(over="ignore"):
        for w in words:
            w = np.asarray(w).astype(np.int64).astype(np.uint64)
            x = _splitmix64(x ^ (w * _G))
    return x


def block_seed(seed, block_id):
    """noise seed of one block of a blockwise task: SplitMix64 chain over (seed, block id, 29)"""
    return int(_hash(seed, np.int64(block_id), np.int64(29)))


def _uniform(h):
    return (h >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def _cells(seed, z, y, x):
    """labels (int64 packed cell id) and membrane b for voxel coordinate arrays."""
    z = np.asarray(z, np.int64)
    y = np.asarray(y, np.int64)
    x = np.asarray(x, np.int64)
    cz, cy, cx = np.floor_divide(z, PITCH[0]), np.floor_divide(y, PITCH[1]), np.floor_divide(x, PITCH[2])
    shape = np.broadcast(z, y, x).shape
    d1 = np.full(shape, np.inf)
    d2 = np.full(shape, np.inf)
    lab = np.zeros(shape, np.int64)
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                iz, iy, ix = cz + dz, cy + dy, cx + dx
                sz = PITCH[0] * (iz + (0.25 + 0.5 * _uniform(_hash(seed, iz, iy, ix, 0))))
                sy = PITCH[1] * (iy + (0.25 + 0.5 * _uniform(_hash(seed, iz, iy, ix, 1))))
                sx = PITCH[2] * (ix + (0.25 + 0.5 * _uniform(_hash(seed, iz, iy, ix, 2))))
                ez = 4.0 * (z - sz)
                ey = y - sy
                ex = x - sx
                dd = (ez * ez + ey * ey) + ex * ex
                cid = ((iz + 1024) * 4096 + (iy + 1024)) * 4096 + (ix + 1024)
                closer = dd < d1
                d2 = np.where(closer, d1, np.minimum(d2, dd))
                lab = np.where(closer, cid, lab)
                d1 = np.where(closer, dd, d1)
    delta = np.sqrt(d2) - np.sqrt(d1)
    t = delta / 2.5
    t2 = t * t
    b = 1.0 / (1.0 + t2 * t2)
    return lab, b


def synth_affs(shape, seed=0, dtype=np.uint8, offset=(0, 0, 0), vol_shape=None, n_channels=3):
    """Affinities (3, Z, Y, X) for the region [offset, offset+shape) of a volume vol_shape."""
    assert n_channels == 3
    vol_shape = tuple(vol_shape) if vol_shape is not None else tuple(o + s for o, s in zip(offset, shape))
    # one extra voxel on the low side of every axis for q = p - e_c
    zz = np.arange(offset[0] - 1, offset[0] + shape[0])[:, None, None]
    yy = np.arange(offset[1] - 1, offset[1] + shape[1])[None, :, None]
    xx = np.arange(offset[2] - 1, offset[2] + shape[2])[None, None, :]
    lab, b = _cells(seed, zz, yy, xx)
    lab = np.broadcast_to(lab, (shape[0] + 1, shape[1] + 1, shape[2] + 1))
    b = np.broadcast_to(b, lab.shape)
    z, y, x = zz[1:], yy[:, 1:], xx[:, :, 1:]
    out = np.empty((3,) + tuple(shape), dtype=dtype)
    lp, bp = lab[1:, 1:, 1:], b[1:, 1:, 1:]
    sl = [(slice(0, -1), slice(1, None), slice(1, None)),
          (slice(1, None), slice(0, -1), slice(1, None)),
          (slice(1, None), slice(1, None), slice(0, -1))]
    coords = (z, y, x)
    for c in range(3):
        lq, bq = lab[sl[c]], b[sl[c]]
        m = np.maximum(bp, bq)
        a = np.where(lp == lq, 1.0 - m, 0.05 * (1.0 - m))
        u = _uniform(_hash(seed, c, z, y, x, 7))
        a = a + 0.08 * (u - 0.5)
        a = np.minimum(np.maximum(a, 0.0), 1.0)
        inside = np.broadcast_to(coords[c] - 1 >= 0, a.shape)
        a = np.where(inside, a, 0.0)
        if np.dtype(dtype) == np.uint8:
            out[c] = np.rint(255.0 * a).astype(np.uint8)
        else:
            out[c] = a.astype(np.float32)
    assert all(o + s <= v for o, s, v in zip(offset, shape, vol_shape))
    return out
