"""numpy restatement of the library's Philox4x32-10 stream contract (csrc/common.cuh):

  uniform(draw, row, unit) = (philox(ctr=(unit, row>>2, draw_lo, draw_hi), key=seed)[row & 3] >> 9) * 2^-23
  normal (draw, row, unit) = box_muller(philox(ctr=(unit, row>>2, draw_lo, draw_hi | 2^31), key=seed))[row & 3]
      box_muller(x, y, z, w) = (ra cos ta, ra sin ta, rb cos tb, rb sin tb),  r = sqrt(-2 ln(1 - (x>>9) 2^-23)),  t = 2 pi (y>>9) 2^-23
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _row_index(rows, row0):
    """`rows` is a count (rows row0 .. row0 + rows - 1) or an explicit array of global row indices."""
    if np.ndim(rows) == 0:
        return (np.arange(int(rows), dtype=np.uint64) + np.uint64(row0))[:, None]
    return (np.asarray(rows, dtype=np.uint64) + np.uint64(row0))[:, None]


def uniforms(seed: int, draw: int, rows, units: int, row0: int = 0) -> np.ndarray:
    """[rows][units] float64 array of the uniforms the device uses for one sampling half-step."""
    r = _row_index(rows, row0)
    u = np.arange(units, dtype=np.uint64)[None, :]
    r, u = np.broadcast_arrays(r, u)
    out = philox4x32_10(u, r >> np.uint64(2), draw & 0xFFFFFFFF, (draw >> 32) & 0x7FFFFFFF, seed & 0xFFFFFFFF, seed >> 32)
    sel = (r & np.uint64(3)).astype(np.int64)
    x = np.choose(sel, out)
    return (x >> np.uint64(9)).astype(np.float64) / float(1 << 23)


def normals(seed: int, draw: int, rows, units: int, row0: int = 0) -> np.ndarray:
    """[rows][units] standard normals of one Gaussian sampling half-step, in Float64 from the exact 23-bit uniforms (the
    device evaluates the same formula in fp32, with approximate lg2 / sqrt / sin / cos in the tensor-core epilogue:
    agreement to ~1e-5 absolute, not bit for bit)."""
    r = _row_index(rows, row0)
    u = np.arange(units, dtype=np.uint64)[None, :]
    r, u = np.broadcast_arrays(r, u)
    out = philox4x32_10(u, r >> np.uint64(2), draw & 0xFFFFFFFF, ((draw >> 32) & 0x7FFFFFFF) | 0x80000000,
                        seed & 0xFFFFFFFF, seed >> 32)
    sel = (r & np.uint64(3)).astype(np.int64)
    hi = sel >= 2                                  # rows 4g+2, 4g+3 use words (z, w); rows 4g, 4g+1 words (x, y)
    a = np.where(hi, out[2], out[0])
    b = np.where(hi, out[3], out[1])
    u1 = 1.0 - (a >> np.uint64(9)).astype(np.float64) / float(1 << 23)
    u2 = (b >> np.uint64(9)).astype(np.float64) / float(1 << 23)
    rad = np.sqrt(-2.0 * np.log(u1))
    return np.where((sel & 1) == 1, rad * np.sin(2.0 * np.pi * u2), rad * np.cos(2.0 * np.pi * u2))
