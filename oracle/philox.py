"""Test infrastructure: numpy restatement of the counter-based generator of nn-fac_b200/csrc/synth.cu
(Philox4x32-10, Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; counter = (row, column, stream, 0),
key = (seed low, seed high), first output word -> 24-bit uniform in [0, 1)).  Only tests/ and bench.py's CPU legs may
import this module; the product never does."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def uniform(rows, cols, row0=0, col0=0, seed=0, stream_id=0, scale=1.0):
    """float32 [rows x cols]: element (i, j) = scale * u(seed, stream_id, row0 + i, col0 + j)."""
    r = (np.arange(rows, dtype=np.uint64) + np.uint64(row0))[:, None]
    c = (np.arange(cols, dtype=np.uint64) + np.uint64(col0))[None, :]
    c0 = np.broadcast_to(r, (rows, cols)).copy() & MASK
    c1 = np.broadcast_to(c, (rows, cols)).copy() & MASK
    c2 = np.full((rows, cols), stream_id, dtype=np.uint64)
    c3 = np.zeros((rows, cols), dtype=np.uint64)
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2                      # 32 x 32 -> 64-bit products, exact in uint64
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    u = (c0 >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return (np.float32(scale) * u).astype(np.float32)
