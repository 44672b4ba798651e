"""Philox4x32-10 (Salmon et al., SC'11; Random123) in vectorised numpy.

Shared random stream of the oracle and the CUDA kernels (csrc/philox.cuh).
numpy's own `Philox` bit generator is the 4x64 variant and is NOT this one.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

# what the 4th counter word means (keeps streams of different uses disjoint)
PURPOSE_NEG = 1        # candidate key of (event, item position)
PURPOSE_NEG_REPL = 2   # j-th draw with replacement of an event
PURPOSE_NBR = 3        # uniform temporal-neighbour slot j of query q
PURPOSE_DROPOUT = 4    # attention-dropout factor of (query, head * n + slot, step)
PURPOSE_NEG_SEQ = 5    # attempt a of slot j of an event's draw without replacement (counter word 2 = j | a << 16)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable uint32-valued arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x).astype(np.uint64) & MASK for x in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0)
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(x.astype(np.uint32) for x in (c0, c1, c2, c3))


def mulhi32(x, n):
    """floor(x * n / 2**32): maps a uint32 draw onto [0, n)."""
    return ((np.asarray(x).astype(np.uint64) * np.asarray(n).astype(np.uint64)) >> np.uint64(32)).astype(np.int64)
