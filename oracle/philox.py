"""oracle/philox.py -- TEST INFRASTRUCTURE.  numpy replay of the boundary sampler (include/bnsgcn.h,
bns_sample_boundary): Philox4x32-10 (Salmon et al., SC'11 -- the generator cuRAND calls
curandStatePhilox4_32_10_t), keys ``(segment << 56) | r56`` sorted stably, first k per segment.

What it restates from the reference: ``select_node`` (train.py:225-236) draws, per peer,
``np.random.choice(b, k, replace=False)`` -- a uniformly random ordered k-subset of the boundary list --
from an unseeded numpy stream, so the reference has no reproducible index sets to pin; the product's
counter-based sampler is pinned by this replay instead ("exact on sampled index sets").
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over uint32 arrays ``c0..c3``; ``k0, k1`` scalars.  Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def sample_boundary(boundaries, sizes, seed: int, offset: int):
    """``boundaries``: list of int64 arrays (one per peer segment, in segment order); ``sizes``: k per segment.
    Returns the list of selected arrays, bit-identical to bns_sample_boundary."""
    lens = [len(b) for b in boundaries]
    B = int(sum(lens))
    if B == 0:
        return [np.empty(0, dtype=np.int64) for _ in boundaries]
    i = np.arange(B, dtype=np.uint64)
    n = len(i)
    r0, r1, _, _ = philox4x32_10((i & MASK32).astype(np.uint32), (i >> np.uint64(32)).astype(np.uint32),
                                 np.full(n, offset & 0xFFFFFFFF, dtype=np.uint32),
                                 np.full(n, (offset >> 32) & 0xFFFFFFFF, dtype=np.uint32),
                                 seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    r56 = (r0.astype(np.uint64) << np.uint64(24)) | (r1.astype(np.uint64) >> np.uint64(8))
    seg = np.repeat(np.arange(len(lens), dtype=np.uint64), lens)
    key = (seg << np.uint64(56)) | r56
    order = np.argsort(key, kind="stable")
    cat = np.concatenate([np.asarray(b, dtype=np.int64) for b in boundaries])
    out, start = [], 0
    for ln, k in zip(lens, sizes):
        out.append(cat[order[start:start + k]])
        start += ln
    return out
