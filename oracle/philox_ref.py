"""numpy restatement of the Philox4x32-10 stream the fused step kernel uses (TEST INFRASTRUCTURE).

Counter = (pixel_group, stream_word, gidx_lo, gidx_hi), key = (seed_lo, seed_hi); stream_word 0 is
the initial state x ~ N(0,1), word 1+i is SDE step i.  Four normals per counter (pixels
4*group..4*group+3) by Box-Muller on (u0,u1) and (u2,u3), u = ((r >> 8) + 0.5) / 2^24.
Pinned against the published Philox4x32-10 known-answer vectors (Random123 kat_vectors) in
tests/test_cpu_oracle.py."""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & MASK for v in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0 & MASK, n1, n2 & MASK, n3
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def _u01(r):
    return ((r >> np.uint32(8)).astype(np.float64) + 0.5) / 16777216.0


def normal_image(seed: int, gidx: int, word: int) -> np.ndarray:
    """[4096] float64 N(0,1) draws of one sample for one stream word."""
    g = np.arange(1024, dtype=np.uint64)
    r0, r1, r2, r3 = philox4x32_10(g, np.full(1024, word), np.full(1024, gidx & 0xFFFFFFFF),
                                   np.full(1024, (gidx >> 32) & 0xFFFFFFFF), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    ra, rb = np.sqrt(-2.0 * np.log(_u01(r0))), np.sqrt(-2.0 * np.log(_u01(r2)))
    ta, tb = 2.0 * np.pi * _u01(r1), 2.0 * np.pi * _u01(r3)
    out = np.stack([ra * np.cos(ta), ra * np.sin(ta), rb * np.cos(tb), rb * np.sin(tb)], axis=1)
    return out.reshape(4096)
