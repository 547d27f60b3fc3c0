"""Philox4x32-10 in numpy — the counter-based noise stream shared by the CUDA kernels and the oracle.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference draws its rounding noise from an
unseeded ``tf.random_uniform(X.shape[1:], 0, 1)`` (dynamic_fixed_point.py:36) which cannot be
replayed, so parity runs feed both sides the same explicit noise tensor.  This module produces the
tensor that ``lbt_quantize(mode=2)`` / ``lbt_noise_fill`` generate in-kernel (include/lbt.h):

    group g = j // 4 of inner index j           (noise ignores dim 0, like X.shape[1:])
    counter = (g & 0xffffffff, g >> 32, offset & 0xffffffff, offset >> 32)
    key     = (seed & 0xffffffff, seed >> 32)
    r[0..3] = philox4x32_10(counter, key)
    u[j]    = (r[j % 4] >> 8) * 2**-24          in [0, 1 - 2**-24]

Philox4x32-10 is the published Random123 algorithm (Salmon et al., SC'11); constants below are
the standard ones.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs uint32-valued arrays/scalars; returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & _MASK
    c1 = np.asarray(c1, dtype=np.uint64) & _MASK
    c2 = np.asarray(c2, dtype=np.uint64) & _MASK
    c3 = np.asarray(c3, dtype=np.uint64) & _MASK
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0                      # 64-bit product of two 32-bit values
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n1 = lo1
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def noise(n_inner, seed, offset):
    """fp32 uniform noise of length ``n_inner`` identical to ``lbt_noise_fill(u, n_inner, seed, offset)``."""
    n_inner = int(n_inner)
    ng = (n_inner + 3) // 4
    g = np.arange(ng, dtype=np.uint64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    offset = int(offset) & 0xFFFFFFFFFFFFFFFF
    r = philox4x32_10(g & _MASK, g >> np.uint64(32),
                      np.full(ng, offset & 0xFFFFFFFF, dtype=np.uint64),
                      np.full(ng, offset >> 32, dtype=np.uint64),
                      seed & 0xFFFFFFFF, seed >> 32)
    r = np.stack(r, axis=1).reshape(-1)[:n_inner]
    return ((r >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def make_offset(qid, step):
    """The (quantiser id, step) -> 64-bit Philox offset convention used by lbt_b200 (low word = id)."""
    return ((int(step) & 0xFFFFFFFF) << 32) | (int(qid) & 0xFFFFFFFF)
