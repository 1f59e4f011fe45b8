"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the engine's Philox4x32-10
index sampler and dropout-mask generator (jsrl_corl_b200/csrc/common.cuh).

Philox4x32-10 is the counter-based generator of Salmon et al., "Parallel Random
Numbers: As Easy as 1, 2, 3" (SC'11); the round function and the constants
below are the published ones (also used by Random123 / cuRAND / torch).  The
known-answer vectors of the Random123 distribution are checked in
tests/test_philox.py.  The *reference* samples with numpy's MT19937
(iql.py:172); that stream is reproduced on the host by the facade's
``sampler="numpy"`` mode, while this generator is the engine's native mode.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
STREAM_SAMPLE = 0
STREAM_DROPOUT_BASE = 1
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays (broadcast). Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(x, dtype=np.uint64) & MASK32 for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK32, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & MASK32, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in (c0, c1, c2, c3)]


def philox_indices(seed: int, step: int, size: int, batch: int) -> np.ndarray:
    """Indices of the b-th draws, b in [0, batch): see common.cuh::philox_index."""
    b = np.arange(batch, dtype=np.uint64)
    x, y, z, w = philox4x32_10(b >> np.uint64(1), step & 0xFFFFFFFF, (step >> 32) & 0xFFFFFFFF, STREAM_SAMPLE,
                               seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    odd = (b & np.uint64(1)).astype(bool)
    lo = np.where(odd, z, x).astype(np.uint64)
    hi = np.where(odd, w, y).astype(np.uint64)
    out = np.empty(batch, dtype=np.int64)
    for i in range(batch):  # 64x64 -> high 64 bits, exact in Python ints
        u = (int(hi[i]) << 32) | int(lo[i])
        out[i] = (u * int(size)) >> 64
    return out


def dropout_threshold(p: float) -> int:
    t = p * 4294967296.0
    if t <= 0.0:
        return 0
    if t >= 4294967295.0:
        return 4294967295
    return int(t)


def philox_dropout_mask(seed: int, step: int, layer: int, n_elems: int, p: float) -> np.ndarray:
    """Keep-mask (uint8) of n_elems consecutive elements (row-major [B][H], n_elems % 4 == 0)."""
    quads = np.arange((n_elems + 3) // 4, dtype=np.uint64)
    x, y, z, w = philox4x32_10(quads, step & 0xFFFFFFFF, (step >> 32) & 0xFFFFFFFF, STREAM_DROPOUT_BASE + layer,
                               seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    words = np.stack([x, y, z, w], axis=1).reshape(-1)[:n_elems]
    return (words >= np.uint32(dropout_threshold(p))).astype(np.uint8)
