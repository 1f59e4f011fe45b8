"""Philox4x32-10: published known-answer vectors, numpy restatement == C restatement."""
import ctypes as C
import os

import numpy as np

from oracle.philox import philox4x32_10, philox_dropout_mask, philox_indices

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Random123 kat_vectors, philox4x32 10 rounds: (counter, key) -> output
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def _c():
    return C.CDLL(os.path.join(ROOT, "oracle", "_build", "libphilox_ref.so"))


def test_known_answer_vectors_numpy_and_c():
    L = _c()
    for ctr, key, want in KAT:
        got = tuple(int(x) for x in philox4x32_10(*ctr, *key))
        assert got == want
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        L.philox_block(c, k, o)
        assert tuple(o) == want


def test_index_stream_numpy_equals_c_and_is_in_range():
    L = _c()
    for seed, step, size, B in ((0, 0, 1, 8), (7, 123456789012, 1000000, 256), (2**63 + 11, 2**40 + 3, 2**31 + 5, 77),
                                (3, 5, 10000, 4096)):
        out = np.zeros(B, dtype=np.int64)
        L.philox_indices(C.c_uint64(seed), C.c_uint64(step), C.c_uint64(size), C.c_int64(B), out.ctypes.data_as(C.c_void_p))
        ref = philox_indices(seed, step, size, B)
        assert np.array_equal(out, ref)
        assert ref.min() >= 0 and ref.max() < size
    # a stream is a pure function of (seed, step, b): a longer batch extends a shorter one
    assert np.array_equal(philox_indices(1, 2, 999, 64)[:32], philox_indices(1, 2, 999, 32))
    # uniformity sanity
    big = philox_indices(5, 9, 10, 4000)
    assert np.all(np.bincount(big, minlength=10) > 300)


def test_dropout_mask_rate():
    m = philox_dropout_mask(3, 17, 1, 256 * 256, 0.1)
    assert abs(m.mean() - 0.9) < 0.01
    assert philox_dropout_mask(3, 17, 1, 64, 0.0).all()
