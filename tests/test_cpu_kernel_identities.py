"""Arithmetic identities the tail and entropy kernels rely on, restated in numpy and checked exhaustively or on
random words (no GPU): the kernels themselves are covered byte for byte by tests/test_gpu_*.py; these pin the
reasoning in their comments.

  csrc/tail.cuh     ff_count4: exact SWAR count of the 0xFF bytes of a word
  csrc/tail.cuh     wide / interior rounds: a group is the funnel of two numeric stream words shifted by g0 & 7
  csrc/entropy.cuh  PrivSinkBF::put: destination select min(wa | (full - 1), ea), nacc = (nacc + n) & 31
"""
import numpy as np


def ff_count4(v):
    v = v.astype(np.uint64)
    t = (~v) & 0xFFFFFFFF
    y = ((t & 0x7F7F7F7F) + 0x7F7F7F7F) & 0xFFFFFFFF
    m = (~(y | t)) & 0x80808080
    return np.array([bin(int(x)).count("1") for x in m], dtype=np.int64)


def ff_count_ref(v):
    b = v.astype(np.uint32).view(np.uint8).reshape(-1, 4)
    return (b == 0xFF).sum(axis=1).astype(np.int64)


def test_ff_count4_is_exact():
    rng = np.random.default_rng(7)
    words = rng.integers(0, 1 << 32, size=20000, dtype=np.uint64).astype(np.uint32)
    # words rich in 0xFF / 0xFE / 0x7F / 0x80 bytes: the cases a borrow-based zero-byte test gets wrong
    special = np.array([0xFF, 0xFE, 0x7F, 0x80, 0x00, 0x01], dtype=np.uint32)
    idx = rng.integers(0, len(special), size=(20000, 4))
    rich = (special[idx[:, 0]] | (special[idx[:, 1]] << 8) | (special[idx[:, 2]] << 16) | (special[idx[:, 3]] << 24)).astype(np.uint32)
    allw = np.concatenate([words, rich, np.array([0, 0xFFFFFFFF, 0xFF00FF00, 0x00FF00FF, 0xFEFFFEFF, 0xFF7FFF80], dtype=np.uint32)])
    assert (ff_count4(allw) == ff_count_ref(allw)).all()


def test_owned_byte_groups_are_funnels_of_stream_words():
    """Owned bytes 4i..4i+3 of a chunk whose first bit sits at global bit g0 are the 32 stream bits at local offset
    -(g0 & 7) + 32 i: (W[i-1] << (32 - nd)) | (W[i] >> nd) with nd = g0 & 7 and W[-1] the 32 bits in front of the chunk."""
    rng = np.random.default_rng(11)
    for nd in range(8):
        front = rng.integers(0, 2, size=32, dtype=np.uint8)
        bits = rng.integers(0, 2, size=32 * 40, dtype=np.uint8)
        stream = np.concatenate([front, bits])  # local bit b of the chunk is stream[32 + b]
        W = [int("".join(map(str, stream[32 * k:32 * k + 32])), 2) for k in range(41)]  # W[0] = front, W[k + 1] = word k
        for i in range(39):
            lo, hi = W[i + 1], W[i]
            v = ((lo >> nd) | (hi << (32 - nd))) & 0xFFFFFFFF if nd else lo
            want = int("".join(map(str, stream[32 + 32 * i - nd:32 + 32 * i - nd + 32])), 2)
            assert v == want, (nd, i)


def test_private_bit_writer_select():
    """put(): t = nacc + n < 64, full = t >> 5, nacc' = t & 31; the store goes to min(wa | (full - 1), ea): the pad word ea
    when no word was completed (full - 1 = all ones) or the region is used up (wa >= ea), else the next word wa."""
    ea = 0x6014
    for nacc in range(32):
        for n in range(0, 28):
            t = nacc + n
            full = t >> 5
            assert full in (0, 1) and (t & 31) == (t - 32 if t >= 32 else t)
            for wa in (0x6000, 0x6004, 0x6010, 0x6014, 0x6018, 0x7000):
                dst = min((wa | ((full - 1) & 0xFFFFFFFF)), ea)
                want = wa if (full and wa < ea) else ea
                assert dst == want, (nacc, n, wa)
    # a word of length 0 changes nothing (acc << 0 | 0, nacc + 0): the writer may be called with an empty symbol
    acc, nacc = 0x1234, 7
    assert ((acc << 0) | 0, (nacc + 0) & 31, (nacc + 0) >> 5) == (acc, nacc, 0)
