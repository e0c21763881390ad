"""Row order in HBM (csrc/fpx_kernels.cuh: row_key): the hash the sketch kernel counts with, d * kRowMult — a bijection
of the docid whose top SKLOG bits are the sketch counter (search_find_kernel in csrc/fpx_kernels.cu: t = h >> (32 - SKLOG),
word byte offset = t & (2^SKLOG - 4), byte = t & 3; SKLOG = 14 or 15), so that one counter's postings are one contiguous
range of a sorted row.  The constants are read from the sources so that the test follows them."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "acoustid-index_b200", "csrc")


def test_row_key_is_a_bijection_with_the_counter_on_top():
    cuh = open(os.path.join(CSRC, "fpx_kernels.cuh")).read()
    cu = open(os.path.join(CSRC, "fpx_kernels.cu")).read()
    mult = int(re.search(r"kRowMult = (0x[0-9A-Fa-f]+)u", cuh).group(1), 16)
    assert re.search(r"kMult = (0x[0-9A-Fa-f]+)u", cu).group(1).lower() == hex(mult)
    assert "uint32_t row_key(uint32_t d) { return d * kRowMult; }" in cuh
    # the counter loop: t = h >> kKeyShift; word byte offset = t & kWordMask; byte = t & 3 (shift amount t << 3, modulo 32);
    # the resolvers: a hot counter id (word * 4 + byte) << kKeyShift is the first row key of its range
    assert "constexpr uint32_t kKeyShift = 32 - SKLOG;" in cu and "constexpr uint32_t kWordMask = kSketchBytes - 4;" in cu
    assert re.search(r"const uint32_t t = \(dd\[e\] \* kMult\) >> kKeyShift;", cu)
    assert "(t & kWordMask) | boff" in cu and "__funnelshift_l(0u, 1u, t << 3)" in cu
    assert "<< kKeyShift; // first row key of the counter" in cu

    rng = np.random.default_rng(7)
    d = np.concatenate([rng.integers(0, 1 << 32, size=200000, dtype=np.uint64),
                        np.arange(0, 70000, dtype=np.uint64), np.array([0xFFFFFFFF], dtype=np.uint64)])
    key = (d * mult) & 0xFFFFFFFF
    inv = pow(mult, -1, 1 << 32)
    assert np.array_equal((key * inv) & 0xFFFFFFFF, d)
    assert len(np.unique(key[:200000])) == len(np.unique(d[:200000]))
    for sklog in (14, 15):
        shift, mask = 32 - sklog, (1 << sklog) - 4
        t = key >> shift
        word, byte = (t & mask) >> 2, t & 3
        assert np.array_equal(key >> shift, word * 4 + byte)     # the counter is the key's top SKLOG bits
        assert word.max() < (1 << sklog) // 4
