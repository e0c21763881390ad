"""Row order in HBM (csrc/fpx_kernels.cuh: row_key): a bijection of the docid whose top 15 bits are the sketch
counter the docid is counted in by search_find_kernel (csrc/fpx_kernels.cu: h = docid * kRowMult, word = h[29:17],
byte = h[16:15]) — shared-memory bank of the word first, so that a row sweeps the banks in order.  The constants are
read from the sources so that the test follows them; the bit arithmetic is restated here in numpy."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "acoustid-index_b200", "csrc")


def _sources():
    return (open(os.path.join(CSRC, "fpx_kernels.cuh")).read(), open(os.path.join(CSRC, "fpx_kernels.cu")).read())


def _row_key(h):
    return (((h << 10) & 0xF8000000) | ((h >> 3) & 0x07F80000) | ((h << 2) & 0x00060000) | ((h >> 15) & 0x00018000) |
            (h & 0x00007FFF))


def _row_key_inv(k):
    return (((k >> 10) & 0x003E0000) | ((k << 3) & 0x3FC00000) | ((k >> 2) & 0x00018000) | ((k << 15) & 0xC0000000) |
            (k & 0x00007FFF))


def test_row_key_is_a_bijection_with_the_counter_on_top():
    cuh, cu = _sources()
    mult = int(re.search(r"kRowMult = (0x[0-9A-Fa-f]+)u", cuh).group(1), 16)
    assert re.search(r"kMult = (0x[0-9A-Fa-f]+)u", cu).group(1).lower() == hex(mult)
    # the masks in the header are the ones restated above
    for m in ("0xF8000000u", "0x07F80000u", "0x00060000u", "0x00018000u", "0x00007FFFu",
              "0x003E0000u", "0x3FC00000u", "0xC0000000u"):
        assert m in cuh, m
    # the counter loop: t = h >> 15; word byte offset = t & 0x7FFC; byte = t & 3 (shift amount t << 3, modulo 32)
    assert re.search(r"const uint32_t t = \(dd\[e\] \* kMult\) >> 15;", cu)
    assert "(t & 0x7FFCu) | boff" in cu and "__funnelshift_l(0u, 1u, t << 3)" in cu

    rng = np.random.default_rng(7)
    d = np.concatenate([rng.integers(0, 1 << 32, size=200000, dtype=np.uint64),
                        np.arange(0, 70000, dtype=np.uint64), np.array([0xFFFFFFFF], dtype=np.uint64)])
    h = (d * mult) & 0xFFFFFFFF
    key = _row_key(h) & 0xFFFFFFFF
    inv = pow(mult, -1, 1 << 32)
    back = ((_row_key_inv(key) & 0xFFFFFFFF) * inv) & 0xFFFFFFFF
    assert np.array_equal(back, d)
    assert len(np.unique(key[:200000])) == len(np.unique(d[:200000]))
    t = h >> 15
    word, byte = (t & 0x7FFC) >> 2, t & 3
    # top 5 bits: the word's bank; top 15 bits: the counter, as the resolvers rebuild it from (word, byte)
    assert np.array_equal(key >> 27, word & 31)
    assert np.array_equal(key >> 17, ((word & 31) << 10) | ((word >> 5) << 2) | byte)
