"""Row order in HBM (csrc/fpx_kernels.cuh: row_key): a bijection of the docid whose top five bits are the
shared-memory bank of the sketch word the docid is counted in (csrc/fpx_kernels.cu: word = hash bits 29..17).
The constants are read from the sources so that the test follows them."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "acoustid-index_b200", "csrc")


def _constants():
    cuh = open(os.path.join(CSRC, "fpx_kernels.cuh")).read()
    cu = open(os.path.join(CSRC, "fpx_kernels.cu")).read()
    mult = int(re.search(r"kRowMult = (0x[0-9A-Fa-f]+)u", cuh).group(1), 16)
    rot = int(re.search(r"return \(h << (\d+)\) \| \(h >> (\d+)\);", cuh).group(1))
    word = re.search(r"sketch_b \+ \(\(hv >> (\d+)\) & (0x[0-9A-Fa-f]+)u\)", cu)
    return mult, rot, int(word.group(1)), int(word.group(2), 16), cu


def test_row_key_is_a_bijection_with_the_bank_bits_on_top():
    mult, rot, shift, mask, cu = _constants()
    assert re.search(r"kMult = (0x[0-9A-Fa-f]+)u", cu).group(1).lower() == hex(mult)
    rng = np.random.default_rng(7)
    d = np.concatenate([rng.integers(0, 1 << 32, size=200000, dtype=np.uint64),
                        np.arange(0, 70000, dtype=np.uint64), np.array([0xFFFFFFFF], dtype=np.uint64)])
    h = (d * mult) & 0xFFFFFFFF
    key = ((h << rot) | (h >> (32 - rot))) & 0xFFFFFFFF
    # inverse: rotate back, multiply by the modular inverse
    inv = pow(mult, -1, 1 << 32)
    back = ((((key >> rot) | (key << (32 - rot))) & 0xFFFFFFFF) * inv) & 0xFFFFFFFF
    assert np.array_equal(back, d)
    assert len(np.unique(key[:200000])) == len(np.unique(d[:200000]))
    # the sketch word's byte offset is (h >> shift) & mask; its bank is the word index modulo 32
    bank = (((h >> shift) & mask) >> 2) & 31
    assert np.array_equal(key >> 27, bank)
