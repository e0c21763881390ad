"""Product segment writer / block decoder vs the oracle's independent restatement, byte for byte (CPU only).
Reference: src/block.zig:438-567, src/filefmt.zig:94-138."""
import numpy as np
import pytest

from _helpers import pkg
from _oracle import OracleIndex, lib as olib, u8p, u32p


def _random_items(rng, n_docs, hashes_per_doc, vocab, first_id=1, hot=None, dup=False):
    ids = np.repeat(np.arange(first_id, first_id + n_docs, dtype=np.uint64), hashes_per_doc)
    hs = rng.integers(0, vocab, size=len(ids), dtype=np.uint64)
    if hot is not None:  # a few very hot hashes so that runs span many blocks
        mask = rng.random(len(ids)) < 0.3
        hs[mask] = rng.choice(np.asarray(hot, dtype=np.uint64), size=int(mask.sum()))
    if dup:
        hs[1::7] = hs[0::7][: len(hs[1::7])]
    hs = (hs * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF) if vocab > 5000 else hs
    items = np.sort((hs << np.uint64(32)) | ids)
    return items, np.arange(first_id, first_id + n_docs, dtype=np.uint32)


@pytest.mark.parametrize("n_docs,H,vocab,first_id,hot,block_size", [
    (1, 1, 10, 1, None, 512),
    (7, 3, 10, 1, None, 512),
    (300, 50, 1000, 1, None, 512),
    (3000, 40, 1 << 20, 1000000, None, 512),
    (5000, 30, 200, 1, [3, 77], 512),
    (5000, 30, 1 << 16, 70000, [3, 77, 4000000000], 512),
    (2000, 30, 500, 1, None, 64),
    (2000, 30, 500, 1, None, 4096),
    (40000, 25, 1 << 18, 5, [123456], 512),
])
def test_writer_matches_oracle_bytes(n_docs, H, vocab, first_id, hot, block_size):
    rng = np.random.default_rng(n_docs * 31 + H)
    items, doc_ids = _random_items(rng, n_docs, H, vocab, first_id, hot, dup=True)
    alive = np.ones(len(doc_ids), np.uint8)
    for threads in (1, 4):
        seg = pkg.FileSegment.from_items(items, doc_ids, alive, commit_id=1, block_size=block_size, threads=threads)
        ix = OracleIndex(block_size)
        ix.add_file_segment_sorted(items, doc_ids, alive)
        v = ix.file_segment(0)
        assert seg.num_blocks == v.num_blocks and seg.min_doc_id == v.min_doc_id
        ob = np.ctypeslib.as_array(v.blocks, shape=((v.num_blocks + 1) * block_size,))
        assert np.array_equal(seg.blocks, ob), "block bytes differ (threads=%d)" % threads
        oi = np.ctypeslib.as_array(v.block_index, shape=(v.num_blocks,))
        assert np.array_equal(seg.block_index, oi)


def test_empty_segment_is_just_the_terminator():
    seg = pkg.FileSegment.from_items(np.zeros(0, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint8), 1)
    assert seg.num_blocks == 0 and len(seg.blocks) == 512 and not seg.blocks.any()


def test_block_decode_matches_oracle():
    rng = np.random.default_rng(5)
    items, doc_ids = _random_items(rng, 4000, 30, 3000, 17, [5, 9], dup=True)
    seg = pkg.FileSegment.from_items(items, doc_ids, np.ones(len(doc_ids), np.uint8), 1)
    got = []
    h = np.zeros(2052, np.uint32)
    d = np.zeros(2052, np.uint32)
    oh = np.zeros(2052, np.uint32)
    od = np.zeros(2052, np.uint32)
    for b in range(seg.num_blocks):
        blk = seg.blocks[b * 512:(b + 2) * 512]  # oracle's SIMD decode may read 16 bytes past the block
        n = pkg.lib().fpx_block_decode(blk.ctypes.data, 512, seg.min_doc_id, h.ctypes.data, d.ctypes.data)
        m = olib().orc_decode_block(blk.ctypes.data_as(u8p), 512, seg.min_doc_id, oh.ctypes.data_as(u32p),
                                    od.ctypes.data_as(u32p))
        assert n == m and np.array_equal(h[:n], oh[:n]) and np.array_equal(d[:n], od[:n])
        got.append((h[:n].astype(np.uint64) << np.uint64(32)) | d[:n].astype(np.uint64))
    assert np.array_equal(np.concatenate(got), items)  # encode -> decode round trip


def test_corrupt_block_is_rejected_not_overread():
    blk = np.zeros(512, np.uint8)
    blk[4:6] = np.array([400], np.uint16).view(np.uint8)      # num_items = 400
    blk[6:8] = np.array([500], np.uint16).view(np.uint8)      # docids_offset past the end
    blk[8:108] = 0xFF                                          # every hash needs 4 bytes
    out = np.zeros(4096, np.uint32)
    assert pkg.lib().fpx_block_decode(blk.ctypes.data, 512, 1, out.ctypes.data, out.ctypes.data) == -1
