"""The hand-assembled golden segment files and manifest of tests/golden/ (every byte derived from the reference's
writer in tests/golden/make_golden.py: filefmt.zig:143-178, segment.zig:64-66, manifest.zig:41-47, block bytes of
SURVEY.md Appendix B) through the product's reader, writer, host snapshot compiler and the oracle; the `-m gpu` test
loads the directory into an HBM snapshot and searches it."""
import importlib.util
import os

import numpy as np
import pytest

from _helpers import have_gpu, pkg
from _oracle import OracleIndex

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
A, B = "0000000000000007-00000000.data", "0000000000000009-00000002.data"
ITEMS = [(1, 100), (1, 200), (3, 300), (4, 400), (5, 500)]          # (hash, id), SURVEY.md Appendix B
DOCS = {50: False, 100: True, 200: True, 300: True, 400: True, 500: True}


def _script():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _read(name):
    return open(os.path.join(GOLDEN, name), "rb").read()


def _items_of(seg):
    out = []
    for b in range(seg.num_blocks):
        blk = np.ascontiguousarray(seg.blocks[b * seg.block_size:(b + 1) * seg.block_size])
        h, d = np.zeros(2048, np.uint32), np.zeros(2048, np.uint32)
        n = pkg.lib().fpx_block_decode(blk.ctypes.data, seg.block_size, seg.min_doc_id, h.ctypes.data, d.ctypes.data)
        assert n >= 0
        out += list(zip(h[:n].tolist(), d[:n].tolist()))
    return out


def test_committed_files_are_what_the_derivation_script_writes():
    for name, data in _script().FILES.items():
        assert _read(name) == data, name


@pytest.mark.parametrize("name,info,meta", [(A, (7, 0, None), {}), (B, (9, 2, 1000), {b"name": b"golden"})])
def test_reader_parses_the_golden_segment_files(name, info, meta):
    f = pkg.SegmentFile.parse(_read(name))
    assert f.info == info and f.metadata == meta and f.num_items == 5
    seg = f.segment
    assert (seg.commit_id, seg.merges, seg.block_size, seg.num_blocks, seg.min_doc_id) == (info[0], info[1], 512, 1, 50)
    assert seg.block_index.tolist() == [5]                            # hash of the block's last item (filefmt.zig:117)
    assert dict(zip(seg.doc_ids.tolist(), [bool(a) for a in seg.doc_alive])) == DOCS
    assert _items_of(seg) == ITEMS
    assert pkg.segment_file_name(*info[:2]) == name


def test_writer_reproduces_golden_a_byte_for_byte():
    """The product's block writer + file serializer on the same items and docs map must emit exactly the
    hand-assembled bytes (block bytes of Appendix B, shortest msgpack forms, padding, index, footer, CRC-64/XZ)."""
    items = np.array([(h << 32) | d for h, d in ITEMS], dtype=np.uint64)
    ids = np.array(sorted(DOCS), dtype=np.uint32)
    alive = np.array([DOCS[i] for i in sorted(DOCS)], dtype=np.uint8)
    seg = pkg.FileSegment.from_items(items, ids, alive, commit_id=7)
    assert seg.min_doc_id == 50
    assert pkg.segment_file_bytes(seg) == _read(A)


def test_manifest_and_index_directory():
    assert pkg.parse_manifest(_read("manifest")) == [(7, 0, None), (9, 2, 1000)]
    files = pkg.open_index_dir(GOLDEN)
    assert [f.info for f in files] == [(7, 0, None), (9, 2, 1000)]


def test_host_compiler_and_oracle_on_the_golden_directory():
    """Both files hold the same documents; the newer one (commit 9) supersedes every id of the older (its docs map
    mentions them all, Index.zig:133-149), so every posting counts once."""
    files = pkg.open_index_dir(GOLDEN)
    ctx = pkg.Context(host_only=True)
    b = pkg.SnapshotBuilder(ctx)
    for f in files:
        b.add_file_segment(f.segment)
    terms, offs, docids = b.csr()
    b.abort()
    ctx.close()
    assert terms.tolist() == [1, 3, 4, 5] and offs.tolist() == [0, 2, 3, 4, 5] and docids.tolist() == [100, 200, 300, 400, 500]
    ix = OracleIndex()
    for f in files:
        s = f.segment
        ix.adopt_file_segment(s.commit_id, s.merges, s.block_size, s.blocks, s.num_blocks, s.block_index, s.doc_ids, s.doc_alive)
    assert ix.search([1, 3, 4, 5], 10, 1, 0) == [(100, 1), (200, 1), (300, 1), (400, 1), (500, 1)]
    assert ix.search([1, 1, 5], 10, 1, 0) == [(100, 1), (200, 1), (500, 1)]


@pytest.mark.gpu
def test_golden_directory_through_the_gpu():
    if not have_gpu():
        pytest.skip("no CUDA device")
    files = pkg.open_index_dir(GOLDEN)
    ctx = pkg.Context(device=0)
    try:
        snap = pkg.swap_snapshot(ctx, [f.segment for f in files])
        info = snap.info()
        assert info["n_segments"] == 2 and info["n_postings"] == 5 and info["n_dropped_superseded"] == 5
        r = pkg.IndexReader(snap)
        assert r.search([1, 3, 4, 5], pkg.SearchOptions(10, 1, 0)) == [(100, 1), (200, 1), (300, 1), (400, 1), (500, 1)]
        assert r.search([1, 1, 5], pkg.SearchOptions(10, 1, 0)) == [(100, 1), (200, 1), (500, 1)]
        assert r.search([3, 4], pkg.SearchOptions(1, 1, 0)) == [(300, 1)]
        assert pkg.multi_index_search(r, pkg.SearchRequest([1, 3, 4, 5])) == [(100, 1), (200, 1), (300, 1), (400, 1), (500, 1)]
        older = pkg.swap_snapshot(ctx, [files[0].segment])           # the older file alone answers the same
        assert pkg.IndexReader(older).search([1, 5], pkg.SearchOptions(10, 1, 0)) == [(100, 1), (200, 1), (500, 1)]
        older.release()
        snap.release()
    finally:
        ctx.close()
