"""Pins the CPU oracle against every known-answer vector the reference's own tests hold
for the `_search` path.  Each test names the reference test it restates
(paths relative to the reference root).  CPU only."""
import ctypes as C
import random

import numpy as np
import pytest

from _oracle import OracleIndex, items_u64, lib, u8p, u32p, u64p


def _u32(vals):
    return (C.c_uint32 * len(vals))(*vals)


def _bytes16(vals):
    b = list(vals) + [0] * (32 - len(vals))
    return (C.c_uint8 * len(b))(*b)


# ---------------------------------------------------------------- streamvbyte.zig KATs


def test_svb_decode_quad_0124():  # streamvbyte.zig:526-540
    out = _u32([0] * 4)
    n = lib().orc_svb_decode_quad(0, 0b01_00_01_01, _bytes16([1, 2, 4]), out)
    assert n == 3 and list(out) == [1, 2, 0, 4]


def test_svb_decode_quad_1234():  # streamvbyte.zig:542-556, 588-599
    out = _u32([0] * 4)
    n = lib().orc_svb_decode_quad(1, 0, _bytes16([1, 2, 3, 4]), out)
    assert n == 4 and list(out) == [1, 2, 3, 4]


def test_svb_decode_quad_fused_delta():  # streamvbyte.zig:558-586
    out = _u32([0] * 4)
    n = lib().orc_svb_decode_quad_delta(1, 0, _bytes16([10, 5, 3, 2]), out, 100)
    assert n == 4 and list(out) == [110, 115, 118, 120]
    n = lib().orc_svb_decode_quad_delta(0, 0b01_00_01_01, _bytes16([1, 2, 4]), out, 50)
    assert n == 3 and list(out) == [51, 53, 53, 57]


def test_svb_decode_quad_0124_minus1():  # streamvbyte.zig:821-835
    out = _u32([0] * 4)
    n = lib().orc_svb_decode_quad(2, 0b01_00_01_00, _bytes16([1, 3]), out)
    assert n == 2 and list(out) == [1, 2, 1, 4]


@pytest.mark.parametrize("data,first,want", [
    ([10, 5, 3, 2], 100, [110, 115, 118, 120]),                       # :601-611
    (list(range(1, 17)), 0, [1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 66, 78, 91, 105, 120, 136]),  # :613-622
    ([], 100, []), ([42], 100, [142]), ([10, 20], 100, [110, 130]), ([1, 2, 3], 0, [1, 3, 6]),  # :624-646
])
def test_svb_delta_decode_in_place(data, first, want):
    a = _u32(data if data else [0])
    lib().orc_svb_delta_decode_in_place(a, len(data), first)
    assert list(a)[:len(data)] == want


def test_svb_decode_values_40_items():  # streamvbyte.zig:648-695
    buf = [0b01_01_01_01] * 10 + list(range(1, 41)) + [0] * 16
    inp = (C.c_uint8 * len(buf))(*buf)
    out = _u32([0] * 40)
    lib().orc_svb_decode_values(40, 0, 40, inp, out, 0, 0, 0)
    assert list(out) == list(range(1, 41))


def test_svb_decode_values_fused_delta():  # streamvbyte.zig:851-908
    buf = [0, 0, 10, 5, 3, 2, 1, 1, 1, 1] + [0] * 16
    inp = (C.c_uint8 * len(buf))(*buf)
    out = _u32([0] * 8)
    lib().orc_svb_decode_values(8, 0, 8, inp, out, 1, 1, 100)
    assert list(out) == [110, 115, 118, 120, 121, 122, 123, 124]


@pytest.mark.parametrize("fn,vals,ctrl,data", [
    ("0124", [1, 2, 0, 4], 0x45, [1, 2, 4]),                                             # :697-712
    ("0124", [0, 255, 65535, 0x12345678], 0xE4, [255, 255, 255, 0x78, 0x56, 0x34, 0x12]),  # :714-731
    ("1234", [1, 2, 3, 4], 0x00, [1, 2, 3, 4]),                                          # :733-749
    ("1234", [255, 65535, 0xFFFFFF, 0x12345678], 0xE4,
     [255, 255, 255, 255, 255, 255, 0x78, 0x56, 0x34, 0x12]),                             # :751-769
    ("1234", [0, 1, 0, 255], 0x00, [0, 1, 0, 255]),                                      # :771-787
    ("0124", [0, 0, 0, 0], 0x00, []),                                                    # :796-800
    ("0124", [0xFFFFFFFF] * 4, 0xFF, [255] * 16),                                        # :802-806
    ("1234", [0xFFFFFFFF] * 4, 0xFF, [255] * 16),                                        # :837-841
])
def test_svb_encode_quad(fn, vals, ctrl, data):
    out = (C.c_uint8 * 32)()
    c = C.c_uint8(0xAA)
    f = getattr(lib(), "orc_svb_encode_quad_" + fn)
    n = f(_u32(vals), out, C.byref(c))
    assert n == len(data)
    assert c.value == ctrl
    assert list(out)[:n] == data


def test_svb_roundtrip_random_ranges():
    """decodeValues range decode == full decode, both variants (streamvbyte.zig:341-412)."""
    rng = random.Random(7)
    L = lib()
    for variant, enc in ((0, L.orc_svb_encode_quad_0124), (1, L.orc_svb_encode_quad_1234)):
        for n in (1, 3, 4, 5, 31, 32, 33, 100, 333):
            vals = [rng.choice([0, 1, 200, 255, 256, 65535, 65536, 1 << 24, 0xFFFFFFFF, rng.getrandbits(32)])
                    for _ in range(n)]
            padded = vals + [0] * ((-n) % 4)
            ctrls, data = [], []
            for q in range(0, len(padded), 4):
                out = (C.c_uint8 * 16)()
                c = C.c_uint8(0)
                k = enc(_u32(padded[q:q + 4]), out, C.byref(c))
                ctrls.append(c.value)
                data.extend(list(out)[:k])
            buf = ctrls + data + [0] * 16
            inp = (C.c_uint8 * len(buf))(*buf)
            out = _u32([0] * (len(padded) + 4))
            L.orc_svb_decode_values(n, 0, n, inp, out, variant, 0, 0)
            assert list(out)[:n] == vals
            s = rng.randrange(0, n)
            e = rng.randrange(s, n + 1)
            out2 = _u32([0xDEAD] * (len(padded) + 4))
            L.orc_svb_decode_values(n, s, e, inp, out2, variant, 0, 0)
            assert list(out2)[s:e] == vals[s:e]


# ---------------------------------------------------------------- block.zig KATs


def _encode_block(pairs, min_doc_id, size):
    items = items_u64(pairs)
    out = np.zeros(size + 16, dtype=np.uint8)
    n = lib().orc_encode_block(items.ctypes.data_as(u64p), len(items), min_doc_id,
                               out.ctypes.data_as(u8p), size)
    return n, out


def _decode_block(block, size, min_doc_id):
    h = np.zeros(2052, dtype=np.uint32)
    d = np.zeros(2052, dtype=np.uint32)
    n = lib().orc_decode_block(block.ctypes.data_as(u8p), size, min_doc_id,
                               h.ctypes.data_as(u32p), d.ctypes.data_as(u32p))
    return list(zip(h[:n].tolist(), d[:n].tolist()))


def _search_hash(block, size, min_doc_id, hash_):
    s, e = C.c_uint32(), C.c_uint32()
    d = np.zeros(2052, dtype=np.uint32)
    n = lib().orc_block_search_hash(block.ctypes.data_as(u8p), size, min_doc_id, hash_,
                                    C.byref(s), C.byref(e), d.ctypes.data_as(u32p))
    return (s.value, e.value), d[:n].tolist()


def test_block_reader_basic():  # block.zig:317-361
    n, blk = _encode_block([(100, 1), (100, 2), (200, 3), (300, 4)], 1, 256)
    assert n == 4
    assert int(blk[:4].view(np.uint32)[0]) == 100 and int(blk[4:6].view(np.uint16)[0]) == 4
    assert _search_hash(blk, 256, 1, 100) == ((0, 2), [1, 2])
    assert _search_hash(blk, 256, 1, 200) == ((2, 3), [3])
    assert _search_hash(blk, 256, 1, 404) == ((4, 4), [])


def test_block_reader_range_decoding():  # block.zig:363-417
    pairs = [(100, 1001), (100, 1005), (100, 1010), (200, 2001), (200, 2002), (300, 3001), (300, 3002), (300, 3003)]
    n, blk = _encode_block(pairs, 1000, 512)
    assert n == 8
    assert _search_hash(blk, 512, 1000, 100) == ((0, 3), [1001, 1005, 1010])
    assert _search_hash(blk, 512, 1000, 200) == ((3, 5), [2001, 2002])
    assert _search_hash(blk, 512, 1000, 300) == ((5, 8), [3001, 3002, 3003])


def test_block_mixed_and_byte_layout():  # block.zig:585-640 + SURVEY.md Appendix B (hand-derived bytes)
    pairs = [(1, 100), (1, 200), (3, 300), (4, 400), (5, 500)]
    n, blk = _encode_block(pairs, 50, 256)
    assert n == 5
    want = [0x01, 0, 0, 0, 0x05, 0, 0x05, 0, 0x50, 0x01, 0x02, 0x01, 0x01, 0x40, 0x01,
            0x32, 0x64, 0xFA, 0x5E, 0x01, 0xC2, 0x01, 0x00, 0x00, 0x00]
    assert blk[:len(want)].tolist() == want
    assert not blk[len(want):256].any()
    assert _decode_block(blk, 256, 50) == pairs
    assert _search_hash(blk, 256, 50, 1) == ((0, 2), [100, 200])
    assert _search_hash(blk, 256, 50, 3) == ((2, 3), [300])
    assert _search_hash(blk, 256, 50, 4) == ((3, 4), [400])
    assert _search_hash(blk, 256, 50, 5) == ((4, 5), [500])
    # the 25-byte accounting of Appendix B: a 25-byte block holds all five, a 24-byte one only the first quad
    assert _encode_block(pairs, 50, 64)[0] == 5


def test_block_duplicate_hashes():  # block.zig:642-679
    n, blk = _encode_block([(100, 1), (100, 2), (100, 3)], 1, 256)
    assert n == 3
    assert _search_hash(blk, 256, 1, 100)[1] == [1, 2, 3]
    assert _decode_block(blk, 256, 1) == [(100, 1), (100, 2), (100, 3)]


def test_block_same_hash_continues_in_next_block():  # block.zig:681-719
    n1, b1 = _encode_block([(100, 1), (100, 2)], 1, 256)
    n2, b2 = _encode_block([(100, 3), (100, 4)], 1, 256)
    assert (n1, n2) == (2, 2)
    assert _search_hash(b1, 256, 1, 100)[1] == [1, 2]
    assert _search_hash(b2, 256, 1, 100)[1] == [3, 4]


def test_block_header_roundtrip():  # block.zig:570-583 (layout: u32 min_hash, u16 num_items, u16 docids_offset)
    pairs = [(12345678, 7 + i) for i in range(25)]
    n, blk = _encode_block(pairs, 7, 512)
    assert n == 25
    assert int(blk[:4].view(np.uint32)[0]) == 12345678
    assert int(blk[4:6].view(np.uint16)[0]) == 25
    # 7 hash control bytes, zero hash data bytes (all deltas 0 => 0 bytes in the 0124 variant)
    assert int(blk[6:8].view(np.uint16)[0]) == 7


def test_block_full_and_capacity():
    """block.zig:479-485: a quad is accepted iff the running size stays <= block_size.
    Same-hash run, ids 1..: per quad 1 hash ctrl + 0 hash data + 1 docid ctrl + 4 docid bytes = 6 B
    => (512-8)/6 = 84 quads = 336 items."""
    pairs = [(9, 1 + i) for i in range(400)]
    n, blk = _encode_block(pairs, 1, 512)
    assert n == 336
    assert _decode_block(blk, 512, 1) == pairs[:336]


def test_item_layout():  # segment.zig:112-143
    assert int(items_u64([(1, 2)])[0]) == 0x0000000100000002
    assert items_u64([(2, 200), (2, 100), (1, 300)]).tolist() == [(1 << 32) | 300, (2 << 32) | 100, (2 << 32) | 200]


# ---------------------------------------------------------------- segment / index KATs


def test_segment_roundtrip_write_read_search():  # filefmt.zig:293-338
    ix = OracleIndex()
    ix.update([("insert", 1, [100, 200, 300]), ("insert", 2, [100, 200])])
    assert ix.search([100, 200, 300], 10, 1) == [(1, 3), (2, 2)]
    ix.checkpoint()
    assert ix.num_file_segments == 1 and ix.num_memory_segments == 0
    v = ix.file_segment(0)
    assert v.n_docs == 2 and v.commit_id == 1 and v.num_items == 5
    assert ix.search([100, 200, 300], 10, 1, 0) == [(1, 3), (2, 2)]


def test_duplicate_query_hashes():  # Index.zig:1056-1096
    ix = OracleIndex()
    ix.update([("insert", 1, [100, 200])])
    assert ix.search([100, 100], 10, 1) == [(1, 1)]
    ix.checkpoint()
    assert ix.num_file_segments == 1
    assert ix.search([100, 100], 10, 1) == [(1, 1)]


def test_checkpoint_and_reload_scores():  # Index.zig:1311-1364
    ix = OracleIndex()
    ix.update([("insert", 1, [100, 200, 300])])
    ix.update([("insert", 2, [100, 200, 300])])
    ix.checkpoint()
    assert ix.num_file_segments == 1 and ix.num_memory_segments == 0
    assert ix.file_segment(0).commit_id == 1 and ix.file_segment(0).merges == 1
    assert ix.search([100, 200, 300], 10, 1) == [(1, 3), (2, 3)]


def test_file_merging_preserves_deletes():  # Index.zig:1366-1401
    ix = OracleIndex()
    for i in range(1, 31):
        ix.update([("insert", i, [100, i])])
        ix.checkpoint()
        if ix.num_file_segments >= 4:
            ix.merge_files(0, ix.num_file_segments)
    ix.update([("delete", 5)])
    out = ix.search([100], 100, 1)
    assert len(out) == 29 and all(i != 5 for i, _ in out)
    # and it stays that way when the tombstone is flushed and merged down
    ix.checkpoint()
    ix.merge_files(0, ix.num_file_segments)
    out = ix.search([100], 100, 1)
    assert len(out) == 29 and all(i != 5 for i, _ in out)


def test_snapshot_many_single_hash_docs():  # Index.zig:1403-1444 (the "fresh" reader part)
    ix = OracleIndex()
    for i in range(1, 31):
        ix.update([("insert", i, [100])])
        if i % 7 == 0:
            ix.checkpoint()
    assert len(ix.search([100], 100, 1)) == 30


def test_memory_merge_keeps_everything_searchable():  # Index.zig:1446-1479
    ix = OracleIndex()
    for i in range(1, 51):
        ix.update([("insert", i, [i])])
        if ix.num_memory_segments >= 16:
            ix.merge_memory(0, 10)
    assert ix.num_file_segments == 0 and ix.num_memory_segments < 50
    assert ix.search([25], 100, 1) == [(25, 1)]


# ---------------------------------------------------------------- e2e (HTTP defaults) KATs


def test_e2e_insert_single():  # tests/test_fingerprint_api.py:5-26
    ix = OracleIndex()
    ix.update([("insert", 1, [101, 201, 301])])
    assert ix.search_http([101, 201, 301]) == [(1, 3)]


def test_e2e_insert_multi_tie_id_ascending():  # tests/test_fingerprint_api.py:29-52
    ix = OracleIndex()
    ix.update([("insert", 1, [101, 201, 301]), ("insert", 2, [102, 202, 302])])
    assert ix.search_http([101, 201, 301, 102, 202, 302]) == [(1, 3), (2, 3)]


def _insert_many_corpus():
    """tests/test_fingerprint_api.py:67-99: 50 000 fps x 100 hashes, CPython random.Random(i)."""
    max_hash = 2 ** 18
    batches, batch = [], []
    for i in range(1, 50001):
        rng = random.Random(i)
        batch.append(("insert", i, [rng.randint(0, max_hash) for _ in range(100)]))
        if len(batch) == 1000:
            batches.append(batch)
            batch = []
    rng = random.Random(100)
    query = [rng.randint(0, max_hash) for _ in range(100)]
    return batches, query


def test_e2e_insert_many():  # tests/test_fingerprint_api.py:67-99
    batches, query = _insert_many_corpus()
    ix = OracleIndex()
    for n, b in enumerate(batches):
        ix.update(b)
        if ix.num_memory_segments >= 16:       # Index.zig:679-687 (memory capped at ~16 segments)
            ix.merge_memory(0, 10)
        if n % 20 == 19:                       # > checkpoint_threshold items => flush (main.zig:44)
            ix.checkpoint()
    assert ix.num_file_segments >= 1 and ix.num_memory_segments >= 1
    assert ix.search_http(query) == [(100, 100)]
    ix.checkpoint()
    ix.merge_files(0, ix.num_file_segments)
    assert ix.num_file_segments == 1
    assert ix.search_http(query) == [(100, 100)]


def test_e2e_update_full():  # tests/test_fingerprint_api.py:102-146
    ix = OracleIndex()
    ix.update([("insert", 1, [100, 200, 300])])
    ix.update([("insert", 1, [1000, 2000, 3000])])
    assert ix.search_http([100, 200, 300]) == []
    assert ix.search_http([1000, 2000, 3000]) == [(1, 3)]
    ix.checkpoint()
    assert ix.search_http([100, 200, 300]) == []
    assert ix.search_http([1000, 2000, 3000]) == [(1, 3)]


def test_e2e_update_partial():  # tests/test_fingerprint_api.py:149-189
    for flush_between in (False, True):
        ix = OracleIndex()
        ix.update([("insert", 1, [100, 200, 300])])
        if flush_between:
            ix.checkpoint()
        ix.update([("insert", 1, [100, 200, 999])])
        assert ix.search_http([100, 200, 300]) == [(1, 2)]
        assert ix.search_http([100, 200, 999]) == [(1, 3)]


def test_e2e_delete_multi_and_single():  # tests/test_fingerprint_api.py:192-260
    ix = OracleIndex()
    ix.update([("insert", 1, [101, 201, 301]), ("insert", 2, [102, 202, 302])])
    ix.update([("delete", 1), ("delete", 2)])
    assert ix.search_http([101, 201, 301, 102, 202, 302]) == []
    ix2 = OracleIndex()
    ix2.update([("insert", 1, [100, 200, 300])])
    ix2.checkpoint()
    ix2.update([("delete", 1)])
    assert ix2.search_http([100, 200, 300]) == []


def test_e2e_legacy_search():  # tests/test_legacy.py:61-69 -> "OK 1001:3 1002:2" (limit 500, min_score 1, pct 10)
    ix = OracleIndex()
    ix.update([("insert", 1001, [11000, 12000, 13000]), ("insert", 1002, [11000, 12000, 19000])])
    assert ix.search([11000, 12000, 13000], 500, 1, 10) == [(1001, 3), (1002, 2)]
    assert ix.search([11000, 12000, 19000], 500, 1, 10) == [(1002, 3), (1001, 2)]


def test_last_change_in_batch_wins():  # MemorySegment.zig:81-148 (reverse pass)
    ix = OracleIndex()
    ix.update([("insert", 1, [10, 20]), ("insert", 1, [30]), ("delete", 2), ("insert", 2, [10])])
    assert ix.search([10, 20, 30], 10, 1, 0) == [(1, 1), (2, 1)]
    v = ix.memory_segment(0)
    assert v.n_items == 2 and v.min_doc_id == 1 and v.max_doc_id == 2
