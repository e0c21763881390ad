"""Segment files (.data) and the manifest — SURVEY.md §8f row 1.

The reference writes / reads them in src/filefmt.zig:143-285 and src/manifest.zig with the un-vendored msgpack.zig, so
its exact byte choices are not pinned; what is pinned here:
  * CRC-64/XZ known-answer value (std.hash.crc.Crc64Xz);
  * the layout of filefmt.zig:1-13, assembled independently in this file with python-msgpack and a table-free CRC,
    is parsed by the product into exactly the segment it was made from;
  * the product's writer emits those same bytes (shortest msgpack encodings), and wider encodings are read too;
  * every rejection of readSegment: magic, block size, counts, checksum, truncation;
  * a snapshot compiled from the parsed file equals the one compiled from the in-memory segment.
"""
import os
import struct

import msgpack
import numpy as np
import pytest

from _helpers import pkg

HEADER_MAGIC, FOOTER_MAGIC = 0x53474D31, 0x314D4753


def crc64_xz(data: bytes) -> int:
    crc = 0xFFFFFFFFFFFFFFFF
    for b in data:
        crc ^= b
        for _ in range(8):
            crc = (crc >> 1) ^ 0xC96C5795D7870F42 if crc & 1 else crc >> 1
    return crc ^ 0xFFFFFFFFFFFFFFFF


def make_segment(n_docs=300, H=40, vocab=5000, seed=1, commit_id=7, merges=2):
    rng = np.random.default_rng(seed)
    ids = np.sort(rng.choice(np.arange(10, 10 + 4 * n_docs), size=n_docs, replace=False)).astype(np.uint64)
    hashes = rng.integers(0, vocab, size=(n_docs, H)).astype(np.uint64)
    items = np.sort(((hashes << np.uint64(32)) | ids[:, None]).reshape(-1))
    doc_ids = np.concatenate([ids.astype(np.uint32), np.array([3, 5000], dtype=np.uint32)])      # + two tombstones
    doc_alive = np.concatenate([np.ones(n_docs, np.uint8), np.zeros(2, np.uint8)])
    return pkg.FileSegment.from_items(items, doc_ids, doc_alive, commit_id=commit_id, merges=merges), items


def python_file_bytes(seg, version=None, metadata=None, wide=False):
    """filefmt.zig:143-178 assembled with python-msgpack (an independent encoder)."""
    bs, nb = seg.block_size, seg.num_blocks
    body = bytes(seg.blocks[:nb * bs])
    if wide:   # every integer as a 64-bit msgpack uint: a reader must not depend on the encoder's width choices
        u = lambda v: b"\xcf" + struct.pack(">Q", v)
        head = b"\x85" + u(0) + u(HEADER_MAGIC) + u(1) + b"\x93" + u(seg.commit_id) + u(seg.merges) + \
            (u(version) if version is not None else b"\xc0") + u(2) + b"\xc3" + u(3) + b"\xc3" + u(4) + u(bs)
        head += b"\xde" + struct.pack(">H", len(metadata or {})) + b"".join(
            msgpack.packb(k) + msgpack.packb(v) for k, v in (metadata or {}).items())
        head += b"\xdf" + struct.pack(">I", len(seg.doc_ids)) + b"".join(
            u(int(i)) + (b"\xc3" if a else b"\xc2") for i, a in zip(seg.doc_ids, seg.doc_alive))
    else:
        head = msgpack.packb({0: HEADER_MAGIC, 1: [seg.commit_id, seg.merges, version], 2: True, 3: True, 4: bs})
        head += msgpack.packb(metadata or {})
        head += msgpack.packb({int(i): bool(a) for i, a in zip(seg.doc_ids, seg.doc_alive)})
    head += b"\0" * (-len(head) % bs)
    n_items = sum(struct.unpack_from("<H", body, b * bs + 4)[0] for b in range(nb))
    index = np.asarray(seg.block_index, dtype="<u4").tobytes()
    footer = msgpack.packb({0: FOOTER_MAGIC, 1: n_items, 2: nb, 3: crc64_xz(body)})
    return head + body + b"\0" * bs + index + footer + struct.pack("<I", len(footer))


def assert_same_segment(a, b):
    assert (a.commit_id, a.merges, a.min_doc_id, a.block_size, a.num_blocks) == \
        (b.commit_id, b.merges, b.min_doc_id, b.block_size, b.num_blocks)
    n = a.num_blocks * a.block_size
    assert np.array_equal(a.blocks[:n], b.blocks[:n]) and np.array_equal(a.block_index, b.block_index)
    assert dict(zip(a.doc_ids.tolist(), a.doc_alive.tolist())) == dict(zip(b.doc_ids.tolist(), b.doc_alive.tolist()))


def test_crc64_xz_known_answer():
    assert pkg.lib().fpx_crc64_xz(b"123456789", 9) == 0x995DC9BBDF1939FA     # CRC-64/XZ check value
    assert crc64_xz(b"123456789") == 0x995DC9BBDF1939FA
    blob = os.urandom(4097)
    assert pkg.lib().fpx_crc64_xz(blob, len(blob)) == crc64_xz(blob)


def test_segment_file_name():
    assert pkg.segment_file_name(0x1234, 5) == "0000000000001234-00000005.data"          # filefmt.zig:36
    assert pkg.segment_file_name(2 ** 64 - 1, 0xFFFFFFFF) == "ffffffffffffffff-ffffffff.data"


@pytest.mark.parametrize("version", [None, 0, 123456789012])
def test_python_assembled_file_parses_to_the_same_segment(version):
    seg, items = make_segment()
    data = python_file_bytes(seg, version=version, metadata={"name": "main", "x" * 40: "y" * 300})
    f = pkg.SegmentFile.parse(data)
    assert f.info == (7, 2, version) and f.num_items == len(items)
    assert f.metadata == {b"name": b"main", b"x" * 40: b"y" * 300}
    assert_same_segment(f.segment, seg)
    assert f.segment.min_doc_id == 3                                                      # tombstones count (filefmt.zig:244-250)


def test_writer_emits_the_python_assembled_bytes_and_wide_encodings_are_read():
    seg, items = make_segment(seed=3)
    for version in (None, 99):
        ours = pkg.segment_file_bytes(seg, version=version)
        assert ours == python_file_bytes(seg, version=version)
        assert_same_segment(pkg.SegmentFile.parse(ours).segment, seg)
    wide = python_file_bytes(seg, version=5, metadata={"k": "v"}, wide=True)
    f = pkg.SegmentFile.parse(wide)
    assert f.info == (7, 2, 5) and f.metadata == {b"k": b"v"}
    assert_same_segment(f.segment, seg)


def test_empty_segment_and_read_from_disk(tmp_path):
    empty = pkg.FileSegment.from_items(np.zeros(0, np.uint64), np.array([9], np.uint32), np.array([0], np.uint8), commit_id=1)
    data = pkg.segment_file_bytes(empty)
    f = pkg.SegmentFile.parse(data)
    assert f.segment.num_blocks == 0 and f.num_items == 0 and f.segment.doc_ids.tolist() == [9]
    seg, _ = make_segment(seed=5, commit_id=0xABC, merges=1)
    (tmp_path / pkg.segment_file_name(0xABC, 1)).write_bytes(pkg.segment_file_bytes(seg))
    (tmp_path / pkg.segment_file_name(1, 0)).write_bytes(data)
    (tmp_path / "manifest").write_bytes(msgpack.packb([[1, 0, None], [0xABC, 1, 77]]))      # manifest.zig:44-46
    files = pkg.open_index_dir(str(tmp_path))
    assert [x.info[:2] for x in files] == [(1, 0), (0xABC, 1)]
    assert_same_segment(files[1].segment, seg)
    with pytest.raises(pkg.FpxError):
        pkg.SegmentFile.read(str(tmp_path / "missing.data"))


def test_manifest():
    assert pkg.parse_manifest(b"") == []                                                  # manifest.zig:23: empty file
    assert pkg.parse_manifest(msgpack.packb([])) == []
    assert pkg.parse_manifest(msgpack.packb([[1, 0, None], [2, 3, 2 ** 40], [2 ** 33, 0]])) == \
        [(1, 0, None), (2, 3, 2 ** 40), (2 ** 33, 0, None)]
    for bad in (b"\x91", msgpack.packb({1: 2}), msgpack.packb([[1]]), msgpack.packb([["a", 0, None]])):
        with pytest.raises(pkg.FpxError):
            pkg.parse_manifest(bad)


def test_rejections():
    seg, _ = make_segment(seed=9)
    good = pkg.segment_file_bytes(seg)
    pkg.SegmentFile.parse(good)
    bs = seg.block_size
    first_block = (good.index(b"\0" * 8) // bs + 1) * bs      # blocks start at the first block_size boundary after the header

    def rejected(data, needle):
        with pytest.raises(pkg.FpxError) as e:
            pkg.SegmentFile.parse(bytes(data))
        assert e.value.status == 3 and needle in str(e.value), str(e.value)       # FPX_INVALID_SEGMENT

    flipped = bytearray(good)
    flipped[first_block + 40] ^= 0x10
    rejected(flipped, "checksum mismatch")                                        # filefmt.zig:284
    rejected(good[:len(good) // 2], "")                                           # truncated
    rejected(good[:-6], "")
    bad_magic = bytearray(good)
    assert bad_magic[2:7] == b"\xce" + struct.pack(">I", HEADER_MAGIC)
    bad_magic[6] ^= 1
    rejected(bad_magic, "bad header magic")                                       # filefmt.zig:236
    rejected(python_file_bytes(seg).replace(msgpack.packb(512), msgpack.packb(32), 1), "block size")   # filefmt.zig:237
    # footer that disagrees with the blocks (filefmt.zig:283)
    body_end = len(good) - 4 - struct.unpack("<I", good[-4:])[0]
    footer = msgpack.unpackb(good[body_end:-4], strict_map_key=False)
    footer[1] += 1
    fb = msgpack.packb(footer)
    rejected(good[:body_end] + fb + struct.pack("<I", len(fb)), "footer counts")
    rejected(b"", "")
    rejected(b"\x00" * 100, "")


def test_snapshot_from_file_equals_snapshot_from_memory():
    seg, _ = make_segment(n_docs=500, H=60, seed=11)
    f = pkg.SegmentFile.parse(pkg.segment_file_bytes(seg))
    ctx = pkg.Context(host_only=True)
    a, b = pkg.SnapshotBuilder(ctx), pkg.SnapshotBuilder(ctx)
    a.add_file_segment(seg)
    b.add_file_segment(f.segment)
    ca, cb = a.csr(), b.csr()
    assert all(np.array_equal(x, y) for x, y in zip(ca, cb)) and len(ca[0]) > 0
    a.abort(), b.abort()
    ctx.close()
