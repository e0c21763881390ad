#!/usr/bin/env python
"""Hand-assembled golden files for the segment-file reader: every byte below is written out by
hand from the reference's writer, not produced by this repo's writer, the oracle, or a msgpack library.

  golden_a.data, golden_b.data   a segment file as filefmt.writeSegment lays it out   /root/reference/src/filefmt.zig:143-178
  manifest                       manifest.write: a msgpack array of SegmentInfo       /root/reference/src/manifest.zig:41-47

Layout of a segment file (filefmt.zig:1-13): header | metadata | docs | zero padding to block_size | blocks |
one all-zero terminator block (filefmt.zig:113-115) | block index (LE u32 per block: hash of the block's last item,
filefmt.zig:117) | footer | LE u32 footer size.

The structs are msgpack maps keyed by FIELD INDEX (filefmt.zig:73-75, 87-89: `.as_map = .{ .key = .field_index }`),
SegmentInfo is a msgpack ARRAY [commit_id, merges, version] (segment.zig:64-66) whose optional version is nil when
null.  The encoder (msgpack.zig @ bef6671) is not vendored in the reference tree, so the integer widths it picks are
not pinned by anything we can read: the files use the shortest msgpack form of every integer (what every mainstream
encoder does); the reader is additionally tested with 64-bit-wide integers elsewhere (tests/test_segment_file.py).

The block is the one derived byte by byte in SURVEY.md Appendix B from block.zig:438-567:
  items (hash, id): (1,100) (1,200) (3,300) (4,400) (5,500), min_doc_id = 50 (the docs map's smallest key, a
  tombstone: filefmt.zig:244-250 takes the minimum over ALL keys of the docs map).
"""
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))


def crc64_xz(data: bytes) -> int:
    """std.hash.crc.Crc64Xz: reflected polynomial 0xC96C5795D7870F42, init and final xor all ones."""
    crc = 0xFFFFFFFFFFFFFFFF
    for b in data:
        crc ^= b
        for _ in range(8):
            crc = (crc >> 1) ^ 0xC96C5795D7870F42 if crc & 1 else crc >> 1
    return crc ^ 0xFFFFFFFFFFFFFFFF


BLOCK_SIZE = 512

# block.zig:545-564: header {min_hash u32, num_items u16, docids_offset u16} LE, hash control bytes, hash data,
# docid control bytes, docid data, zero padding  
BLOCK = bytes([
    0x01, 0x00, 0x00, 0x00,              # min_hash = 1
    0x05, 0x00,                          # num_items = 5
    0x05, 0x00,                          # docids_offset = 2 control + 3 data bytes of the hash column
    0x50, 0x01,                          # hash control: quad 0 deltas [0,0,2,1] -> 0124 codes [0,0,1,1]; quad 1 [1,0,0,0]
    0x02, 0x01, 0x01,                    # hash data
    0x40, 0x01,                          # docid control: quad 0 deltas [50,100,250,350] -> 1234 codes [0,0,0,1]; quad 1 [450,..]
    0x32, 0x64, 0xFA, 0x5E, 0x01,        # docid data quad 0: 50, 100, 250, 350 = 0x015E
    0xC2, 0x01, 0x00, 0x00, 0x00,        # docid data quad 1: 450 = 0x01C2, then the three padding values (1 byte each)
]).ljust(BLOCK_SIZE, b"\0")

DOCS = bytes([
    0x86,                                # map of 6: doc id -> alive?   (filefmt.zig:162 packer.writeMap(segment.docs))
    0x32, 0xC2,                          # 50: false  (tombstone; it sets min_doc_id)
    0x64, 0xC3,                          # 100: true
    0xCC, 0xC8, 0xC3,                    # 200: true  (uint8)
    0xCD, 0x01, 0x2C, 0xC3,              # 300: true  (uint16)
    0xCD, 0x01, 0x90, 0xC3,              # 400: true
    0xCD, 0x01, 0xF4, 0xC3,              # 500: true
])


def header(commit_id_bytes, merges_bytes, version_bytes):
    return bytes([
        0x85,                            # map of 5, keys = field indices (filefmt.zig:66-75)
        0x00, 0xCE, 0x53, 0x47, 0x4D, 0x31,   # 0 magic: 0x53474D31 "SGM1" (filefmt.zig:38)
        0x01, 0x93]) + commit_id_bytes + merges_bytes + version_bytes + bytes([   # 1 info: [commit_id, merges, version]
        0x02, 0xC3,                      # 2 has_metadata: true (filefmt.zig:157)
        0x03, 0xC3,                      # 3 has_docs: true
        0x04, 0xCD, 0x02, 0x00,          # 4 block_size: 512
    ])


def footer(block):
    crc = crc64_xz(block)                # over the data blocks only, not the terminator (filefmt.zig:120)
    body = bytes([
        0x84,                            # map of 4 (filefmt.zig:78-89)
        0x00, 0xCE, 0x31, 0x4D, 0x47, 0x53,   # 0 magic: @byteSwap(header_magic) (filefmt.zig:39)
        0x01, 0x05,                      # 1 num_items: 5
        0x02, 0x01,                      # 2 num_blocks: 1
        0x03, 0xCF]) + struct.pack(">Q", crc)   # 3 checksum: uint64
    return body + struct.pack("<I", len(body))   # footer size, LE u32 (filefmt.zig:176-177)


def segment_file(head, metadata):
    top = head + metadata + DOCS
    top += b"\0" * (-len(top) % BLOCK_SIZE)          # filefmt.zig:164-165
    index = struct.pack("<I", 5)                      # block_index[0] = hash of the last item (filefmt.zig:117)
    return top + BLOCK + b"\0" * BLOCK_SIZE + index + footer(BLOCK)


FILES = {
    # commit_id 7, merges 0, version null; empty metadata map
    "0000000000000007-00000000.data": segment_file(header(b"\x07", b"\x00", b"\xC0"), b"\x80"),
    # commit_id 9, merges 2, version 1000 (uint16); metadata {"name": "golden"}
    "0000000000000009-00000002.data": segment_file(header(b"\x09", b"\x02", b"\xCD\x03\xE8"),
                                                   b"\x81\xA4name\xA6golden"),
    # manifest.zig:41-47: [SegmentInfo, SegmentInfo]
    "manifest": bytes([0x92, 0x93, 0x07, 0x00, 0xC0, 0x93, 0x09, 0x02, 0xCD, 0x03, 0xE8]),
}

if __name__ == "__main__":
    for name, data in FILES.items():
        with open(os.path.join(HERE, name), "wb") as f:
            f.write(data)
        print(name, len(data), "bytes, crc64/xz of the file %016x" % crc64_xz(data))
