// fpx_filefmt.h — fpindex segment files (.data) and the manifest, host side.
//
// Follows src/filefmt.zig:1-13 (layout), :66-87 (header / footer structs, msgpack maps keyed by field index),
// :143-206 (writeSegment), :209-285 (readSegment) and src/manifest.zig:17-39 (msgpack array of SegmentInfo);
// SegmentInfo is a msgpack array [commit_id, merges, version?] (src/segment.zig:23-66).
//
//   header (msgpack) | metadata map str->str | docs map u32->bool | zero pad to block_size | blocks |
//   all-zero terminator block | block index (u32 LE per block) | footer (msgpack) | footer size (u32 LE)
//
// The reader accepts any valid msgpack encoding of these values (the reference's encoder, msgpack.zig, is not
// vendored, so its exact width choices are not pinned); the writer emits the shortest encodings.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace fpx {

constexpr uint32_t kHeaderMagic = 0x53474D31u; // "SGM1", filefmt.zig:39
constexpr uint32_t kFooterMagic = 0x314D4753u; // byte-swapped, filefmt.zig:40

// ---- CRC-64/XZ (std.hash.crc.Crc64Xz: reflected, poly 0x42F0E1EBA9EA3693, init/xorout all ones) ----
struct Crc64Xz {
    uint64_t table[256];
    Crc64Xz() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint64_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0xC96C5795D7870F42ull : c >> 1;
            table[i] = c;
        }
    }
    uint64_t update(uint64_t crc, const uint8_t *p, size_t n) const {
        for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
        return crc;
    }
    uint64_t of(const uint8_t *p, size_t n) const { return ~update(~0ull, p, n); }
};
inline const Crc64Xz &crc64xz() {
    static const Crc64Xz c;
    return c;
}

// ---- a small msgpack reader: the subset the segment files use, any width ----
struct MsgpackReader {
    const uint8_t *p, *end;
    bool ok = true;
    MsgpackReader(const uint8_t *b, size_t n) : p(b), end(b + n) {}

    bool need(size_t n) {
        if ((size_t)(end - p) < n) ok = false;
        return ok;
    }
    uint64_t be(size_t n) {
        uint64_t v = 0;
        if (!need(n)) return 0;
        for (size_t i = 0; i < n; ++i) v = (v << 8) | p[i];
        p += n;
        return v;
    }
    uint8_t peek() { return need(1) ? *p : 0xC1; }
    bool read_nil() {
        if (peek() == 0xC0) {
            ++p;
            return true;
        }
        return false;
    }
    bool read_bool(bool &v) {
        const uint8_t t = peek();
        if (t != 0xC2 && t != 0xC3) return ok = false;
        v = t == 0xC3;
        ++p;
        return true;
    }
    bool read_uint(uint64_t &v) { // accepts non-negative signed encodings too
        const uint8_t t = peek();
        if (!ok) return false;
        ++p;
        if (t <= 0x7F) v = t;
        else if (t == 0xCC) v = be(1);
        else if (t == 0xCD) v = be(2);
        else if (t == 0xCE) v = be(4);
        else if (t == 0xCF) v = be(8);
        else if (t == 0xD0 || t == 0xD1 || t == 0xD2 || t == 0xD3) {
            const size_t n = (size_t)1 << (t - 0xD0);
            const uint64_t raw = be(n);
            const uint64_t sign = 1ull << (8 * n - 1);
            if (raw & sign) return ok = false; // negative
            v = raw;
        } else
            return ok = false;
        return ok;
    }
    bool read_len(uint8_t fix_lo, uint8_t fix_mask, uint8_t t16, uint8_t t32, uint64_t &n) {
        const uint8_t t = peek();
        if (!ok) return false;
        ++p;
        if ((t & fix_mask) == fix_lo) n = t & (uint8_t)~fix_mask;
        else if (t == t16) n = be(2);
        else if (t == t32) n = be(4);
        else
            return ok = false;
        return ok;
    }
    bool read_map(uint64_t &n) { return read_len(0x80, 0xF0, 0xDE, 0xDF, n); }
    bool read_array(uint64_t &n) { return read_len(0x90, 0xF0, 0xDC, 0xDD, n); }
    bool read_str(std::string &s) {
        const uint8_t t = peek();
        if (!ok) return false;
        ++p;
        uint64_t n = 0;
        if ((t & 0xE0) == 0xA0) n = t & 0x1F;
        else if (t == 0xD9 || t == 0xC4) n = be(1); // str8 / bin8
        else if (t == 0xDA || t == 0xC5) n = be(2);
        else if (t == 0xDB || t == 0xC6) n = be(4);
        else
            return ok = false;
        if (!need(n)) return false;
        s.assign(reinterpret_cast<const char *>(p), (size_t)n);
        p += n;
        return true;
    }
    bool skip() { // any value of the subset
        const uint8_t t = peek();
        if (!ok) return false;
        uint64_t n, u;
        bool b;
        std::string s;
        if (t == 0xC0) return ++p, true;
        if (t == 0xC2 || t == 0xC3) return read_bool(b);
        if (t <= 0x7F || (t >= 0xCC && t <= 0xD3)) {
            if (t >= 0xD0) { // signed: skip raw
                ++p;
                be((size_t)1 << (t - 0xD0));
                return ok;
            }
            return read_uint(u);
        }
        if (t >= 0xE0) return ++p, true; // negative fixint
        if ((t & 0xE0) == 0xA0 || t == 0xD9 || t == 0xDA || t == 0xDB || t == 0xC4 || t == 0xC5 || t == 0xC6) return read_str(s);
        if ((t & 0xF0) == 0x90 || t == 0xDC || t == 0xDD) {
            if (!read_array(n)) return false;
            for (uint64_t i = 0; i < n && ok; ++i) skip();
            return ok;
        }
        if ((t & 0xF0) == 0x80 || t == 0xDE || t == 0xDF) {
            if (!read_map(n)) return false;
            for (uint64_t i = 0; i < 2 * n && ok; ++i) skip();
            return ok;
        }
        return ok = false;
    }
};

struct MsgpackWriter {
    std::vector<uint8_t> &out;
    explicit MsgpackWriter(std::vector<uint8_t> &o) : out(o) {}
    void be(uint64_t v, int n) {
        for (int i = n - 1; i >= 0; --i) out.push_back((uint8_t)(v >> (8 * i)));
    }
    void nil() { out.push_back(0xC0); }
    void boolean(bool v) { out.push_back(v ? 0xC3 : 0xC2); }
    void uint(uint64_t v) {
        if (v <= 0x7F) out.push_back((uint8_t)v);
        else if (v <= 0xFF) out.push_back(0xCC), be(v, 1);
        else if (v <= 0xFFFF) out.push_back(0xCD), be(v, 2);
        else if (v <= 0xFFFFFFFFull) out.push_back(0xCE), be(v, 4);
        else out.push_back(0xCF), be(v, 8);
    }
    void map(uint64_t n) {
        if (n <= 15) out.push_back((uint8_t)(0x80 | n));
        else if (n <= 0xFFFF) out.push_back(0xDE), be(n, 2);
        else out.push_back(0xDF), be(n, 4);
    }
    void array(uint64_t n) {
        if (n <= 15) out.push_back((uint8_t)(0x90 | n));
        else if (n <= 0xFFFF) out.push_back(0xDC), be(n, 2);
        else out.push_back(0xDD), be(n, 4);
    }
    void str(const std::string &s) {
        const uint64_t n = s.size();
        if (n <= 31) out.push_back((uint8_t)(0xA0 | n));
        else if (n <= 0xFF) out.push_back(0xD9), be(n, 1);
        else if (n <= 0xFFFF) out.push_back(0xDA), be(n, 2);
        else out.push_back(0xDB), be(n, 4);
        out.insert(out.end(), s.begin(), s.end());
    }
};

struct SegmentInfoHost { // segment.zig:23-26
    uint64_t commit_id = 0, merges = 0, version = 0;
    bool has_version = false;
};

inline bool read_segment_info(MsgpackReader &r, SegmentInfoHost &info) { // msgpack array, segment.zig:64-66
    uint64_t n = 0;
    if (!r.read_array(n) || n < 2) return r.ok = false;
    if (!r.read_uint(info.commit_id) || !r.read_uint(info.merges)) return false;
    info.has_version = false;
    if (n >= 3) {
        if (!r.read_nil()) {
            if (!r.read_uint(info.version)) return false;
            info.has_version = true;
        }
    }
    for (uint64_t i = 3; i < n && r.ok; ++i) r.skip();
    return r.ok;
}
inline void write_segment_info(MsgpackWriter &w, const SegmentInfoHost &info) {
    w.array(3);
    w.uint(info.commit_id);
    w.uint(info.merges);
    if (info.has_version) w.uint(info.version);
    else w.nil();
}

// A parsed segment file: views into `data` plus the decoded docs map and metadata.
struct SegmentFile {
    std::vector<uint8_t> owned; // the file's bytes when we own them
    const uint8_t *data = nullptr;
    size_t size = 0;
    SegmentInfoHost info;
    uint32_t block_size = 0;
    const uint8_t *blocks = nullptr; // num_blocks blocks, followed by the terminator block
    uint64_t num_blocks = 0, num_items = 0;
    std::vector<uint32_t> block_index; // copied: the file offset need not be 4-byte aligned in a caller's buffer
    std::vector<uint32_t> doc_ids;
    std::vector<uint8_t> doc_alive;
    uint32_t min_doc_id = 0, max_doc_id = 0;
    std::vector<std::pair<std::string, std::string>> metadata;
    std::string error;
    bool checksum_mismatch = false;

    bool fail(const char *m) {
        error = m;
        return false;
    }

    // filefmt.zig:209-285 readSegment, on bytes already in memory
    bool parse(const uint8_t *d, size_t n) {
        data = d;
        size = n;
        MsgpackReader r(d, n);
        uint64_t nf = 0, magic = 0, bs = 0;
        bool has_metadata = false, has_docs = false, seen[5] = {false, false, false, false, false};
        if (!r.read_map(nf)) return fail("header is not a msgpack map");
        for (uint64_t i = 0; i < nf; ++i) { // keys are field indices, filefmt.zig:73-75
            uint64_t key = 0;
            if (!r.read_uint(key)) return fail("header key is not an integer");
            bool good = true;
            switch (key) {
            case 0: good = r.read_uint(magic); break;
            case 1: good = read_segment_info(r, info); break;
            case 2: good = r.read_bool(has_metadata); break;
            case 3: good = r.read_bool(has_docs); break;
            case 4: good = r.read_uint(bs); break;
            default: good = r.skip();
            }
            if (!good) return fail("malformed header field");
            if (key < 5) seen[key] = true;
        }
        for (bool sn : seen)
            if (!sn) return fail("header field missing");
        if (magic != kHeaderMagic) return fail("bad header magic");            // filefmt.zig:236
        if (bs < 64 || bs > 4096) return fail("block size out of range");      // filefmt.zig:237, block.zig:41-42
        block_size = (uint32_t)bs;
        if (has_metadata) {
            uint64_t m = 0;
            if (!r.read_map(m)) return fail("malformed metadata map");
            for (uint64_t i = 0; i < m; ++i) {
                std::string k, v;
                if (!r.read_str(k) || !r.read_str(v)) return fail("malformed metadata entry");
                metadata.emplace_back(std::move(k), std::move(v));
            }
        }
        if (has_docs) {
            uint64_t m = 0;
            if (!r.read_map(m)) return fail("malformed docs map");
            if (m > (uint64_t)(r.end - r.p)) return fail("malformed docs map");
            doc_ids.reserve((size_t)m);
            doc_alive.reserve((size_t)m);
            for (uint64_t i = 0; i < m; ++i) {
                uint64_t id = 0;
                bool alive = false;
                if (!r.read_uint(id) || id > 0xFFFFFFFFull || !r.read_bool(alive)) return fail("malformed docs entry");
                doc_ids.push_back((uint32_t)id);
                doc_alive.push_back(alive ? 1 : 0);
            }
        }
        min_doc_id = max_doc_id = 0; // filefmt.zig:244-250
        for (uint32_t id : doc_ids) {
            if (min_doc_id == 0 || id < min_doc_id) min_doc_id = id;
            if (max_doc_id == 0 || id > max_doc_id) max_doc_id = id;
        }
        const size_t hdr_end = (size_t)(r.p - d);
        const size_t blocks_start = (hdr_end + block_size - 1) / block_size * block_size; // filefmt.zig:253-254
        uint64_t crc = ~0ull;
        size_t ptr = blocks_start;
        num_blocks = num_items = 0;
        bool terminated = false;
        while (ptr + block_size <= n) { // filefmt.zig:260-268
            const uint8_t *blk = d + ptr;
            ptr += block_size;
            const uint32_t items = (uint32_t)blk[4] | ((uint32_t)blk[5] << 8); // block.zig:46-50 header.num_items
            if (items == 0) {
                terminated = true;
                break;
            }
            num_items += items;
            num_blocks += 1;
            crc = crc64xz().update(crc, blk, block_size);
        }
        if (!terminated) return fail("blocks are not terminated by an empty block");
        if (blocks_start > n) return fail("truncated file");
        blocks = d + blocks_start;
        const size_t index_start = ptr, index_end = index_start + (size_t)num_blocks * 4;
        if (index_end > n) return fail("truncated block index"); // filefmt.zig:275
        block_index.resize((size_t)num_blocks);
        for (size_t i = 0; i < (size_t)num_blocks; ++i) {
            const uint8_t *q = d + index_start + 4 * i;
            block_index[i] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
        }
        MsgpackReader fr(d + index_end, n - index_end);
        uint64_t fn = 0, fmagic = 0, f_items = 0, f_blocks = 0, f_crc = 0;
        bool fseen[4] = {false, false, false, false};
        if (!fr.read_map(fn)) return fail("footer is not a msgpack map");
        for (uint64_t i = 0; i < fn; ++i) {
            uint64_t key = 0;
            if (!fr.read_uint(key)) return fail("footer key is not an integer");
            bool good = true;
            switch (key) {
            case 0: good = fr.read_uint(fmagic); break;
            case 1: good = fr.read_uint(f_items); break;
            case 2: good = fr.read_uint(f_blocks); break;
            case 3: good = fr.read_uint(f_crc); break;
            default: good = fr.skip();
            }
            if (!good) return fail("malformed footer field");
            if (key < 4) fseen[key] = true;
        }
        for (bool sn : fseen)
            if (!sn) return fail("footer field missing");
        if (fmagic != kFooterMagic) return fail("bad footer magic");                                     // filefmt.zig:282
        if (f_items != num_items || f_blocks != num_blocks) return fail("footer counts do not match");   // :283
        if (f_crc != ~crc) {                                                                              // :284
            checksum_mismatch = true;
            return fail("checksum mismatch");
        }
        return true;
    }
};

// filefmt.zig:143-178: header, (empty) metadata, docs, pad, blocks + terminator, block index, footer, footer size
inline void serialize_segment_file(const SegmentInfoHost &info, uint32_t block_size, const uint8_t *blocks, uint64_t num_blocks,
                                   const uint32_t *block_index, const uint32_t *doc_ids, const uint8_t *doc_alive,
                                   uint64_t n_docs, const std::vector<std::pair<std::string, std::string>> &metadata,
                                   std::vector<uint8_t> &out) {
    out.clear();
    MsgpackWriter w(out);
    w.map(5);
    w.uint(0), w.uint(kHeaderMagic);
    w.uint(1), write_segment_info(w, info);
    w.uint(2), w.boolean(true);
    w.uint(3), w.boolean(true);
    w.uint(4), w.uint(block_size);
    w.map(metadata.size());
    for (const auto &kv : metadata) w.str(kv.first), w.str(kv.second);
    w.map(n_docs);
    for (uint64_t i = 0; i < n_docs; ++i) w.uint(doc_ids[i]), w.boolean(doc_alive[i] != 0);
    const size_t rem = out.size() % block_size;
    if (rem) out.insert(out.end(), block_size - rem, 0);
    out.insert(out.end(), blocks, blocks + (size_t)num_blocks * block_size);
    out.insert(out.end(), block_size, 0); // empty terminator block, filefmt.zig:113-115
    uint64_t n_items = 0;
    for (uint64_t b = 0; b < num_blocks; ++b) {
        const uint8_t *blk = blocks + (size_t)b * block_size;
        n_items += (uint32_t)blk[4] | ((uint32_t)blk[5] << 8);
        const uint32_t mh = block_index[b];
        out.push_back((uint8_t)mh), out.push_back((uint8_t)(mh >> 8)), out.push_back((uint8_t)(mh >> 16)), out.push_back((uint8_t)(mh >> 24));
    }
    const size_t footer_start = out.size();
    w.map(4);
    w.uint(0), w.uint(kFooterMagic);
    w.uint(1), w.uint(n_items);
    w.uint(2), w.uint(num_blocks);
    w.uint(3), w.uint(crc64xz().of(blocks, (size_t)num_blocks * block_size));
    const uint32_t fs = (uint32_t)(out.size() - footer_start);
    out.push_back((uint8_t)fs), out.push_back((uint8_t)(fs >> 8)), out.push_back((uint8_t)(fs >> 16)), out.push_back((uint8_t)(fs >> 24));
}

// manifest.zig:17-39: msgpack array of SegmentInfo; empty input = no segments
inline bool parse_manifest(const uint8_t *d, size_t n, std::vector<SegmentInfoHost> &out) {
    out.clear();
    if (n == 0) return true;
    MsgpackReader r(d, n);
    uint64_t cnt = 0;
    if (!r.read_array(cnt)) return false;
    for (uint64_t i = 0; i < cnt; ++i) {
        SegmentInfoHost s;
        if (!read_segment_info(r, s)) return false;
        out.push_back(s);
    }
    return r.ok;
}

} // namespace fpx
