// fpindex_oracle.cpp — CPU ORACLE for the fpindex `_search` hot path.
//
// TEST INFRASTRUCTURE, NOT PRODUCT CODE (see fpindex_oracle.h).  A restatement, in
// C++17, of the algorithm in acoustid-index's Zig sources.  Every section cites
// the reference file:line it follows (paths relative to the reference root).
// It deliberately keeps the reference's COST STRUCTURE (block-compressed segments
// re-decoded per query with a pshufb StreamVByte decoder, one hash-map update per
// matched posting, full sort of the candidates) because it doubles as the timed
// CPU baseline ("port") in bench.py.
//
// Parity pinning: checked against the reference's own known-answer tests in
// tests/test_oracle_kat.py; "parity unpinned" items are listed in DESIGN.md.

#include "fpindex_oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#if defined(__SSSE3__)
#include <tmmintrin.h>
#define ORC_SSSE3 1
#else
#define ORC_SSSE3 0
#endif

namespace {

// ---------------------------------------------------------------------------------------
// StreamVByte  (src/streamvbyte.zig)
// ---------------------------------------------------------------------------------------

constexpr size_t kSimdPadding = 16; // streamvbyte.zig:5

enum Variant { V0124 = 0, V1234 = 1, V0124M1 = 2 }; // streamvbyte.zig:63-67

struct SvbTables {
    // streamvbyte.zig:76-211 — per control byte: pshufb mask and total data length.
    alignas(16) uint8_t shuf0124[256][16];
    alignas(16) uint8_t shuf1234[256][16];
    uint8_t len0124[256];
    uint8_t len1234[256];
    SvbTables() {
        static const uint8_t w0124[4] = {0, 1, 2, 4};
        static const uint8_t w1234[4] = {1, 2, 3, 4};
        for (int c = 0; c < 256; ++c) {
            for (int v = 0; v < 2; ++v) {
                const uint8_t *w = v ? w1234 : w0124;
                uint8_t(*mask)[16] = v ? shuf1234 : shuf0124;
                uint8_t off = 0;
                for (int b = 0; b < 16; ++b) mask[c][b] = 0x80; // high bit => output zero byte
                for (int i = 0; i < 4; ++i) {
                    uint8_t n = w[(c >> (2 * i)) & 3];
                    for (uint8_t j = 0; j < n; ++j) mask[c][4 * i + j] = (uint8_t)(off + j);
                    off = (uint8_t)(off + n);
                }
                (v ? len1234 : len0124)[c] = off;
            }
        }
    }
};
const SvbTables g_tab;

inline const uint8_t *length_table(int variant) {
    return variant == V1234 ? g_tab.len1234 : g_tab.len0124; // streamvbyte.zig:351-355
}

// streamvbyte.zig:216-247 svbDecodeQuadBase.  Reads 16 bytes at `in`.
inline size_t decode_quad(int variant, uint8_t control, const uint8_t *in, uint32_t out[4]) {
#if ORC_SSSE3
    const uint8_t(*mask)[16] = variant == V1234 ? g_tab.shuf1234 : g_tab.shuf0124;
    __m128i data = _mm_loadu_si128(reinterpret_cast<const __m128i *>(in));
    __m128i m = _mm_load_si128(reinterpret_cast<const __m128i *>(mask[control]));
    __m128i r = _mm_shuffle_epi8(data, m);
    if (variant == V0124M1) r = _mm_add_epi32(r, _mm_set1_epi32(1));
    _mm_storeu_si128(reinterpret_cast<__m128i *>(out), r);
#else
    // scalar equivalent of the pshufb path (streamvbyte.zig:47-59)
    const uint8_t(*mask)[16] = variant == V1234 ? g_tab.shuf1234 : g_tab.shuf0124;
    uint8_t bytes[16];
    for (int b = 0; b < 16; ++b) {
        uint8_t mm = mask[control][b];
        bytes[b] = (mm & 0x80) ? 0 : in[mm & 0x0F];
    }
    for (int i = 0; i < 4; ++i) {
        uint32_t v = (uint32_t)bytes[4 * i] | ((uint32_t)bytes[4 * i + 1] << 8) |
                     ((uint32_t)bytes[4 * i + 2] << 16) | ((uint32_t)bytes[4 * i + 3] << 24);
        out[i] = v + (variant == V0124M1 ? 1u : 0u);
    }
#endif
    return length_table(variant)[control];
}

// streamvbyte.zig:264-283 svbDecodeQuadWithDelta: in-register prefix sum + carry.
inline size_t decode_quad_delta(int variant, uint8_t control, const uint8_t *in, uint32_t out[4],
                                uint32_t carry) {
#if ORC_SSSE3
    const uint8_t(*mask)[16] = variant == V1234 ? g_tab.shuf1234 : g_tab.shuf0124;
    __m128i data = _mm_loadu_si128(reinterpret_cast<const __m128i *>(in));
    __m128i v = _mm_shuffle_epi8(data, _mm_load_si128(reinterpret_cast<const __m128i *>(mask[control])));
    if (variant == V0124M1) v = _mm_add_epi32(v, _mm_set1_epi32(1));
    v = _mm_add_epi32(v, _mm_slli_si128(v, 4));
    v = _mm_add_epi32(v, _mm_slli_si128(v, 8));
    v = _mm_add_epi32(v, _mm_set1_epi32((int)carry));
    _mm_storeu_si128(reinterpret_cast<__m128i *>(out), v);
    return length_table(variant)[control];
#else
    uint32_t t[4];
    size_t n = decode_quad(variant, control, in, t);
    uint32_t acc = carry;
    for (int i = 0; i < 4; ++i) {
        acc += t[i];
        out[i] = acc;
    }
    return n;
#endif
}

// streamvbyte.zig:287-339 svbDeltaDecodeInPlace (result is a plain running sum).
inline void delta_decode_in_place(uint32_t *data, size_t n, uint32_t first_value) {
    if (n == 0) return;
    data[0] += first_value;
    for (size_t i = 1; i < n; ++i) data[i] += data[i - 1];
}

// streamvbyte.zig:341-412 decodeValues: decode the quads covering [start_item,end_item).
void decode_values(size_t total_items, size_t start_item, size_t end_item, const uint8_t *in,
                   uint32_t *out, int variant, bool delta, uint32_t first_value) {
    const uint8_t *len = length_table(variant);
    const size_t start_quad = start_item / 4;
    const size_t end_quad = (end_item + 3) / 4;
    const size_t total_quads = (total_items + 3) / 4;
    size_t data_offset = total_quads;
    for (size_t q = 0; q < start_quad; ++q) data_offset += len[in[q]]; // :361-365
    const uint8_t *ctrl = in + start_quad;
    const uint8_t *data = in + data_offset;
    uint32_t *o = out + start_quad * 4;
    size_t remaining = end_quad - start_quad;
    if (delta) {
        uint32_t carry = first_value;
        while (remaining--) {
            data += decode_quad_delta(variant, *ctrl++, data, o, carry);
            carry = o[3];
            o += 4;
        }
    } else {
        while (remaining--) {
            data += decode_quad(variant, *ctrl++, data, o);
            o += 4;
        }
    }
}

// streamvbyte.zig:418-480 encoders.
inline size_t encode_value_0124(uint32_t v, uint8_t *out, uint8_t *ctrl, int idx) {
    if (v == 0) return 0;
    if (v < (1u << 8)) {
        out[0] = (uint8_t)v;
        *ctrl |= (uint8_t)(1u << (2 * idx));
        return 1;
    }
    if (v < (1u << 16)) {
        out[0] = (uint8_t)v;
        out[1] = (uint8_t)(v >> 8);
        *ctrl |= (uint8_t)(2u << (2 * idx));
        return 2;
    }
    out[0] = (uint8_t)v;
    out[1] = (uint8_t)(v >> 8);
    out[2] = (uint8_t)(v >> 16);
    out[3] = (uint8_t)(v >> 24);
    *ctrl |= (uint8_t)(3u << (2 * idx));
    return 4;
}
inline size_t encode_value_1234(uint32_t v, uint8_t *out, uint8_t *ctrl, int idx) {
    size_t n = v < (1u << 8) ? 1 : v < (1u << 16) ? 2 : v < (1u << 24) ? 3 : 4;
    for (size_t j = 0; j < n; ++j) out[j] = (uint8_t)(v >> (8 * j));
    *ctrl |= (uint8_t)((n - 1) << (2 * idx));
    return n;
}
inline size_t encode_quad_0124(const uint32_t in[4], uint8_t *out, uint8_t *ctrl) {
    *ctrl = 0;
    size_t n = 0;
    for (int i = 0; i < 4; ++i) n += encode_value_0124(in[i], out + n, ctrl, i);
    return n;
}
inline size_t encode_quad_1234(const uint32_t in[4], uint8_t *out, uint8_t *ctrl) {
    *ctrl = 0;
    size_t n = 0;
    for (int i = 0; i < 4; ++i) n += encode_value_1234(in[i], out + n, ctrl, i);
    return n;
}

// ---------------------------------------------------------------------------------------
// Item  (src/segment.zig:87-106): packed u64, id in the low half, hash in the high half.
// ---------------------------------------------------------------------------------------
inline uint32_t item_hash(uint64_t it) { return (uint32_t)(it >> 32); }
inline uint32_t item_id(uint64_t it) { return (uint32_t)it; }
inline uint64_t make_item(uint32_t hash, uint32_t id) { return ((uint64_t)hash << 32) | id; }

// ---------------------------------------------------------------------------------------
// Block codec  (src/block.zig)
// ---------------------------------------------------------------------------------------
constexpr size_t kMaxBlockSize = 4096;                 // block.zig:42
constexpr size_t kMaxItemsPerBlock = kMaxBlockSize / 2; // block.zig:43
constexpr size_t kBlockHeaderSize = 8;                 // block.zig:44

struct BlockHeader { // block.zig:46-50 (little-endian, extern struct)
    uint32_t min_hash;
    uint16_t num_items;
    uint16_t docids_offset;
};
inline BlockHeader read_header(const uint8_t *p) {
    BlockHeader h;
    std::memcpy(&h, p, sizeof h);
    return h;
}

// block.zig:420-567 BlockEncoder.
struct BlockEncoder {
    uint16_t num_items = 0;
    uint32_t last_hash = 0, last_docid = 0;
    uint8_t hash_data[kMaxBlockSize + 16];
    uint8_t hash_ctrl[kMaxBlockSize];
    uint8_t doc_data[kMaxBlockSize + 16];
    uint8_t doc_ctrl[kMaxBlockSize];
    size_t n_hash_data = 0, n_hash_ctrl = 0, n_doc_data = 0, n_doc_ctrl = 0;

    // block.zig:438-495 encodeChunk: returns false on BlockFull.
    bool encode_chunk(const uint64_t *items, size_t n, uint32_t min_doc_id, size_t block_size) {
        uint32_t dh[4] = {0, 0, 0, 0}, dd[4] = {0, 0, 0, 0};
        for (size_t i = 0; i < n; ++i) {
            uint32_t h = item_hash(items[i]), d = item_id(items[i]);
            dh[i] = h - last_hash;
            dd[i] = (h != last_hash) ? d - min_doc_id : d - last_docid; // reset at hash change
            last_hash = h;
            last_docid = d;
        }
        size_t eh = encode_quad_0124(dh, hash_data + n_hash_data, &hash_ctrl[n_hash_ctrl]);
        size_t ed = encode_quad_1234(dd, doc_data + n_doc_data, &doc_ctrl[n_doc_ctrl]);
        size_t new_size = kBlockHeaderSize + n_hash_data + eh + n_hash_ctrl + 1 + n_doc_data + ed +
                          n_doc_ctrl + 1; // block.zig:479-481
        if (new_size > block_size) return false;
        n_hash_data += eh;
        n_hash_ctrl += 1;
        n_doc_data += ed;
        n_doc_ctrl += 1;
        num_items = (uint16_t)(num_items + n);
        return true;
    }

    // block.zig:501-567 encodeBlock.
    size_t encode_block(const uint64_t *items, size_t n, uint32_t min_doc_id, uint8_t *out,
                        size_t block_size) {
        if (n == 0) {
            std::memset(out, 0, block_size);
            return 0;
        }
        num_items = 0;
        n_hash_data = n_hash_ctrl = n_doc_data = n_doc_ctrl = 0;
        last_hash = item_hash(items[0]);
        last_docid = min_doc_id;
        const uint64_t *p = items;
        size_t left = n;
        bool full = false;
        while (left >= 4) {
            if (!encode_chunk(p, 4, min_doc_id, block_size)) {
                full = true;
                break;
            }
            p += 4;
            left -= 4;
        }
        if (!full && left > 0) (void)encode_chunk(p, left, min_doc_id, block_size); // :535-542
        BlockHeader hdr;
        hdr.min_hash = item_hash(items[0]);
        hdr.num_items = num_items;
        hdr.docids_offset = (uint16_t)(n_hash_data + n_hash_ctrl);
        uint8_t *w = out;
        std::memcpy(w, &hdr, sizeof hdr);
        w += sizeof hdr;
        std::memcpy(w, hash_ctrl, n_hash_ctrl);
        w += n_hash_ctrl;
        std::memcpy(w, hash_data, n_hash_data);
        w += n_hash_data;
        std::memcpy(w, doc_ctrl, n_doc_ctrl);
        w += n_doc_ctrl;
        std::memcpy(w, doc_data, n_doc_data);
        w += n_doc_data;
        std::memset(w, 0, block_size - (size_t)(w - out));
        return num_items;
    }
};

// block.zig:66-312 BlockReader (lazy decode of the hash column, range decode of docids).
struct BlockReader {
    uint32_t min_doc_id;
    const uint8_t *block = nullptr;
    bool hashes_loaded = false;
    uint32_t hashes[kMaxItemsPerBlock + 4];
    uint32_t docids[kMaxItemsPerBlock + 4];

    explicit BlockReader(uint32_t m) : min_doc_id(m) {}
    BlockReader() : min_doc_id(0) {}

    void load(const uint8_t *data) { // block.zig:104-117 (lazy=true)
        block = data;
        hashes_loaded = false;
    }
    BlockHeader header() const { return read_header(block); }
    uint32_t min_hash() const { return header().min_hash; } // block.zig:206-208

    void ensure_hashes_loaded() { // block.zig:137-158
        if (hashes_loaded) return;
        BlockHeader h = header();
        if (h.num_items != 0)
            decode_values(h.num_items, 0, h.num_items, block + kBlockHeaderSize, hashes, V0124, true,
                          h.min_hash);
        hashes_loaded = true;
    }
    void find_hash(uint32_t hash, size_t *start, size_t *end) { // block.zig:217-231
        BlockHeader h = header();
        if (h.num_items == 0) {
            *start = *end = 0;
            return;
        }
        ensure_hashes_loaded();
        auto r = std::equal_range(hashes, hashes + h.num_items, hash);
        *start = (size_t)(r.first - hashes);
        *end = (size_t)(r.second - hashes);
    }
    const uint32_t *docids_for_range(size_t start, size_t end) { // block.zig:235-265
        if (start >= end) return docids;
        BlockHeader h = header();
        decode_values(h.num_items, start, end, block + kBlockHeaderSize + h.docids_offset, docids,
                      V1234, false, 0);
        delta_decode_in_place(docids + start, end - start, min_doc_id);
        return docids + start;
    }
    // block.zig:160-203 full decode with base reset at hash boundaries (merge reader path).
    size_t decode_all(uint32_t *out_hashes, uint32_t *out_docids) {
        BlockHeader h = header();
        if (h.num_items == 0) return 0;
        ensure_hashes_loaded();
        decode_values(h.num_items, 0, h.num_items, block + kBlockHeaderSize + h.docids_offset, docids,
                      V1234, false, 0);
        uint32_t last_docid = min_doc_id, last_hash = hashes[0];
        for (size_t i = 0; i < h.num_items; ++i) {
            if (hashes[i] != last_hash) {
                last_docid = min_doc_id;
                last_hash = hashes[i];
            }
            docids[i] += last_docid;
            last_docid = docids[i];
            out_hashes[i] = hashes[i];
            out_docids[i] = docids[i];
        }
        return h.num_items;
    }
};

// ---------------------------------------------------------------------------------------
// Small open-addressing maps (stand-ins for Zig's std.HashMapUnmanaged, 80% max load).
// ---------------------------------------------------------------------------------------
inline uint64_t hit_hash(uint32_t key) { // common.zig:61-71 HitContext.hash
    uint64_t x = key;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

struct DocMap { // docs: id -> alive? (FileSegment.zig:39, MemorySegment.zig:25)
    std::vector<uint32_t> keys;
    std::vector<uint8_t> state; // 0 empty, 1 tombstone(false), 2 alive(true)
    size_t count = 0, mask = 0;
    void reserve(size_t n) {
        size_t cap = 8;
        while (cap * 4 < n * 5 + 5) cap <<= 1;
        keys.assign(cap, 0);
        state.assign(cap, 0);
        mask = cap - 1;
        count = 0;
    }
    // returns true if newly inserted
    bool put_if_absent(uint32_t id, bool alive) {
        size_t i = (size_t)hit_hash(id) & mask;
        while (state[i]) {
            if (keys[i] == id) return false;
            i = (i + 1) & mask;
        }
        keys[i] = id;
        state[i] = alive ? 2 : 1;
        ++count;
        return true;
    }
    bool contains(uint32_t id) const {
        if (count == 0) return false;
        size_t i = (size_t)hit_hash(id) & mask;
        while (state[i]) {
            if (keys[i] == id) return true;
            i = (i + 1) & mask;
        }
        return false;
    }
};

struct Segments;

// common.zig:73-176 SearchResults.
struct SearchResults {
    struct Entry {
        uint64_t commit_id;
        uint32_t key;
        uint32_t score;
    };
    std::vector<Entry> slots;
    std::vector<uint8_t> used;
    size_t count = 0, mask = 0;
    uint32_t max_results = 10, min_score = 1, min_score_pct = 10; // common.zig:50-54
    std::vector<std::pair<uint32_t, uint32_t>> results;           // (id, score)

    SearchResults() { rebuild(64); }
    void rebuild(size_t cap) {
        slots.assign(cap, Entry{0, 0, 0});
        used.assign(cap, 0);
        mask = cap - 1;
        count = 0;
    }
    void reset(uint32_t mr, uint32_t ms, uint32_t pct) { // pool acquire/release: keep capacity
        max_results = mr;
        min_score = ms;
        min_score_pct = pct;
        if (slots.size() > 64 * 1024) rebuild(64); // common.zig:108-119, 201
        else if (count) {
            std::fill(used.begin(), used.end(), 0);
            count = 0;
        }
        results.clear();
    }
    void grow() {
        std::vector<Entry> old;
        std::vector<uint8_t> old_used;
        old.swap(slots);
        old_used.swap(used);
        rebuild(old.size() * 2);
        for (size_t i = 0; i < old.size(); ++i)
            if (old_used[i]) {
                size_t j = (size_t)hit_hash(old[i].key) & mask;
                while (used[j]) j = (j + 1) & mask;
                slots[j] = old[i];
                used[j] = 1;
                ++count;
            }
    }
    // common.zig:121-129 incr.
    inline void incr(uint32_t id, uint64_t commit_id) {
        if ((count + 1) * 5 > slots.size() * 4) grow();
        size_t i = (size_t)hit_hash(id) & mask;
        while (used[i]) {
            if (slots[i].key == id) {
                Entry &e = slots[i];
                if (e.commit_id < commit_id) {
                    e.score = 1;
                    e.commit_id = commit_id;
                } else if (e.commit_id == commit_id) {
                    e.score += 1;
                }
                return;
            }
            i = (i + 1) & mask;
        }
        used[i] = 1;
        slots[i] = Entry{commit_id, id, 1};
        ++count;
    }
    const Entry *get(uint32_t id) const {
        size_t i = (size_t)hit_hash(id) & mask;
        while (used[i]) {
            if (slots[i].key == id) return &slots[i];
            i = (i + 1) & mask;
        }
        return nullptr;
    }
    void finish(const Segments &segs); // common.zig:131-167
};

// ---------------------------------------------------------------------------------------
// Segments
// ---------------------------------------------------------------------------------------
struct SegmentBase {
    uint64_t commit_id = 0, merges = 0; // segment.zig:23-26 SegmentInfo
    DocMap docs;
    std::vector<uint32_t> doc_ids; // ascending copy of the docs map, for views / merging
    std::vector<uint8_t> doc_alive;
    uint32_t min_doc_id = 0, max_doc_id = 0;

    void set_docs(const uint32_t *ids, const uint8_t *alive, size_t n) {
        std::vector<std::pair<uint32_t, uint8_t>> tmp(n);
        for (size_t i = 0; i < n; ++i) tmp[i] = {ids[i], (uint8_t)(alive[i] ? 1 : 0)};
        std::sort(tmp.begin(), tmp.end());
        doc_ids.resize(n);
        doc_alive.resize(n);
        docs.reserve(n);
        min_doc_id = max_doc_id = 0;
        for (size_t i = 0; i < n; ++i) {
            doc_ids[i] = tmp[i].first;
            doc_alive[i] = tmp[i].second;
            docs.put_if_absent(tmp[i].first, tmp[i].second != 0);
            // filefmt.zig:244-250 / MemorySegment.zig:115-120: 0 is the "unset" sentinel
            if (min_doc_id == 0 || tmp[i].first < min_doc_id) min_doc_id = tmp[i].first;
            if (max_doc_id == 0 || tmp[i].first > max_doc_id) max_doc_id = tmp[i].first;
        }
    }
};

// FileSegment.zig
struct FileSegment : SegmentBase {
    uint32_t block_size = 512;
    size_t num_blocks = 0, num_items = 0;
    std::vector<uint8_t> owned_blocks; // num_blocks*block_size + terminator block
    std::vector<uint32_t> owned_index;
    const uint8_t *blocks = nullptr;
    const uint32_t *block_index = nullptr;

    static constexpr size_t kMaxBlocksPerHash = 4;  // FileSegment.zig:25
    static constexpr size_t kMaxDocsPerHash = 1000; // FileSegment.zig:26

    // FileSegment.zig:135-180 search.
    void search(const uint32_t *sorted_hashes, size_t n, SearchResults &results) const {
        struct CacheEntry {
            size_t block_no;
            BlockReader reader;
        };
        // 4-entry direct-mapped block cache (FileSegment.zig:138-141); arrays left uninitialised.
        std::unique_ptr<CacheEntry[]> cache(new CacheEntry[kMaxBlocksPerHash]);
        for (size_t i = 0; i < kMaxBlocksPerHash; ++i) {
            cache[i].block_no = (size_t)-1;
            cache[i].reader.min_doc_id = min_doc_id;
        }
        size_t prev_start = 0;
        for (size_t qi = 0; qi < n; ++qi) {
            const uint32_t hash = sorted_hashes[qi];
            // :145-151 lowerBound over the max-hash block index, resuming from the previous hit
            size_t block_no = (size_t)(std::lower_bound(block_index + prev_start, block_index + num_blocks, hash) -
                                       block_index);
            prev_start = block_no;
            size_t num_docs = 0, nb = 0;
            for (; block_no < num_blocks; ++block_no) {
                CacheEntry &ce = cache[block_no % kMaxBlocksPerHash];
                if (ce.block_no != block_no) {
                    ce.block_no = block_no;
                    ce.reader.load(blocks + block_no * (size_t)block_size); // loadBlockData :83-89
                }
                BlockReader &br = ce.reader;
                if (br.min_hash() > hash) break; // :164
                size_t s, e;
                br.find_hash(hash, &s, &e); // searchHash block.zig:268-271
                const uint32_t *d = br.docids_for_range(s, e);
                for (size_t i = 0; i < e - s; ++i) results.incr(d[i], commit_id); // :167-169
                nb += 1;
                num_docs += e - s;
                if (nb >= kMaxBlocksPerHash) break;     // :173
                if (num_docs > kMaxDocsPerHash) break;  // :174
            }
        }
    }

    // FileSegment.Reader (FileSegment.zig:99-133): all items in (hash,id) order.
    void read_all(std::vector<uint64_t> &out) const {
        std::unique_ptr<BlockReader> br(new BlockReader(min_doc_id));
        std::vector<uint32_t> h(kMaxItemsPerBlock + 4), d(kMaxItemsPerBlock + 4);
        for (size_t b = 0; b < num_blocks; ++b) {
            br->load(blocks + b * (size_t)block_size);
            size_t n = br->decode_all(h.data(), d.data());
            for (size_t i = 0; i < n; ++i) out.push_back(make_item(h[i], d[i]));
        }
    }
};

// MemorySegment.zig
struct MemorySegment : SegmentBase {
    std::vector<uint64_t> items; // sorted by u64 (hash, id)

    // MemorySegment.zig:44-54 search: equalRange by hash over a shrinking suffix.
    void search(const uint32_t *sorted_hashes, size_t n, SearchResults &results) const {
        const uint64_t *lo = items.data();
        const uint64_t *end = items.data() + items.size();
        for (size_t qi = 0; qi < n; ++qi) {
            uint32_t hash = sorted_hashes[qi];
            const uint64_t *a = std::lower_bound(lo, end, make_item(hash, 0));
            const uint64_t *b = a;
            while (b < end && item_hash(*b) == hash) ++b; // upper edge of the equal range
            for (const uint64_t *p = a; p < b; ++p) results.incr(item_id(*p), commit_id);
            lo = b;
        }
    }
};

struct Segments { // Index.zig:36-150
    std::vector<std::shared_ptr<FileSegment>> file;     // oldest -> newest
    std::vector<std::shared_ptr<MemorySegment>> memory; // oldest -> newest, all newer than file

    // Index.zig:133-149 hasNewerCommit.
    bool has_newer_commit(uint32_t id, uint64_t commit_id) const {
        for (size_t i = memory.size(); i-- > 0;) {
            const MemorySegment &s = *memory[i];
            if (s.commit_id <= commit_id) return false;
            if (id >= s.min_doc_id && id <= s.max_doc_id && s.docs.contains(id)) return true;
        }
        for (size_t j = file.size(); j-- > 0;) {
            const FileSegment &s = *file[j];
            if (s.commit_id <= commit_id) return false;
            if (id >= s.min_doc_id && id <= s.max_doc_id && s.docs.contains(id)) return true;
        }
        return false;
    }
};

// common.zig:131-171 finish + compareResults.
void SearchResults::finish(const Segments &segs) {
    results.clear();
    results.reserve(count);
    uint32_t ms = min_score;
    for (size_t i = 0; i < slots.size(); ++i)
        if (used[i] && slots[i].score >= ms) results.emplace_back(slots[i].key, slots[i].score);
    std::sort(results.begin(), results.end(),
              [](const std::pair<uint32_t, uint32_t> &a, const std::pair<uint32_t, uint32_t> &b) {
                  return a.second > b.second || (a.second == b.second && a.first < b.first);
              });
    size_t out = 0;
    for (size_t i = 0; i < results.size(); ++i) {
        auto cand = results[i];
        if (out == max_results) break;
        const Entry *hit = get(cand.first);
        if (segs.has_newer_commit(cand.first, hit->commit_id)) continue; // superseded version
        if (cand.second < ms) break;
        // relative cutoff anchored on the best survivor; u32 arithmetic (wrapping), truncating div
        if (out == 0) ms = std::max(ms, (uint32_t)(cand.second * min_score_pct) / 100u);
        results[out++] = cand;
    }
    results.resize(out);
}

// Index.zig:165-177 IndexReader.search + :489-499 dedupSorted.
void reader_search(const Segments &segs, const uint32_t *query, size_t n, SearchResults &r,
                   std::vector<uint32_t> &scratch) {
    scratch.assign(query, query + n);
    std::sort(scratch.begin(), scratch.end());
    scratch.erase(std::unique(scratch.begin(), scratch.end()), scratch.end());
    for (auto &s : segs.file) s->search(scratch.data(), scratch.size(), r);
    for (auto &s : segs.memory) s->search(scratch.data(), scratch.size(), r);
    r.finish(segs);
}

// filefmt.zig:94-138 writeBlocks: greedy packing from a 2048-item look-ahead window.
void write_blocks(const uint64_t *items, size_t n, uint32_t min_doc_id, uint32_t block_size,
                  FileSegment &seg) {
    std::unique_ptr<BlockEncoder> enc(new BlockEncoder());
    seg.owned_blocks.clear();
    seg.owned_index.clear();
    seg.block_size = block_size;
    seg.num_items = 0;
    std::vector<uint8_t> blk(block_size);
    size_t p = 0;
    for (;;) {
        size_t window = std::min(kMaxItemsPerBlock, n - p); // items_buffer refill :106-111
        size_t consumed = enc->encode_block(items + p, window, min_doc_id, blk.data(), block_size);
        seg.owned_blocks.insert(seg.owned_blocks.end(), blk.begin(), blk.end());
        if (consumed == 0) break; // empty terminator block written above :115
        seg.owned_index.push_back(item_hash(items[p + consumed - 1])); // :117
        seg.num_items += consumed;
        p += consumed;
    }
    seg.num_blocks = seg.owned_index.size();
    seg.blocks = seg.owned_blocks.data();
    seg.block_index = seg.owned_index.data();
}

// segment_merger.zig:85-151 prepare + read, materialised.
struct Merged {
    uint64_t commit_id = 0, merges = 0;
    std::vector<uint32_t> doc_ids;
    std::vector<uint8_t> doc_alive;
    std::vector<uint64_t> items;
};
template <class Seg>
Merged merge_segments(const std::vector<std::shared_ptr<Seg>> &sources, const Segments &collection) {
    Merged m;
    for (size_t i = 0; i < sources.size(); ++i) {
        const Seg &s = *sources[i];
        if (i == 0) {
            m.commit_id = s.commit_id;
            m.merges = s.merges;
        } else { // segment.zig:38-51 SegmentInfo.merge
            m.commit_id = std::min(m.commit_id, s.commit_id);
            m.merges = m.merges + s.merges + 1;
        }
    }
    std::vector<std::pair<uint32_t, uint8_t>> docs;
    for (auto &sp : sources) {
        const Seg &s = *sp;
        DocMap skip;
        skip.reserve(s.doc_ids.size());
        for (size_t i = 0; i < s.doc_ids.size(); ++i) {
            if (!collection.has_newer_commit(s.doc_ids[i], s.commit_id)) // :119
                docs.emplace_back(s.doc_ids[i], s.doc_alive[i]);
            else
                skip.put_if_absent(s.doc_ids[i], true);
        }
        std::vector<uint64_t> src;
        if constexpr (std::is_same<Seg, FileSegment>::value) s.read_all(src);
        else src = s.items;
        for (uint64_t it : src)
            if (!skip.contains(item_id(it))) m.items.push_back(it);
    }
    std::sort(m.items.begin(), m.items.end()); // k-way merge by Item.order (:131-151)
    std::sort(docs.begin(), docs.end());
    for (auto &d : docs) {
        m.doc_ids.push_back(d.first);
        m.doc_alive.push_back(d.second);
    }
    return m;
}

} // namespace

struct orc_index {
    Segments segs;
    uint64_t commit_id = 0; // last minted commit id (dense, one per write; segment.zig:6-9)
    uint32_t block_size = 512;
};

namespace {
std::shared_ptr<FileSegment> file_segment_from(const Merged &m, uint32_t block_size) {
    auto f = std::make_shared<FileSegment>();
    f->commit_id = m.commit_id;
    f->merges = m.merges;
    f->set_docs(m.doc_ids.data(), m.doc_alive.data(), m.doc_ids.size());
    write_blocks(m.items.data(), m.items.size(), f->min_doc_id, block_size, *f);
    return f;
}
} // namespace

extern "C" {

orc_index *orc_index_new(uint32_t block_size) {
    auto *ix = new orc_index();
    ix->block_size = block_size ? block_size : 512; // filefmt.zig:29
    return ix;
}
void orc_index_free(orc_index *ix) { delete ix; }

// MemorySegment.zig:81-148 build (reverse pass: the LAST change per id in the batch wins).
int orc_update(orc_index *ix, size_t n_changes, const uint8_t *kinds, const uint32_t *ids,
               const uint64_t *hash_offsets, const uint32_t *hashes) {
    auto m = std::make_shared<MemorySegment>();
    m->commit_id = ++ix->commit_id;
    m->merges = 0;
    DocMap seen;
    seen.reserve(n_changes);
    std::vector<std::pair<uint32_t, uint8_t>> docs;
    for (size_t i = n_changes; i-- > 0;) {
        uint32_t id = ids[i];
        if (kinds[i] == ORC_INSERT) {
            if (seen.put_if_absent(id, true)) {
                docs.emplace_back(id, 1);
                for (uint64_t j = hash_offsets[i]; j < hash_offsets[i + 1]; ++j)
                    m->items.push_back(make_item(hashes[j], id));
            }
        } else if (kinds[i] == ORC_DELETE) {
            if (seen.put_if_absent(id, false)) docs.emplace_back(id, 0);
        } else {
            return -1;
        }
    }
    std::sort(m->items.begin(), m->items.end()); // MemorySegment.zig:139
    std::vector<uint32_t> dids(docs.size());
    std::vector<uint8_t> dal(docs.size());
    for (size_t i = 0; i < docs.size(); ++i) {
        dids[i] = docs[i].first;
        dal[i] = docs[i].second;
    }
    m->set_docs(dids.data(), dal.data(), dids.size());
    ix->segs.memory.push_back(m);
    return 0;
}

// Index.zig:770-862 checkpoint(force=true).
int orc_checkpoint(orc_index *ix) {
    if (ix->segs.memory.empty()) return 0;
    Merged m = merge_segments(ix->segs.memory, ix->segs);
    ix->segs.file.push_back(file_segment_from(m, ix->block_size));
    ix->segs.memory.clear();
    return 1;
}

int orc_merge_memory(orc_index *ix, size_t lo, size_t count) {
    auto &mem = ix->segs.memory;
    if (count < 2 || lo + count > mem.size()) return -1;
    std::vector<std::shared_ptr<MemorySegment>> src(mem.begin() + lo, mem.begin() + lo + count);
    Merged m = merge_segments(src, ix->segs);
    auto out = std::make_shared<MemorySegment>(); // MemorySegment.zig:63-79 buildFromMerger
    out->commit_id = m.commit_id;
    out->merges = m.merges;
    out->items = std::move(m.items);
    out->set_docs(m.doc_ids.data(), m.doc_alive.data(), m.doc_ids.size());
    mem.erase(mem.begin() + lo, mem.begin() + lo + count);
    mem.insert(mem.begin() + lo, out);
    return 0;
}

int orc_merge_files(orc_index *ix, size_t lo, size_t count) {
    auto &fl = ix->segs.file;
    if (count < 2 || lo + count > fl.size()) return -1;
    std::vector<std::shared_ptr<FileSegment>> src(fl.begin() + lo, fl.begin() + lo + count);
    Merged m = merge_segments(src, ix->segs);
    auto out = file_segment_from(m, ix->block_size);
    fl.erase(fl.begin() + lo, fl.begin() + lo + count);
    fl.insert(fl.begin() + lo, out);
    return 0;
}

int orc_add_file_segment_sorted(orc_index *ix, const uint64_t *items, size_t n_items,
                                const uint32_t *doc_ids, const uint8_t *doc_alive, size_t n_docs) {
    if (!ix->segs.memory.empty()) return -1; // file segments are older than all memory segments
    for (size_t i = 1; i < n_items; ++i)
        if (items[i] < items[i - 1]) return -2;
    auto f = std::make_shared<FileSegment>();
    f->commit_id = ++ix->commit_id;
    f->merges = 0;
    f->set_docs(doc_ids, doc_alive, n_docs);
    write_blocks(items, n_items, f->min_doc_id, ix->block_size, *f);
    ix->segs.file.push_back(f);
    return 0;
}

int orc_adopt_file_segment(orc_index *ix, uint64_t commit_id, uint64_t merges, uint32_t block_size,
                           const uint8_t *blocks, size_t n_blocks, const uint32_t *block_index,
                           const uint32_t *doc_ids, const uint8_t *doc_alive, size_t n_docs) {
    if (!ix->segs.memory.empty()) return -1;
    auto f = std::make_shared<FileSegment>();
    f->commit_id = commit_id;
    f->merges = merges;
    f->block_size = block_size;
    f->num_blocks = n_blocks;
    f->blocks = blocks;
    f->block_index = block_index;
    f->num_items = 0;
    for (size_t b = 0; b < n_blocks; ++b) f->num_items += read_header(blocks + b * (size_t)block_size).num_items;
    f->set_docs(doc_ids, doc_alive, n_docs);
    ix->segs.file.push_back(f);
    ix->commit_id = std::max(ix->commit_id, commit_id + merges);
    return 0;
}

size_t orc_num_file_segments(const orc_index *ix) { return ix->segs.file.size(); }
size_t orc_num_memory_segments(const orc_index *ix) { return ix->segs.memory.size(); }

int orc_file_segment(const orc_index *ix, size_t i, orc_file_segment_view *v) {
    if (i >= ix->segs.file.size()) return -1;
    const FileSegment &f = *ix->segs.file[i];
    v->commit_id = f.commit_id;
    v->merges = f.merges;
    v->min_doc_id = f.min_doc_id;
    v->max_doc_id = f.max_doc_id;
    v->block_size = f.block_size;
    v->_pad = 0;
    v->num_blocks = f.num_blocks;
    v->num_items = f.num_items;
    v->blocks = f.blocks;
    v->block_index = f.block_index;
    v->doc_ids = f.doc_ids.data();
    v->doc_alive = f.doc_alive.data();
    v->n_docs = f.doc_ids.size();
    return 0;
}

int orc_memory_segment(const orc_index *ix, size_t i, orc_memory_segment_view *v) {
    if (i >= ix->segs.memory.size()) return -1;
    const MemorySegment &m = *ix->segs.memory[i];
    v->commit_id = m.commit_id;
    v->merges = m.merges;
    v->min_doc_id = m.min_doc_id;
    v->max_doc_id = m.max_doc_id;
    v->items = m.items.data();
    v->n_items = m.items.size();
    v->doc_ids = m.doc_ids.data();
    v->doc_alive = m.doc_alive.data();
    v->n_docs = m.doc_ids.size();
    return 0;
}

int64_t orc_search(const orc_index *ix, const uint32_t *query, size_t n_terms, uint32_t max_results,
                   uint32_t min_score, uint32_t min_score_pct, uint32_t *out_ids,
                   uint32_t *out_scores, size_t cap) {
    SearchResults r;
    std::vector<uint32_t> scratch;
    r.reset(max_results, min_score, min_score_pct);
    reader_search(ix->segs, query, n_terms, r, scratch);
    size_t n = std::min(cap, r.results.size());
    for (size_t i = 0; i < n; ++i) {
        out_ids[i] = r.results[i].first;
        out_scores[i] = r.results[i].second;
    }
    return (int64_t)r.results.size();
}

double orc_search_batch(const orc_index *ix, size_t n_queries, const uint32_t *terms,
                        const uint64_t *term_offsets, const uint32_t *opts3, uint32_t k_stride,
                        uint32_t *out_ids, uint32_t *out_scores, uint32_t *out_counts,
                        unsigned n_threads) {
    if (n_threads == 0) n_threads = 1;
    std::atomic<size_t> next{0};
    const size_t grain = 16;
    auto worker = [&]() {
        SearchResults r; // one pooled collector per worker (common.zig:186-300)
        std::vector<uint32_t> scratch;
        for (;;) {
            size_t q0 = next.fetch_add(grain);
            if (q0 >= n_queries) break;
            size_t q1 = std::min(n_queries, q0 + grain);
            for (size_t q = q0; q < q1; ++q) {
                r.reset(opts3[3 * q], opts3[3 * q + 1], opts3[3 * q + 2]);
                reader_search(ix->segs, terms + term_offsets[q], (size_t)(term_offsets[q + 1] - term_offsets[q]), r,
                              scratch);
                size_t n = std::min<size_t>(k_stride, r.results.size());
                for (size_t i = 0; i < n; ++i) {
                    out_ids[q * k_stride + i] = r.results[i].first;
                    out_scores[q * k_stride + i] = r.results[i].second;
                }
                out_counts[q] = (uint32_t)n;
            }
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (unsigned t = 1; t < n_threads; ++t) th.emplace_back(worker);
    worker();
    for (auto &t : th) t.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

size_t orc_svb_encode_quad_0124(const uint32_t in[4], uint8_t *out_data, uint8_t *out_control) {
    return encode_quad_0124(in, out_data, out_control);
}
size_t orc_svb_encode_quad_1234(const uint32_t in[4], uint8_t *out_data, uint8_t *out_control) {
    return encode_quad_1234(in, out_data, out_control);
}
size_t orc_svb_decode_quad(int variant, uint8_t control, const uint8_t *in, uint32_t out[4]) {
    return decode_quad(variant, control, in, out);
}
size_t orc_svb_decode_quad_delta(int variant, uint8_t control, const uint8_t *in, uint32_t out[4],
                                 uint32_t carry) {
    return decode_quad_delta(variant, control, in, out, carry);
}
void orc_svb_delta_decode_in_place(uint32_t *data, size_t n, uint32_t first_value) {
    delta_decode_in_place(data, n, first_value);
}
void orc_svb_decode_values(size_t total_items, size_t start_item, size_t end_item, const uint8_t *in,
                           uint32_t *out, int variant, int delta, uint32_t first_value) {
    decode_values(total_items, start_item, end_item, in, out, variant, delta != 0, first_value);
}
size_t orc_encode_block(const uint64_t *items, size_t n_items, uint32_t min_doc_id, uint8_t *out,
                        size_t block_size) {
    std::unique_ptr<BlockEncoder> enc(new BlockEncoder());
    return enc->encode_block(items, n_items, min_doc_id, out, block_size);
}
size_t orc_decode_block(const uint8_t *block, size_t block_size, uint32_t min_doc_id,
                        uint32_t *out_hashes, uint32_t *out_docids) {
    (void)block_size;
    std::unique_ptr<BlockReader> br(new BlockReader(min_doc_id));
    br->load(block);
    return br->decode_all(out_hashes, out_docids);
}
size_t orc_block_search_hash(const uint8_t *block, size_t block_size, uint32_t min_doc_id,
                             uint32_t hash, uint32_t *out_start, uint32_t *out_end,
                             uint32_t *out_docids) {
    (void)block_size;
    std::unique_ptr<BlockReader> br(new BlockReader(min_doc_id));
    br->load(block);
    size_t s, e;
    br->find_hash(hash, &s, &e);
    const uint32_t *d = br->docids_for_range(s, e);
    *out_start = (uint32_t)s;
    *out_end = (uint32_t)e;
    for (size_t i = 0; i < e - s; ++i) out_docids[i] = d[i];
    return e - s;
}
int orc_uses_ssse3(void) { return ORC_SSSE3; }

} // extern "C"
