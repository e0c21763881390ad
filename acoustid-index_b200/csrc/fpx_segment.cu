// fpx_segment.cu — C entry points of the segment writer (include/fpx_segment.h).
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fpx_segment.h"
#include "fpx_codec.h"
#include "fpx_snapshot_host.h"

using namespace fpx;

struct fpx_segment_buf {
    std::vector<uint8_t> blocks; // num_blocks * block_size + terminator
    std::vector<uint32_t> index;
    uint64_t num_items = 0;
    uint32_t block_size = 512;
};

namespace {

struct PackedRange {
    std::vector<uint8_t> bytes;
    std::vector<uint32_t> index;
    std::vector<uint64_t> starts; // item index at which each block starts
};

// Packs blocks starting at item `begin` until a block would start at or after `stop`
// (a block may extend past `stop`).  Returns the item index after the last packed block.
uint64_t pack_range(const uint64_t *items, uint64_t n, uint64_t begin, uint64_t stop, uint32_t min_doc_id,
                    uint32_t block_size, PackedRange &out) {
    BlockPacker packer;
    std::vector<uint8_t> blk(block_size);
    uint64_t p = begin;
    while (p < stop) {
        const size_t window = (size_t)std::min<uint64_t>(kWriterWindow, n - p); // filefmt.zig:106-111
        const size_t took = packer.pack(items + p, window, min_doc_id, blk.data(), block_size);
        if (took == 0) break;
        out.starts.push_back(p);
        out.bytes.insert(out.bytes.end(), blk.begin(), blk.end());
        out.index.push_back((uint32_t)(items[p + took - 1] >> 32)); // filefmt.zig:117
        p += took;
    }
    return p;
}

} // namespace

extern "C" {

fpx_status fpx_segment_write(const uint64_t *items, uint64_t n_items, uint32_t min_doc_id, uint32_t block_size,
                             uint32_t threads, fpx_segment_buf **out) {
    if (!out || (n_items && !items)) return FPX_INVALID_ARGUMENT;
    if (block_size == 0) block_size = 512;
    if (block_size < kMinBlockSize || block_size > kMaxBlockSize) return FPX_INVALID_ARGUMENT;
    if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
    fpx_segment_buf *b = new (std::nothrow) fpx_segment_buf();
    if (!b) return FPX_OUT_OF_MEMORY;
    b->block_size = block_size;
    try {
        // Greedy packing is sequential by definition, but two greedy walks that ever share a block
        // boundary coincide from there on.  So: each thread packs its own stretch speculatively from a
        // guessed boundary; the stitcher then replays from the true boundary only until it lands on a
        // boundary the next stretch already produced, and adopts the rest unchanged.
        const uint64_t min_stretch = 1u << 16;
        unsigned parts = (unsigned)std::min<uint64_t>(threads, std::max<uint64_t>(1, n_items / min_stretch));
        std::vector<PackedRange> spec(parts);
        std::vector<uint64_t> cut(parts + 1);
        for (unsigned t = 0; t <= parts; ++t) cut[t] = (n_items * t / parts) & ~3ull; // quads never straddle a guess
        cut[parts] = n_items;
        std::vector<std::thread> th;
        for (unsigned t = 0; t < parts; ++t)
            th.emplace_back([&, t] { pack_range(items, n_items, cut[t], cut[t + 1], min_doc_id, block_size, spec[t]); });
        for (auto &x : th) x.join();

        uint64_t pos = 0; // true boundary reached so far
        for (unsigned t = 0; t < parts; ++t) {
            const PackedRange &s = spec[t];
            // replay from `pos` until it coincides with one of this stretch's block starts
            size_t k = 0;
            PackedRange fix;
            while (pos < cut[t + 1]) {
                while (k < s.starts.size() && s.starts[k] < pos) ++k;
                if (k < s.starts.size() && s.starts[k] == pos) break;
                const uint64_t next = pack_range(items, n_items, pos, pos + 1, min_doc_id, block_size, fix);
                if (next == pos) break;
                pos = next;
            }
            b->blocks.insert(b->blocks.end(), fix.bytes.begin(), fix.bytes.end());
            b->index.insert(b->index.end(), fix.index.begin(), fix.index.end());
            if (pos < cut[t + 1] && k < s.starts.size() && s.starts[k] == pos) {
                b->blocks.insert(b->blocks.end(), s.bytes.begin() + (size_t)k * block_size, s.bytes.end());
                b->index.insert(b->index.end(), s.index.begin() + k, s.index.end());
                // end of the adopted stretch = start of its last block + that block's item count
                const uint8_t *last = s.bytes.data() + (s.starts.size() - 1) * (size_t)block_size;
                pos = s.starts.back() + read_block_head(last).num_items;
            }
        }
        b->num_items = pos;
        b->blocks.resize(b->blocks.size() + block_size, 0); // terminator block (filefmt.zig:113-115)
    } catch (const std::bad_alloc &) {
        delete b;
        return FPX_OUT_OF_MEMORY;
    }
    *out = b;
    return FPX_OK;
}

const uint8_t *fpx_segment_buf_blocks(const fpx_segment_buf *b) { return b->blocks.data(); }
const uint32_t *fpx_segment_buf_block_index(const fpx_segment_buf *b) { return b->index.data(); }
uint64_t fpx_segment_buf_num_blocks(const fpx_segment_buf *b) { return b->index.size(); }
uint64_t fpx_segment_buf_num_items(const fpx_segment_buf *b) { return b->num_items; }
uint32_t fpx_segment_buf_block_size(const fpx_segment_buf *b) { return b->block_size; }
void fpx_segment_buf_free(fpx_segment_buf *b) { delete b; }

int32_t fpx_block_decode(const uint8_t *block, uint32_t block_size, uint32_t min_doc_id, uint32_t *out_hashes,
                         uint32_t *out_docids) {
    return decode_block(block, block_size, min_doc_id, out_hashes, out_docids);
}

} // extern "C"
