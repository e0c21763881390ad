// fpx_segment.cu — C entry points of the segment writer (include/fpx_segment.h).
#include <cstdio>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fpx_segment.h"
#include "fpx_codec.h"
#include "fpx_filefmt.h"
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

// ------------------------------------------------------------------------------------------------
// segment files and the manifest (fpx_filefmt.h)
// ------------------------------------------------------------------------------------------------
namespace fpx {
void set_last_error(const std::string &msg); // fpx_api.cu
}

struct fpx_segment_file {
    fpx::SegmentFile f;
};

extern "C" {

fpx_status fpx_segment_file_parse(const uint8_t *data, uint64_t size, fpx_segment_file **out) {
    if (!out || (size && !data)) return FPX_INVALID_ARGUMENT;
    *out = nullptr;
    fpx_segment_file *h = new (std::nothrow) fpx_segment_file();
    if (!h) return FPX_OUT_OF_MEMORY;
    try {
        h->f.owned.assign(data, data + size);
        if (!h->f.parse(h->f.owned.data(), h->f.owned.size())) {
            fpx::set_last_error(h->f.error);
            delete h;
            return FPX_INVALID_SEGMENT;
        }
    } catch (const std::bad_alloc &) {
        delete h;
        return FPX_OUT_OF_MEMORY;
    }
    *out = h;
    return FPX_OK;
}

fpx_status fpx_segment_file_read(const char *path, fpx_segment_file **out) {
    if (!out || !path) return FPX_INVALID_ARGUMENT;
    *out = nullptr;
    std::FILE *fp = std::fopen(path, "rb");
    if (!fp) {
        fpx::set_last_error(std::string("cannot open ") + path);
        return FPX_INVALID_SEGMENT;
    }
    fpx_segment_file *h = new (std::nothrow) fpx_segment_file();
    if (!h) {
        std::fclose(fp);
        return FPX_OUT_OF_MEMORY;
    }
    fpx_status rc = FPX_OK;
    try {
        std::fseek(fp, 0, SEEK_END);
        const long sz = std::ftell(fp);
        std::fseek(fp, 0, SEEK_SET);
        h->f.owned.resize(sz > 0 ? (size_t)sz : 0);
        if (sz > 0 && std::fread(h->f.owned.data(), 1, (size_t)sz, fp) != (size_t)sz) {
            fpx::set_last_error("short read"); // error.UnexpectedEndOfFile, filefmt.zig:226
            rc = FPX_INVALID_SEGMENT;
        } else if (!h->f.parse(h->f.owned.data(), h->f.owned.size())) {
            fpx::set_last_error(h->f.error);
            rc = FPX_INVALID_SEGMENT;
        }
    } catch (const std::bad_alloc &) {
        rc = FPX_OUT_OF_MEMORY;
    }
    std::fclose(fp);
    if (rc != FPX_OK) {
        delete h;
        return rc;
    }
    *out = h;
    return FPX_OK;
}

fpx_status fpx_segment_file_view(const fpx_segment_file *h, fpx_file_segment *out, fpx_segment_info *info) {
    if (!h || !out) return FPX_INVALID_ARGUMENT;
    const fpx::SegmentFile &f = h->f;
    out->commit_id = f.info.commit_id;
    out->merges = f.info.merges;
    out->min_doc_id = f.min_doc_id;
    out->block_size = f.block_size;
    out->blocks = f.blocks;
    out->num_blocks = f.num_blocks;
    out->block_index = f.block_index.data();
    out->doc_ids = f.doc_ids.data();
    out->doc_alive = f.doc_alive.data();
    out->n_docs = f.doc_ids.size();
    if (info) {
        info->commit_id = f.info.commit_id;
        info->merges = f.info.merges;
        info->version = f.info.version;
        info->has_version = f.info.has_version ? 1u : 0u;
        info->reserved = 0;
    }
    return FPX_OK;
}

uint64_t fpx_segment_file_num_items(const fpx_segment_file *h) { return h ? h->f.num_items : 0; }
uint64_t fpx_segment_file_metadata_count(const fpx_segment_file *h) { return h ? h->f.metadata.size() : 0; }

fpx_status fpx_segment_file_metadata_get(const fpx_segment_file *h, uint64_t i, const char **key, uint64_t *key_len,
                                         const char **value, uint64_t *value_len) {
    if (!h || i >= h->f.metadata.size() || !key || !key_len || !value || !value_len) return FPX_INVALID_ARGUMENT;
    *key = h->f.metadata[i].first.data();
    *key_len = h->f.metadata[i].first.size();
    *value = h->f.metadata[i].second.data();
    *value_len = h->f.metadata[i].second.size();
    return FPX_OK;
}

void fpx_segment_file_close(fpx_segment_file *h) { delete h; }

fpx_status fpx_segment_file_serialize(const fpx_file_segment *seg, const fpx_segment_info *info, uint8_t **out,
                                      uint64_t *out_size) {
    if (!seg || !out || !out_size) return FPX_INVALID_ARGUMENT;
    if (seg->block_size < kMinBlockSize || seg->block_size > kMaxBlockSize) return FPX_INVALID_ARGUMENT;
    if ((seg->num_blocks && (!seg->blocks || !seg->block_index)) || (seg->n_docs && (!seg->doc_ids || !seg->doc_alive)))
        return FPX_INVALID_ARGUMENT;
    fpx::SegmentInfoHost si;
    si.commit_id = info ? info->commit_id : seg->commit_id;
    si.merges = info ? info->merges : seg->merges;
    si.has_version = info && info->has_version;
    si.version = info ? info->version : 0;
    try {
        std::vector<uint8_t> bytes;
        fpx::serialize_segment_file(si, seg->block_size, seg->blocks, seg->num_blocks, seg->block_index, seg->doc_ids,
                                    seg->doc_alive, seg->n_docs, {}, bytes);
        uint8_t *p = static_cast<uint8_t *>(std::malloc(bytes.size() ? bytes.size() : 1));
        if (!p) return FPX_OUT_OF_MEMORY;
        std::memcpy(p, bytes.data(), bytes.size());
        *out = p;
        *out_size = bytes.size();
    } catch (const std::bad_alloc &) {
        return FPX_OUT_OF_MEMORY;
    }
    return FPX_OK;
}

void fpx_bytes_free(uint8_t *p) { std::free(p); }

int32_t fpx_segment_file_name(uint64_t commit_id, uint64_t merges, char *buf, uint64_t cap) {
    const int n = std::snprintf(buf, (size_t)cap, "%016llx-%08llx.data", (unsigned long long)commit_id,
                                (unsigned long long)merges);
    return (n < 0 || (uint64_t)n >= cap) ? -1 : n;
}

fpx_status fpx_manifest_parse(const uint8_t *data, uint64_t size, fpx_segment_info *out, uint64_t cap, uint64_t *n) {
    if (!n || (size && !data)) return FPX_INVALID_ARGUMENT;
    std::vector<fpx::SegmentInfoHost> v;
    try {
        if (!fpx::parse_manifest(data, (size_t)size, v)) {
            fpx::set_last_error("malformed manifest");
            return FPX_INVALID_SEGMENT;
        }
    } catch (const std::bad_alloc &) {
        return FPX_OUT_OF_MEMORY;
    }
    *n = v.size();
    if (v.size() > cap || (v.size() && !out)) return FPX_INVALID_ARGUMENT;
    for (size_t i = 0; i < v.size(); ++i) {
        out[i].commit_id = v[i].commit_id;
        out[i].merges = v[i].merges;
        out[i].version = v[i].version;
        out[i].has_version = v[i].has_version ? 1u : 0u;
        out[i].reserved = 0;
    }
    return FPX_OK;
}

uint64_t fpx_crc64_xz(const uint8_t *data, uint64_t size) { return fpx::crc64xz().of(data, (size_t)size); }

} // extern "C"
