// fpx_codec.h — host-side reader/writer for fpindex's 512-byte segment blocks.
//
// Wire format (reference: src/block.zig:30-50, src/streamvbyte.zig:76-211, 418-480):
//   header  { u32 min_hash; u16 num_items; u16 docids_offset; }           (little endian)
//   hashes  ceil(n/4) control bytes, then data: per value 2 control bits -> {0,1,2,4} bytes,
//           values are deltas from the previous hash, first base = min_hash
//   docids  at 8 + docids_offset: ceil(n/4) control bytes, then data: 2 bits -> {1,2,3,4} bytes,
//           values are deltas from the previous docid of the SAME hash; the base resets to the
//           segment's min_doc_id at every hash change and at block start.
//
// This is the product's own codec (a byte-cursor design, no shuffle tables); the oracle under
// oracle/ has an independent restatement and tests/ cross-check the two byte for byte.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace fpx {

constexpr uint32_t kMinBlockSize = 64;    // block.zig:41
constexpr uint32_t kMaxBlockSize = 4096;  // block.zig:42
constexpr uint32_t kBlockHeaderBytes = 8; // block.zig:44
constexpr uint32_t kWriterWindow = kMaxBlockSize / 2; // filefmt.zig:96 look-ahead (MAX_ITEMS_PER_BLOCK)

struct BlockHead {
    uint32_t min_hash;
    uint32_t num_items;
    uint32_t docids_offset;
};

inline BlockHead read_block_head(const uint8_t *p) {
    BlockHead h;
    h.min_hash = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    h.num_items = (uint32_t)p[4] | ((uint32_t)p[5] << 8);
    h.docids_offset = (uint32_t)p[6] | ((uint32_t)p[7] << 8);
    return h;
}

inline uint32_t load_le(const uint8_t *p, uint32_t nbytes) {
    uint32_t v = 0;
    for (uint32_t i = 0; i < nbytes; ++i) v |= (uint32_t)p[i] << (8 * i);
    return v;
}

// Decodes one block into hashes[0..n) / docids[0..n).  Returns n, or -1 if the block's streams
// run past `block_size` (corrupt block).  Never reads outside [block, block+block_size).
inline int decode_block(const uint8_t *block, uint32_t block_size, uint32_t min_doc_id,
                        uint32_t *hashes, uint32_t *docids) {
    const BlockHead hd = read_block_head(block);
    const uint32_t n = hd.num_items;
    if (n == 0) return 0;
    const uint32_t quads = (n + 3) / 4;
    if (n > kWriterWindow) return -1;
    // hash column
    uint32_t cpos = kBlockHeaderBytes, dpos = kBlockHeaderBytes + quads;
    if (dpos > block_size) return -1;
    uint32_t acc = hd.min_hash;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t code = (block[cpos + (i >> 2)] >> (2 * (i & 3))) & 3u;
        const uint32_t len = code == 3 ? 4 : code;
        if (dpos + len > block_size) return -1;
        acc += load_le(block + dpos, len);
        dpos += len;
        hashes[i] = acc;
    }
    // docid column
    cpos = kBlockHeaderBytes + hd.docids_offset;
    dpos = cpos + quads;
    if (dpos > block_size) return -1;
    uint32_t prev_hash = hashes[0], base = min_doc_id;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t len = ((block[cpos + (i >> 2)] >> (2 * (i & 3))) & 3u) + 1;
        if (dpos + len > block_size) return -1;
        if (hashes[i] != prev_hash) {
            prev_hash = hashes[i];
            base = min_doc_id;
        }
        base += load_le(block + dpos, len);
        dpos += len;
        docids[i] = base;
    }
    return (int)n;
}

// Greedy block packer (reference behaviour: block.zig:438-567 + filefmt.zig:94-138).
// A quad of items is accepted iff header + both streams (incl. the new control bytes) still fit.
class BlockPacker {
  public:
    // Packs a prefix of items[0..n) (each (hash<<32)|id, ascending) into out[0..block_size).
    // Returns how many items went in.  n == 0 writes the all-zero terminator block.
    size_t pack(const uint64_t *items, size_t n, uint32_t min_doc_id, uint8_t *out, uint32_t block_size) {
        std::memset(out, 0, block_size);
        if (n == 0) return 0;
        hn_ = dn_ = quads_ = 0;
        uint32_t prev_hash = (uint32_t)(items[0] >> 32), prev_doc = min_doc_id;
        size_t taken = 0;
        while (taken < n) {
            const size_t m = n - taken < 4 ? n - taken : 4;
            // stage one quad (missing lanes encode the value 0: 0 hash bytes, 1 docid byte)
            uint8_t hbuf[16], dbuf[16];
            uint32_t hb = 0, db = 0;
            uint8_t hc = 0, dc = 0;
            uint32_t ph = prev_hash, pd = prev_doc;
            for (size_t i = 0; i < 4; ++i) {
                uint32_t hv = 0, dv = 0;
                if (i < m) {
                    const uint32_t h = (uint32_t)(items[taken + i] >> 32), d = (uint32_t)items[taken + i];
                    hv = h - ph;
                    dv = (h != ph) ? d - min_doc_id : d - pd;
                    ph = h;
                    pd = d;
                }
                const uint32_t hl = hv == 0 ? 0 : hv < 0x100u ? 1 : hv < 0x10000u ? 2 : 4;
                const uint32_t dl = dv < 0x100u ? 1 : dv < 0x10000u ? 2 : dv < 0x1000000u ? 3 : 4;
                hc |= (uint8_t)((hl == 4 ? 3 : hl) << (2 * i));
                dc |= (uint8_t)((dl - 1) << (2 * i));
                for (uint32_t k = 0; k < hl; ++k) hbuf[hb++] = (uint8_t)(hv >> (8 * k));
                for (uint32_t k = 0; k < dl; ++k) dbuf[db++] = (uint8_t)(dv >> (8 * k));
            }
            const uint32_t need = kBlockHeaderBytes + (quads_ + 1) + hn_ + hb + (quads_ + 1) + dn_ + db;
            if (need > block_size) break; // block full; a short tail quad is only ever tried last
            hctrl_[quads_] = hc;
            dctrl_[quads_] = dc;
            std::memcpy(hdata_ + hn_, hbuf, hb);
            std::memcpy(ddata_ + dn_, dbuf, db);
            hn_ += hb;
            dn_ += db;
            quads_ += 1;
            prev_hash = ph;
            prev_doc = pd;
            taken += m;
            if (m < 4) break;
        }
        const uint32_t doff = quads_ + hn_;
        const uint32_t mh = (uint32_t)(items[0] >> 32);
        out[0] = (uint8_t)mh;
        out[1] = (uint8_t)(mh >> 8);
        out[2] = (uint8_t)(mh >> 16);
        out[3] = (uint8_t)(mh >> 24);
        out[4] = (uint8_t)taken;
        out[5] = (uint8_t)(taken >> 8);
        out[6] = (uint8_t)doff;
        out[7] = (uint8_t)(doff >> 8);
        uint8_t *w = out + kBlockHeaderBytes;
        std::memcpy(w, hctrl_, quads_);
        w += quads_;
        std::memcpy(w, hdata_, hn_);
        w += hn_;
        std::memcpy(w, dctrl_, quads_);
        w += quads_;
        std::memcpy(w, ddata_, dn_);
        return taken;
    }

  private:
    uint8_t hctrl_[kMaxBlockSize], dctrl_[kMaxBlockSize];
    uint8_t hdata_[kMaxBlockSize + 16], ddata_[kMaxBlockSize + 16];
    uint32_t hn_ = 0, dn_ = 0, quads_ = 0;
};

} // namespace fpx
