// fpx_snapshot_host.h — snapshot compiler: reference segments -> flat CSR of live, reachable postings.
//
// The reference re-decodes compressed blocks on every query (src/FileSegment.zig:135-180).  Which
// postings a query term can reach in a segment does not depend on the query, so we decide it ONCE
// here, at the point where the reference installs a new snapshot (src/Index.zig:469-485):
//
//   reachable(S,h): postings the scan loop would visit for hash h in file segment S — the blocks
//       starting at lowerBound(block_index, h), while block.min_hash <= h, stopping after 4 blocks or
//       once more than 1000 docs were seen (FileSegment.zig:25-26, 145-175).  Memory segments are
//       never truncated (MemorySegment.zig:44-54).
//   live(S,p): no segment newer than S mentions p.id in its docs map — insert OR tombstone
//       (Index.zig:133-149 hasNewerCommit, combined with common.zig:121-129 incr: the newest
//       segment with a matching posting owns the score, and finish() drops it if any newer
//       segment mentions the id).
//   CSR[h] = multiset union over S of { p.id : p in reachable(S,h), live(S,p) }, sorted by id.
//
// With that, score(id) for a query with unique term set Q is sum_{h in Q} mult(CSR[h], id) and the
// ranking of common.zig:131-167 applies unchanged (its hasNewerCommit test can never fire).
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "fpx_codec.h"

namespace fpx {

constexpr uint32_t kMaxBlocksPerHash = 4;  // FileSegment.zig:25
constexpr uint32_t kMaxDocsPerHash = 1000; // FileSegment.zig:26

template <class F> void parallel_for(size_t n, unsigned threads, F fn) {
    if (threads <= 1 || n < 2) {
        fn((size_t)0, n, 0u);
        return;
    }
    unsigned t = (unsigned)std::min<size_t>(threads, n);
    std::vector<std::thread> th;
    for (unsigned i = 0; i < t; ++i) {
        size_t a = n * i / t, b = n * (i + 1) / t;
        th.emplace_back([=] { fn(a, b, i); });
    }
    for (auto &x : th) x.join();
}

// One input segment, decoded: runs of equal hash over a flat docid array.
struct DecodedSegment {
    uint64_t commit_id = 0, merges = 0;
    bool is_file = false;
    std::vector<uint32_t> terms;     // distinct hashes, ascending
    std::vector<uint64_t> run_start; // terms.size()+1 offsets into docids
    std::vector<uint32_t> reach;     // reachable prefix length of each run
    std::vector<uint32_t> docids;    // all postings, run-major
    std::vector<uint32_t> doc_ids;   // docs map keys
    uint64_t n_unreachable = 0;
};

struct CompiledCsr {
    std::vector<uint32_t> terms;      // ascending unique
    std::vector<uint32_t> row_len;    // live reachable postings per term (> 0)
    std::vector<uint32_t> row_start4; // start of the padded row, in units of 4 docids (16 bytes)
    std::vector<uint32_t> docids;     // padded rows; 4*total4 entries
    uint32_t pad_id = 0;
    bool pad_spread = false;          // row padding (and pad_id) lies above every live docid
    uint64_t n_postings = 0, n_postings_total = 0, n_unreachable = 0, n_superseded = 0, n_out_of_range = 0;
    uint64_t max_row_len = 0;
    // dense debug copy (fpx_snapshot_csr)
    std::vector<uint64_t> dense_offsets;
    std::vector<uint32_t> dense_docids;
};

// id -> index of the newest segment whose docs map mentions it.
class NewestMention {
  public:
    void build(const std::vector<std::unique_ptr<DecodedSegment>> &segs) {
        uint64_t total = 0;
        uint32_t max_id = 0;
        for (auto &s : segs) {
            total += s->doc_ids.size();
            for (uint32_t id : s->doc_ids) max_id = std::max(max_id, id);
        }
        if ((uint64_t)max_id <= 4 * total + (1u << 24)) {
            direct_.assign((size_t)max_id + 1, 0);
            for (size_t si = 0; si < segs.size(); ++si)
                for (uint32_t id : segs[si]->doc_ids) direct_[id] = (uint16_t)(si + 1);
            use_direct_ = true;
            return;
        }
        size_t cap = 16;
        while (cap < total * 2 + 2) cap <<= 1;
        keys_.assign(cap, 0);
        vals_.assign(cap, 0);
        mask_ = cap - 1;
        for (size_t si = 0; si < segs.size(); ++si)
            for (uint32_t id : segs[si]->doc_ids) {
                size_t i = slot(id);
                while (vals_[i] && keys_[i] != id) i = (i + 1) & mask_;
                keys_[i] = id;
                vals_[i] = (uint16_t)(si + 1);
            }
    }
    // 1-based index of the newest mentioning segment, 0 if none
    uint32_t newest(uint32_t id) const {
        if (use_direct_) return id < direct_.size() ? direct_[id] : 0;
        if (keys_.empty()) return 0;
        size_t i = slot(id);
        while (vals_[i]) {
            if (keys_[i] == id) return vals_[i];
            i = (i + 1) & mask_;
        }
        return 0;
    }

  private:
    size_t slot(uint32_t id) const { return (size_t)((id * 0x9E3779B97F4A7C15ull) >> 20) & mask_; }
    bool use_direct_ = false;
    std::vector<uint16_t> direct_;
    std::vector<uint32_t> keys_;
    std::vector<uint16_t> vals_;
    size_t mask_ = 0;
};

class SnapshotCompiler {
  public:
    explicit SnapshotCompiler(unsigned threads) : threads_(threads ? threads : 1) {}

    std::string error;

    // Decode a file segment (FileSegment.zig:33-53 fields) and apply the per-hash scan caps.
    bool add_file_segment(uint64_t commit_id, uint64_t merges, uint32_t min_doc_id, uint32_t block_size,
                          const uint8_t *blocks, uint64_t num_blocks, const uint32_t *block_index,
                          const uint32_t *doc_ids, uint64_t n_docs) {
        if (!check_order(commit_id, true)) return false;
        if (block_size < kMinBlockSize || block_size > kMaxBlockSize) return fail("block_size out of range");
        if (num_blocks && (!blocks || !block_index)) return fail("null blocks / block_index");
        auto seg = std::make_unique<DecodedSegment>();
        seg->commit_id = commit_id;
        seg->merges = merges;
        seg->is_file = true;
        seg->doc_ids.assign(doc_ids, doc_ids + n_docs);

        // pass 1: item offset of every block
        std::vector<uint64_t> blk_off(num_blocks + 1, 0);
        for (uint64_t b = 0; b < num_blocks; ++b) {
            BlockHead h = read_block_head(blocks + b * (uint64_t)block_size);
            if (h.num_items == 0) return fail("empty block inside the segment");
            blk_off[b + 1] = blk_off[b] + h.num_items;
        }
        const uint64_t n_items = blk_off[num_blocks];
        std::vector<uint32_t> hashes(n_items);
        seg->docids.resize(n_items);
        // pass 2: decode every block (parallel)
        std::atomic<int> bad{0};
        parallel_for(num_blocks, threads_, [&](size_t b0, size_t b1, unsigned) {
            for (size_t b = b0; b < b1; ++b) {
                const uint8_t *blk = blocks + b * (uint64_t)block_size;
                int n = decode_block(blk, block_size, min_doc_id, hashes.data() + blk_off[b],
                                     seg->docids.data() + blk_off[b]);
                if (n < 0 || (uint64_t)n != blk_off[b + 1] - blk_off[b]) bad = 1;
                // filefmt.zig:117: block_index[b] is the hash of the block's last item
                else if (block_index[b] != hashes[blk_off[b + 1] - 1]) bad = 2;
            }
        });
        if (bad == 1) return fail("corrupt block (stream overruns the block)");
        if (bad == 2) return fail("block_index does not match the blocks");
        // pass 3: hashes must ascend across the whole segment (segment_merger.zig:131-151 order)
        parallel_for(n_items, threads_, [&](size_t a, size_t b, unsigned) {
            for (size_t i = std::max<size_t>(a, 1); i < b; ++i)
                if (hashes[i] < hashes[i - 1]) bad = 3;
        });
        if (bad == 3) return fail("hashes not ascending");
        find_runs(hashes, *seg);
        hashes = std::vector<uint32_t>();
        // pass 4: reachable prefix of every run under the 4-block / >1000-doc caps
        seg->reach.resize(seg->terms.size());
        std::vector<uint64_t> cut(threads_, 0);
        parallel_for(seg->terms.size(), threads_, [&](size_t r0, size_t r1, unsigned tid) {
            if (r0 >= r1) return;
            size_t b = (size_t)(std::upper_bound(blk_off.begin(), blk_off.end(), seg->run_start[r0]) - blk_off.begin()) - 1;
            uint64_t dropped = 0;
            for (size_t r = r0; r < r1; ++r) {
                const uint64_t s = seg->run_start[r], e = seg->run_start[r + 1];
                while (blk_off[b + 1] <= s) ++b; // first block holding this hash == lowerBound(block_index, h)
                uint64_t pos = s;
                uint32_t nb = 0;
                size_t bb = b;
                while (pos < e) { // FileSegment.zig:156-175
                    const uint64_t piece_end = std::min<uint64_t>(e, blk_off[bb + 1]);
                    pos = piece_end;
                    nb += 1;
                    ++bb;
                    if (nb >= kMaxBlocksPerHash) break;
                    if (pos - s > kMaxDocsPerHash) break;
                }
                seg->reach[r] = (uint32_t)(pos - s);
                dropped += e - pos;
            }
            cut[tid] = dropped;
        });
        for (uint64_t c : cut) seg->n_unreachable += c;
        segs_.push_back(std::move(seg));
        compiled_.reset();
        return true;
    }

    // MemorySegment.zig:21-28: items sorted by (hash<<32)|id; every match counts, no caps.
    bool add_memory_segment(uint64_t commit_id, uint64_t merges, const uint64_t *items, uint64_t n_items,
                            const uint32_t *doc_ids, uint64_t n_docs) {
        if (!check_order(commit_id, false)) return false;
        if (n_items && !items) return fail("null items");
        for (uint64_t i = 1; i < n_items; ++i)
            if (items[i] < items[i - 1]) return fail("memory segment items not sorted");
        auto seg = std::make_unique<DecodedSegment>();
        seg->commit_id = commit_id;
        seg->merges = merges;
        seg->doc_ids.assign(doc_ids, doc_ids + n_docs);
        std::vector<uint32_t> hashes(n_items);
        seg->docids.resize(n_items);
        for (uint64_t i = 0; i < n_items; ++i) {
            hashes[i] = (uint32_t)(items[i] >> 32);
            seg->docids[i] = (uint32_t)items[i];
        }
        find_runs(hashes, *seg);
        seg->reach.resize(seg->terms.size());
        for (size_t r = 0; r < seg->terms.size(); ++r) {
            uint64_t len = seg->run_start[r + 1] - seg->run_start[r];
            if (len > 0xFFFFFFFFull) return fail("posting list too long");
            seg->reach[r] = (uint32_t)len;
        }
        segs_.push_back(std::move(seg));
        compiled_.reset();
        return true;
    }

    void set_doc_range(uint32_t lo, uint32_t hi) {
        lo_ = lo;
        hi_ = hi;
        compiled_.reset();
    }
    uint32_t doc_lo() const { return lo_; }
    uint32_t doc_hi() const { return hi_; }
    size_t n_segments() const { return segs_.size(); }

    const CompiledCsr *compile() {
        if (compiled_) return compiled_.get();
        auto out = std::make_unique<CompiledCsr>();
        const bool ranged = !(lo_ == 0 && hi_ == 0);
        const bool multi = segs_.size() > 1;
        NewestMention newest;
        if (multi) newest.build(segs_);
        auto keep = [&](uint32_t id, size_t si, uint64_t &sup, uint64_t &oor) -> bool {
            if (multi) {
                if (newest.newest(id) > si + 1) { // a newer segment mentions the id
                    ++sup;
                    return false;
                }
            }
            if (ranged && !(id >= lo_ && (hi_ == 0u || id < hi_))) { // hi == 0 with lo > 0: open-ended
                ++oor;
                return false;
            }
            return true;
        };
        const bool filter = multi || ranged;

        // live count per (segment, run)
        std::vector<std::vector<uint32_t>> live(segs_.size());
        for (size_t si = 0; si < segs_.size(); ++si) {
            DecodedSegment &s = *segs_[si];
            out->n_postings_total += s.docids.size();
            out->n_unreachable += s.n_unreachable;
            live[si].resize(s.terms.size());
            if (!filter) {
                live[si] = s.reach;
                continue;
            }
            std::vector<uint64_t> sup(threads_, 0), oor(threads_, 0);
            parallel_for(s.terms.size(), threads_, [&](size_t r0, size_t r1, unsigned tid) {
                uint64_t a = 0, b = 0;
                for (size_t r = r0; r < r1; ++r) {
                    const uint32_t *d = s.docids.data() + s.run_start[r];
                    uint32_t c = 0;
                    for (uint32_t i = 0; i < s.reach[r]; ++i) c += keep(d[i], si, a, b) ? 1u : 0u;
                    live[si][r] = c;
                }
                sup[tid] = a;
                oor[tid] = b;
            });
            for (unsigned t = 0; t < threads_; ++t) {
                out->n_superseded += sup[t];
                out->n_out_of_range += oor[t];
            }
        }

        // global term directory
        std::vector<uint32_t> &terms = out->terms;
        if (segs_.size() == 1) {
            const DecodedSegment &s = *segs_[0];
            terms.reserve(s.terms.size());
            for (size_t r = 0; r < s.terms.size(); ++r)
                if (live[0][r]) terms.push_back(s.terms[r]);
        } else {
            for (size_t si = 0; si < segs_.size(); ++si)
                for (size_t r = 0; r < segs_[si]->terms.size(); ++r)
                    if (live[si][r]) terms.push_back(segs_[si]->terms[r]);
            std::sort(terms.begin(), terms.end());
            terms.erase(std::unique(terms.begin(), terms.end()), terms.end());
        }
        const size_t nt = terms.size();
        std::vector<uint64_t> len64(nt, 0);
        std::vector<uint8_t> contributors(nt, 0);
        // global row index of each (segment, run) with live postings
        std::vector<std::vector<uint32_t>> grow(segs_.size());
        for (size_t si = 0; si < segs_.size(); ++si) {
            const DecodedSegment &s = *segs_[si];
            grow[si].assign(s.terms.size(), 0xFFFFFFFFu);
            size_t g = 0;
            for (size_t r = 0; r < s.terms.size(); ++r) {
                if (!live[si][r]) continue;
                while (terms[g] < s.terms[r]) ++g;
                grow[si][r] = (uint32_t)g;
                len64[g] += live[si][r];
                if (contributors[g] < 2) contributors[g]++;
            }
        }
        out->row_len.resize(nt);
        out->row_start4.resize(nt);
        uint64_t total4 = 0;
        for (size_t g = 0; g < nt; ++g) {
            if (len64[g] > 0xFFFFFFF0ull) {
                error = "posting list too long";
                return nullptr;
            }
            out->row_len[g] = (uint32_t)len64[g];
            if (total4 > 0xFFFFFFFFull) {
                error = "snapshot too large for 32-bit row starts";
                return nullptr;
            }
            out->row_start4[g] = (uint32_t)total4;
            total4 += (len64[g] + 3) / 4;
            out->n_postings += len64[g];
            out->max_row_len = std::max<uint64_t>(out->max_row_len, len64[g]);
        }
        out->docids.assign(total4 * 4, 0);

        // fill rows: segments oldest -> newest, runs of one segment in parallel (distinct rows)
        std::vector<uint32_t> cursor(nt, 0);
        for (size_t si = 0; si < segs_.size(); ++si) {
            const DecodedSegment &s = *segs_[si];
            parallel_for(s.terms.size(), threads_, [&](size_t r0, size_t r1, unsigned) {
                uint64_t dummy1 = 0, dummy2 = 0;
                for (size_t r = r0; r < r1; ++r) {
                    const uint32_t g = grow[si][r];
                    if (g == 0xFFFFFFFFu) continue;
                    uint32_t *dst = out->docids.data() + (uint64_t)out->row_start4[g] * 4 + cursor[g];
                    const uint32_t *d = s.docids.data() + s.run_start[r];
                    if (!filter) {
                        std::memcpy(dst, d, (size_t)s.reach[r] * 4);
                        cursor[g] += s.reach[r];
                    } else {
                        uint32_t c = 0;
                        for (uint32_t i = 0; i < s.reach[r]; ++i)
                            if (keep(d[i], si, dummy1, dummy2)) dst[c++] = d[i];
                        cursor[g] += c;
                    }
                }
            });
        }
        // rows fed by several segments: restore ascending docid order
        if (multi)
            parallel_for(nt, threads_, [&](size_t g0, size_t g1, unsigned) {
                for (size_t g = g0; g < g1; ++g)
                    if (contributors[g] > 1) {
                        uint32_t *p = out->docids.data() + (uint64_t)out->row_start4[g] * 4;
                        std::sort(p, p + out->row_len[g]);
                    }
            });

        // Rows are padded to 16 bytes.  No kernel tests "is this padding?" per posting: the exact kernels mask
        // by position, and the sketch kernel simply counts the padding too — so padding must not pile up in
        // one sketch counter.  It is therefore drawn from the unused docid space above the largest live id
        // (64 Ki distinct values, varying from row to row); such values can never be live, and an exact
        // recount (which uses the rows' true lengths) gives them score 0.  pad_id, the "empty" marker of the
        // small candidate sets, is one more unused value.  If the ids reach up to 2^32 there is no such
        // range and a single unused value is used for everything (still exact, the sketch is just less sharp).
        uint32_t max_live = 0;
        {
            std::vector<uint32_t> mx(threads_, 0);
            parallel_for(nt, threads_, [&](size_t g0, size_t g1, unsigned tid) {
                uint32_t m = 0;
                for (size_t g = g0; g < g1; ++g) {
                    const uint32_t *p = out->docids.data() + (uint64_t)out->row_start4[g] * 4;
                    for (uint32_t i = 0; i < out->row_len[g]; ++i) m = std::max(m, p[i]);
                }
                mx[tid] = m;
            });
            for (uint32_t m : mx) max_live = std::max(max_live, m);
        }
        const bool spread = max_live < 0xFFFE0000u;
        out->pad_id = spread ? max_live + 1 : choose_pad(*out);
        out->pad_spread = spread;
        const uint32_t pad_base = max_live + 2;
        parallel_for(nt, threads_, [&](size_t g0, size_t g1, unsigned) {
            for (size_t g = g0; g < g1; ++g) {
                uint32_t *p = out->docids.data() + (uint64_t)out->row_start4[g] * 4;
                for (uint32_t i = out->row_len[g]; i < ((out->row_len[g] + 3) & ~3u); ++i)
                    p[i] = spread ? pad_base + (uint32_t)((g * 3 + i) & 0xFFFFu) : out->pad_id;
            }
        });
        compiled_ = std::move(out);
        return compiled_.get();
    }

    CompiledCsr *compiled() { return compiled_.get(); }
    std::unique_ptr<CompiledCsr> take_compiled() { return std::move(compiled_); }

    // Dense CSR (debug view)
    void make_dense(CompiledCsr &c) {
        if (!c.dense_offsets.empty()) return;
        const size_t nt = c.terms.size();
        c.dense_offsets.assign(nt + 1, 0);
        for (size_t g = 0; g < nt; ++g) c.dense_offsets[g + 1] = c.dense_offsets[g] + c.row_len[g];
        c.dense_docids.resize(c.dense_offsets[nt]);
        for (size_t g = 0; g < nt; ++g)
            std::memcpy(c.dense_docids.data() + c.dense_offsets[g], c.docids.data() + (uint64_t)c.row_start4[g] * 4,
                        (size_t)c.row_len[g] * 4);
    }

  private:
    bool fail(const char *msg) {
        error = msg;
        return false;
    }
    bool check_order(uint64_t commit_id, bool is_file) {
        if (!segs_.empty()) {
            const DecodedSegment &last = *segs_.back();
            if (is_file && !last.is_file) return fail("file segments must precede memory segments (Index.zig:33-41)");
            if (commit_id <= last.commit_id) return fail("segments must be added oldest to newest (ascending commit_id)");
        }
        if (segs_.size() >= 65534) return fail("too many segments");
        return true;
    }
    void find_runs(const std::vector<uint32_t> &hashes, DecodedSegment &seg) {
        const size_t n = hashes.size();
        std::vector<size_t> counts(threads_ + 1, 0);
        unsigned used = 1;
        if (n >= 2 && threads_ > 1) used = (unsigned)std::min<size_t>(threads_, n);
        parallel_for(n, used, [&](size_t a, size_t b, unsigned tid) {
            size_t c = 0;
            for (size_t i = a; i < b; ++i) c += (i == 0 || hashes[i] != hashes[i - 1]) ? 1 : 0;
            counts[tid + 1] = c;
        });
        for (unsigned t = 0; t < used; ++t) counts[t + 1] += counts[t];
        const size_t nr = counts[used];
        seg.terms.resize(nr);
        seg.run_start.resize(nr + 1);
        parallel_for(n, used, [&](size_t a, size_t b, unsigned tid) {
            size_t w = counts[tid];
            for (size_t i = a; i < b; ++i)
                if (i == 0 || hashes[i] != hashes[i - 1]) {
                    seg.terms[w] = hashes[i];
                    seg.run_start[w] = i;
                    ++w;
                }
        });
        seg.run_start[nr] = n;
    }
    uint32_t choose_pad(const CompiledCsr &c) {
        const size_t nt = c.terms.size();
        for (uint64_t base = 0; base < (1ull << 32); base += 64) {
            // candidates base..base+63 (first round also tries 0xFFFFFFFF implicitly later)
            std::vector<uint64_t> used(threads_, 0);
            parallel_for(nt, threads_, [&](size_t g0, size_t g1, unsigned tid) {
                uint64_t m = 0;
                for (size_t g = g0; g < g1; ++g) {
                    const uint32_t *p = c.docids.data() + (uint64_t)c.row_start4[g] * 4;
                    for (uint32_t i = 0; i < c.row_len[g]; ++i) {
                        uint64_t d = p[i];
                        if (d >= base && d < base + 64) m |= 1ull << (d - base);
                    }
                }
                used[tid] = m;
            });
            uint64_t m = 0;
            for (uint64_t u : used) m |= u;
            if (m != ~0ull)
                for (unsigned k = 0; k < 64; ++k)
                    if (!(m >> k & 1)) return (uint32_t)(base + k);
        }
        return 0; // unreachable: fewer than 2^32 postings can not cover every u32
    }

    unsigned threads_;
    uint32_t lo_ = 0, hi_ = 0;
    std::vector<std::unique_ptr<DecodedSegment>> segs_;
    std::unique_ptr<CompiledCsr> compiled_;
};

} // namespace fpx
