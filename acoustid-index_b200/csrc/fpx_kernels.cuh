// fpx_kernels.cuh — device-side data structures shared by the kernels and the host API.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fpx {

// Term directory entry (open addressing, linear probing, 16 bytes = one 128-bit load).
struct TermEntry {
    uint32_t term;
    uint32_t len;    // postings in the row (> 0)
    uint32_t start4; // row start in units of 4 docids (16 bytes)
    uint32_t used;   // 0 = empty slot
};

struct SnapshotDev {
    const TermEntry *table;
    uint32_t table_mask;
    uint32_t table_shift; // 32 - log2(capacity)
    const uint32_t *docids; // padded rows
    uint32_t pad_id; // a docid no posting uses: "empty" marker of candidate sets (row padding uses other unused
                     // docids, see fpx_snapshot_host.h)
    uint32_t pad_spread; // 1: pad_id and all row padding are larger than every live docid
};

// Order of the docids inside a row in HBM: ascending row_key(d) = d * kRowMult (an odd multiplier: a bijection on u32),
// the hash the sketch kernel counts with.  Its top 15 bits are the sketch counter of d (32768 8-bit counters, four per
// 32-bit word: word = h[31:19], byte = h[18:17]), so the postings of one counter are one contiguous range of every
// row — the resolvers of search_find_kernel binary-search it with a multiply and a compare per probe — equal docids
// stay adjacent, and a row is still searchable by key.  (Round 1 ordered rows by the shared-memory bank of the counter
// word to save a fifth of the atomics' bank conflicts; that measured -0.5 % and would cost the search a bit
// permutation per probe.)
constexpr uint32_t kRowMult = 0x9E3779B1u;
constexpr uint32_t row_inv32(uint32_t a) {
    uint32_t x = a;
    for (int i = 0; i < 5; ++i) x *= 2u - a * x;
    return x;
}
constexpr uint32_t kRowMultInv = row_inv32(kRowMult);
__host__ __device__ __forceinline__ uint32_t row_key(uint32_t d) { return d * kRowMult; }
__host__ __device__ __forceinline__ uint32_t row_key_inv(uint32_t k) { return k * kRowMultInv; }

struct SearchOpts { // == fpx_search_opts
    uint32_t max_results, min_score, min_score_pct;
};

// One prepared query, as the search kernels consume it (32 bytes).
struct WorkItem {
    uint32_t q;         // query index in the batch
    uint32_t rows_off;  // first row descriptor in BatchArgs::rows
    uint32_t n_rows;    // unique query terms present in the snapshot
    uint32_t total4;    // sum over rows of ceil(len/4): padded posting volume in 16-byte units
    uint32_t postings;  // sum of row lengths (saturating)
    uint32_t k_eff;     // min(max_results, k_stride)
    uint32_t min_score;
    uint32_t min_score_pct;
};

// Work classes.
//   0      sketch path: TMA-staged rows (<= 40832 bytes = 10208 padded postings per query), u8 count sketch, hot
//          counters resolved exactly in the staged rows (needs 2 <= min_score <= 128)
//   1..3   exact shared-memory count table of 2^13 / 2^14 / 2^15 packed slots
//   4      global-memory table: whatever the others cannot represent exactly
constexpr int kNumClasses = 5;
constexpr int kSketchClass = 0;
constexpr int kWideClass = 4;

struct BatchCounters {
    uint32_t qcount[kNumClasses];
    uint32_t qhead[kNumClasses];
    uint32_t long_count;
    uint32_t long_head;
    uint32_t error; // FPX_UNSUPPORTED etc. raised on device
    uint32_t pad;
};

struct DeviceStats { // accumulated across batches (profiling)
    unsigned long long queries, unique_terms, postings, results, wide_queries, overflow_requeues, sketch_queries;
    unsigned long long dbg[16]; // phase timers of the sketch kernel (CTA 0 only), FPX_DEBUG_ABLATE bit 9
};

struct BatchArgs {
    SnapshotDev snap;
    uint32_t n_queries;
    uint32_t k_stride;
    const uint32_t *terms;
    const uint64_t *term_offsets;
    uint64_t term_base; // subtracted from term_offsets (chunked host batches)
    uint64_t n_terms_total; // terms of the batch: term_offsets must stay within [term_base, term_base + n_terms_total]
    const SearchOpts *opts;
    uint32_t *out_ids, *out_scores, *out_counts;
    // workspace
    uint4 *rows;            // per query at [term_offsets[q]-term_base ...): {start4, len, off4, 0}; off4 = start of
                            // the row inside a stage that holds the query's rows back to back (16-byte units)
    WorkItem *items;        // kNumClasses * n_queries
    uint32_t *long_queue;   // n_queries
    BatchCounters *counters;
    DeviceStats *stats;
    unsigned long long *wide_tables; // per wide CTA: wide_cap 64-bit slots
    uint32_t wide_cap_log2;
    uint32_t use_sketch;    // 0 disables class 0 (A/B testing)
    uint32_t debug;         // ablation bits for profiling runs (results are wrong when set): see fpx_kernels.cu
};

constexpr uint32_t kWarpQueryTerms = 128; // queries up to this many raw terms are prepared by one warp
constexpr uint32_t kMaxQueryTerms = 8192; // FPX_MAX_QUERY_TERMS
constexpr uint32_t kFastKbuf = 512;       // candidate buffer of the shared-memory paths
constexpr uint32_t kWideKbuf = 2048;      // candidate buffer of the global-memory path
constexpr uint32_t kMaxResults = 1024;    // FPX_MAX_RESULTS
constexpr uint32_t kRowsChunk = 256;      // row descriptors staged per round
constexpr uint32_t kStageU4 = 2552;       // sketch path: one query's padded rows must fit a 40832-byte stage (four of them and
                                          // the two 32 KB sketches are all the shared memory of an SM)
constexpr uint32_t kSketchMaxRows = 128;  // sketch path: row descriptors live in producer registers

void launch_build_table(TermEntry *table, uint32_t log2cap, const uint32_t *terms, const uint32_t *lens,
                        const uint32_t *start4, uint64_t n_terms, cudaStream_t st);
void launch_prepare(const BatchArgs &a, cudaStream_t st);
void launch_prepare_long(const BatchArgs &a, cudaStream_t st, int n_sms);
void launch_search_sketch(const BatchArgs &a, cudaStream_t st, int n_sms); // both sketch classes
void launch_search_class(const BatchArgs &a, int cls, cudaStream_t st, int n_sms);
void launch_search_wide(const BatchArgs &a, cudaStream_t st, int n_ctas);
// pack the k_stride-wide result arrays of n queries: out_counts[q] and the (id, score) pairs back to back
// (out_* may be mapped pinned host memory); offsets = n+1 words of device scratch, offsets[n] = number of pairs;
// pairs beyond `capacity` are dropped
void launch_result_pack(const uint32_t *ids, const uint32_t *scores, const uint32_t *counts, uint32_t *offsets, uint32_t n,
                        uint32_t k_stride, uint32_t *out_counts, uint2 *out_pairs, cudaStream_t st,
                        uint32_t capacity = 0xFFFFFFFFu);
// merge n_shards packed result blocks (layout of launch_result_pack with out_counts = block, offsets = block + n,
// pairs = block + 2n + 2; blocks stride_words apart) into k_stride-wide arrays: global top-k + cutoffs per query
void launch_merge_packed_shards(const uint32_t *packed, uint64_t stride_words, uint32_t n_shards, uint32_t n,
                                const SearchOpts *opts, uint32_t k_stride, uint32_t *out_ids, uint32_t *out_scores,
                                uint32_t *out_counts, cudaStream_t st);
cudaError_t configure_kernels();
int wide_ctas(int n_sms);

} // namespace fpx
