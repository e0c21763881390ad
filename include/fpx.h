/*
 * fpx.h — C ABI of the B200-native `_search` path for acoustid-index (fpindex).
 *
 * This is the drop-in boundary.  The reference (Zig) has no FFI; the seams this ABI is
 * designed to be bound at are (paths relative to the reference root):
 *
 *   - src/Index.zig:469-485  Index.swapSnapshot        -> fpx_snapshot_begin / add_..._segment / commit
 *   - src/Index.zig:57-63    Segments.deinit           -> fpx_snapshot_release
 *   - src/Index.zig:430-434  Index.acquireReader       -> fpx_snapshot_acquire
 *   - src/Index.zig:170-177  IndexReader.search  }
 *   - src/common.zig:131-167 SearchResults.finish }    -> fpx_search / fpx_search_batch
 *   - src/MultiIndex.zig:302-306 option mapping        -> fpx_default_min_score
 *
 * INTEGRATION.md shows the Zig `extern fn` stub for each entry point.
 *
 * Conventions: plain pointers and sizes, no exceptions cross the boundary, every function
 * returns an fpx_status; fpx_last_error_message() gives a thread-local description of the
 * last failure on the calling thread.  All input buffers are borrowed for the duration of
 * the call only.  Handles are opaque; snapshots are immutable and atomically refcounted, so
 * fpx_search* may be called concurrently from many threads on one snapshot while another
 * thread commits a newer one (mirrors the reference's lock-free readers, MultiIndex.zig:5-10).
 */
#ifndef FPX_H
#define FPX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPX_ABI_VERSION 1u

typedef int32_t fpx_status;
enum {
    FPX_OK = 0,
    FPX_OUT_OF_MEMORY = 1,
    FPX_INVALID_ARGUMENT = 2,
    FPX_INVALID_SEGMENT = 3,     /* error.InvalidSegment / ChecksumMismatch, filefmt.zig:235-284 */
    FPX_TIMEOUT = 4,             /* error.SearchTimeout, MultiIndex.zig:320 */
    FPX_CUDA_ERROR = 5,
    FPX_BACKEND_UNAVAILABLE = 6, /* no CUDA device / driver: the caller keeps its CPU path */
    FPX_UNSUPPORTED = 7          /* query outside the documented limits below */
};

/* Documented limits of the device path (callers keep the CPU path beyond). */
#define FPX_MAX_QUERY_TERMS 8192u /* raw terms per query */
#define FPX_MAX_RESULTS 1024u     /* effective per-query limit = min(max_results, k_stride) */

typedef struct fpx_ctx fpx_ctx;
typedef struct fpx_snapshot_builder fpx_snapshot_builder;
typedef struct fpx_snapshot fpx_snapshot;

typedef struct fpx_config {
    int32_t device;            /* CUDA device ordinal; -1 = current device */
    uint32_t host_threads;     /* threads for snapshot compilation; 0 = hardware concurrency */
    uint32_t chunk_queries;    /* queries per pipelined H2D/compute/D2H chunk in fpx_search_batch; 0 = default */
    uint32_t flags;            /* FPX_FLAG_* */
} fpx_config;
#define FPX_FLAG_PROFILE 1u    /* record CUDA events around every kernel (fpx_profile_read) */
#define FPX_FLAG_HOST_ONLY 2u  /* no device: only snapshot compilation / introspection work (CPU tests) */
#define FPX_FLAG_NO_SKETCH 4u  /* route every query through the exact count-table kernels (A/B testing) */
#define FPX_FLAG_HOST_BUILD 8u /* compile snapshots on host threads instead of on the device */

/* An immutable file segment exactly as FileSegment holds it in RAM (FileSegment.zig:33-53). */
typedef struct fpx_file_segment {
    uint64_t commit_id;          /* SegmentInfo.commit_id (segment.zig:23-26) */
    uint64_t merges;             /* SegmentInfo.merges */
    uint32_t min_doc_id;         /* FileSegment.min_doc_id: the docid delta base (block.zig:68) */
    uint32_t block_size;         /* FileSegment.block_size, 64..4096 (block.zig:41-42) */
    const uint8_t *blocks;       /* FileSegment.blocks: num_blocks*block_size bytes */
    uint64_t num_blocks;         /* FileSegment.num_blocks (terminator block excluded) */
    const uint32_t *block_index; /* FileSegment.block_index: max hash per block (filefmt.zig:117) */
    const uint32_t *doc_ids;     /* keys of FileSegment.docs (any order) */
    const uint8_t *doc_alive;    /* values of FileSegment.docs: 1 insert, 0 tombstone */
    uint64_t n_docs;
} fpx_file_segment;

/* A memory segment as MemorySegment holds it (MemorySegment.zig:21-28). */
typedef struct fpx_memory_segment {
    uint64_t commit_id;
    uint64_t merges;
    const uint64_t *items; /* MemorySegment.items: packed Item = (hash<<32)|id, ascending (segment.zig:87-106) */
    uint64_t n_items;
    const uint32_t *doc_ids;
    const uint8_t *doc_alive;
    uint64_t n_docs;
} fpx_memory_segment;

/* common.zig:50-54 SearchOptions (already resolved: see fpx_default_min_score). */
typedef struct fpx_search_opts {
    uint32_t max_results;
    uint32_t min_score;
    uint32_t min_score_pct;
} fpx_search_opts;

typedef struct fpx_snapshot_info {
    uint64_t n_segments;
    uint64_t n_terms;            /* CSR rows */
    uint64_t n_postings;         /* live, reachable postings kept */
    uint64_t n_postings_total;   /* postings seen in the input segments */
    uint64_t n_dropped_unreachable; /* cut by the 4-block / >1000-doc scan caps (FileSegment.zig:173-174) */
    uint64_t n_dropped_superseded;  /* newer segment mentions the id (Index.zig:133-149) */
    uint64_t n_dropped_out_of_range;/* outside this shard's docid range */
    uint64_t device_bytes;       /* HBM held by this snapshot */
    uint64_t max_row_len;
    uint32_t pad_id;             /* docid value used to pad rows to 16 bytes (not a live id) */
    uint32_t table_log2;         /* log2 of the term hash table capacity */
    uint32_t doc_lo, doc_hi;     /* docid range [lo, hi) of this shard (0, 0 = everything) */
} fpx_snapshot_info;

/* Host view of the compiled CSR (debug / tests; valid until the builder is committed or aborted). */
typedef struct fpx_csr_view {
    uint64_t n_terms;
    const uint32_t *terms;       /* ascending, unique */
    const uint64_t *row_offsets; /* n_terms+1, in docids (unpadded, dense) */
    const uint32_t *docids;      /* row r = docids[row_offsets[r] .. row_offsets[r+1]), ascending */
} fpx_csr_view;

typedef struct fpx_profile {
    /* accumulated since the last fpx_profile_reset, FPX_FLAG_PROFILE only */
    double prepare_ms;  uint64_t prepare_launches;
    double sketch_ms;   uint64_t sketch_launches;   /* search_sketch_kernel: TMA gather + count sketch + top-k */
    double search_ms;   uint64_t search_launches;   /* search_smem_kernel<13|14|15>: exact count-table path */
    double wide_ms;     uint64_t wide_launches;     /* search_wide_kernel: overflow / oversized queries */
    double h2d_ms, d2h_ms;
    uint64_t queries, unique_terms, postings, results;
    uint64_t sketch_queries, wide_queries, overflow_requeues;
    uint64_t h2d_bytes, d2h_bytes;
} fpx_profile;

/* ---- lifecycle ---- */
uint32_t fpx_abi_version(void);
const char *fpx_last_error_message(void);
fpx_status fpx_init(const fpx_config *config /* may be NULL */, fpx_ctx **out);
void fpx_shutdown(fpx_ctx *ctx);

/* ---- snapshot build: called where Index.swapSnapshot installs a new Segments ---- */
fpx_status fpx_snapshot_begin(fpx_ctx *ctx, fpx_snapshot_builder **out);
/* Segments must be added oldest -> newest, all file segments before all memory segments
 * (Index.zig:33-41).  Input is fully consumed (uploaded or decoded) during the call. */
fpx_status fpx_snapshot_add_file_segment(fpx_snapshot_builder *b, const fpx_file_segment *seg);
fpx_status fpx_snapshot_add_memory_segment(fpx_snapshot_builder *b, const fpx_memory_segment *seg);
/* Restrict the snapshot to docids in [lo, hi) (multi-GPU docid-range sharding).  The scan caps and
 * supersession rules are applied on the whole snapshot first, so shards union to the full result.
 * (0, 0) = no restriction; hi = 0 with lo > 0 = open-ended, [lo, 2^32): the last shard, so that the docid
 * 0xFFFFFFFF belongs to a shard too. */
fpx_status fpx_snapshot_set_doc_range(fpx_snapshot_builder *b, uint32_t lo, uint32_t hi);
/* Compile to CSR on the host (idempotent).  Host-built snapshots only (FPX_FLAG_HOST_BUILD / FPX_FLAG_HOST_ONLY):
 * by default a device context decodes the segments and assembles the rows on the GPU at commit. */
fpx_status fpx_snapshot_compile(fpx_snapshot_builder *b);
fpx_status fpx_snapshot_csr(fpx_snapshot_builder *b, fpx_csr_view *out);
/* Upload to HBM; consumes the builder on success.  refcount starts at 1. */
fpx_status fpx_snapshot_commit(fpx_snapshot_builder *b, fpx_snapshot **out);
void fpx_snapshot_abort(fpx_snapshot_builder *b);

fpx_status fpx_snapshot_acquire(fpx_snapshot *s);
fpx_status fpx_snapshot_release(fpx_snapshot *s);
fpx_status fpx_snapshot_get_info(const fpx_snapshot *s, fpx_snapshot_info *out);
/* The row of `term` as it lies in HBM (without padding): debug / tests.  *out_len receives the row length (0 if the
 * term is absent); FPX_INVALID_ARGUMENT if it does not fit `capacity`. */
fpx_status fpx_snapshot_read_row(const fpx_snapshot *s, uint32_t term, uint32_t *out_docids, uint64_t capacity,
                                 uint64_t *out_len);
/* Row length of each term (0 if absent) from the host-side copy of the term directory. */
fpx_status fpx_snapshot_row_lengths(const fpx_snapshot *s, const uint32_t *terms, uint64_t n,
                                    uint32_t *out_lengths);

/* ---- search: called where IndexReader.search + SearchResults.finish are ---- */
/* MultiIndex.zig:304: min_score = (raw query length + 19) / 20 when the request has none. */
uint32_t fpx_default_min_score(uint64_t raw_query_len);

/* One query (IndexReader.search + finish).  `terms` is NOT modified (the reference sorts in place). */
fpx_status fpx_search(fpx_snapshot *s, const uint32_t *terms, uint64_t n_terms,
                      const fpx_search_opts *opts, uint32_t *out_ids, uint32_t *out_scores,
                      uint32_t capacity, uint32_t *out_count);

/* A batch of independent queries; host buffers.  Query q = terms[term_offsets[q] .. term_offsets[q+1]).
 * Results of query q go to out_ids/out_scores[q*k_stride ...], count to out_counts[q]; at most
 * min(opts[q].max_results, k_stride) results per query (a prefix of the reference's list). */
fpx_status fpx_search_batch(fpx_snapshot *s, uint64_t n_queries, const uint32_t *terms,
                            const uint64_t *term_offsets, const fpx_search_opts *opts,
                            uint32_t k_stride, uint32_t *out_ids, uint32_t *out_scores,
                            uint32_t *out_counts);

/* The same search with the results as the reference hands them out — a list per query (SearchResponse.results,
 * api.zig:56-72; MultiIndex.zig:327-329 copies exactly the found results into the request arena) instead of
 * k_stride-wide arrays: out_counts[q] results for query q, whose (id, score) pairs are the words
 * out_pairs[2*o ...) with o = out_counts[0] + ... + out_counts[q-1].  For a 100 K-query batch of the metric's workload
 * that is ~1.1 MB coming back from the GPU instead of 32 MB.  capacity_pairs = room in out_pairs (pairs);
 * *out_n_pairs = pairs needed; FPX_INVALID_ARGUMENT if that exceeds the capacity (counts complete, pairs cut).
 * k_stride still bounds the results per query. */
fpx_status fpx_search_batch_packed(fpx_snapshot *s, uint64_t n_queries, const uint32_t *terms,
                                   const uint64_t *term_offsets, const fpx_search_opts *opts, uint32_t k_stride,
                                   uint32_t *out_counts, uint32_t *out_pairs, uint64_t capacity_pairs,
                                   uint64_t *out_n_pairs);

/* fpx_search_batch with the request's deadline (api.SearchRequest.timeout, api.zig:7-8; MultiIndex.zig:311-322 cancels
 * the search and returns error.SearchTimeout): FPX_TIMEOUT when the batch is not answered within timeout_ms
 * (0 = no deadline); the output arrays are then undefined.  The call still returns only after the GPU work it
 * enqueued has drained (it writes the library's own buffers). */
fpx_status fpx_search_batch_timeout(fpx_snapshot *s, uint64_t n_queries, const uint32_t *terms,
                                    const uint64_t *term_offsets, const fpx_search_opts *opts, uint32_t k_stride,
                                    uint32_t *out_ids, uint32_t *out_scores, uint32_t *out_counts,
                                    uint32_t timeout_ms);

/* The batch with all buffers already in device memory, enqueued on `cuda_stream` (a cudaStream_t; NULL = the legacy
 * default stream).  fpx_search_batch_device_async never waits for the device: the caller states where the batch's
 * terms lie — term_base = term_offsets[0], n_terms_total = term_offsets[n_queries] - term_offsets[0] (this sizes the
 * row workspace; offsets outside that window are rejected on the device) — and results are ready when the stream
 * reaches this point, so calls can be queued back to back or captured in a CUDA graph.  What the kernels raise cannot
 * come back as a return value: *d_status (device memory, optional) receives 0, or FPX_UNSUPPORTED /
 * FPX_INVALID_ARGUMENT when a query lies outside the limits below / has bad offsets; such a query has count 0.
 * fpx_search_batch_device reads the two offsets back itself (one stream synchronisation) and has no status word. */
fpx_status fpx_search_batch_device_async(fpx_snapshot *s, uint64_t n_queries, uint64_t term_base,
                                         uint64_t n_terms_total, const uint32_t *d_terms,
                                         const uint64_t *d_term_offsets, const fpx_search_opts *d_opts,
                                         uint32_t k_stride, uint32_t *d_out_ids, uint32_t *d_out_scores,
                                         uint32_t *d_out_counts, uint32_t *d_status, void *cuda_stream);
fpx_status fpx_search_batch_device(fpx_snapshot *s, uint64_t n_queries, const uint32_t *d_terms,
                                   const uint64_t *d_term_offsets, const fpx_search_opts *d_opts,
                                   uint32_t k_stride, uint32_t *d_out_ids, uint32_t *d_out_scores,
                                   uint32_t *d_out_counts, void *cuda_stream);

/* Pack the k_stride-wide device result arrays of a batch for an exchange between GPUs (most queries return one
 * or two results): d_packed, 2*n_queries + 2 + 2*capacity_pairs u32 words, receives
 *   [0, n) counts | [n, 2n] row offsets into the pairs, word 2n = number of pairs needed | word 2n+1 unused |
 *   capacity_pairs x (id, score).
 * Pairs beyond capacity_pairs are dropped: the receiver compares word 2n with its capacity.  Asynchronous on
 * `cuda_stream`. */
fpx_status fpx_pack_results_device(uint64_t n_queries, uint32_t k_stride, const uint32_t *d_ids,
                                   const uint32_t *d_scores, const uint32_t *d_counts, uint32_t *d_packed,
                                   uint32_t capacity_pairs, void *cuda_stream);

/* Docid-range sharded corpus, device side: merge the shards' packed top-k lists of a batch (n_shards blocks in the
 * layout of fpx_pack_results_device, shard g's block at d_packed + g * shard_stride_words — e.g. the receive buffer of
 * an ncclAllGather of every rank's packed block, cut to the largest rank's size) into k_stride-wide arrays: per query
 * the global best min(max_results, k_stride) under (score desc, id asc), then the relative cutoff anchored on the
 * global best (common.zig:153-166).  Shards must have been searched with min_score_pct = 0 and hold disjoint docid
 * ranges; every block's pairs must be complete (word 2n <= its capacity).  One warp per query, lane g walks shard g:
 * at most 32 shards.  Asynchronous on `cuda_stream`. */
fpx_status fpx_merge_packed_shards_device(uint32_t n_shards, uint64_t n_queries, const uint32_t *d_packed,
                                          uint64_t shard_stride_words, const fpx_search_opts *d_opts,
                                          uint32_t k_stride, uint32_t *d_out_ids, uint32_t *d_out_scores,
                                          uint32_t *d_out_counts, void *cuda_stream);

/* Merge per-shard top-k lists (docid-range sharded corpus): for each query take the global best
 * min(max_results, k_stride) under (score desc, id asc), then apply the relative cutoff anchored on the
 * global best (common.zig:153-166).  Shards must have been searched with min_score_pct = 0.
 * Host buffers; shard g's lists at ids[g*n_queries*k_stride ...]. */
fpx_status fpx_merge_shard_results(uint32_t n_shards, uint64_t n_queries, uint32_t k_stride,
                                   const uint32_t *ids, const uint32_t *scores, const uint32_t *counts,
                                   const fpx_search_opts *opts, uint32_t *out_ids, uint32_t *out_scores,
                                   uint32_t *out_counts);

/* ---- request micro-batcher: the single-query seam of MultiIndex.search (MultiIndex.zig:287-330) ----
 * Many host threads (the reference runs one coroutine per request on `executors=.auto` OS threads,
 * main.zig:272-276) call fpx_batcher_search concurrently; a worker thread turns whatever has accumulated into
 * one fpx_search_batch call on the snapshot that is current at that moment.  fpx_batcher_set_snapshot is the
 * Index.swapSnapshot hook: it never waits for searches, batches in flight keep the old snapshot alive. */
typedef struct fpx_batcher fpx_batcher;
typedef struct fpx_batcher_config {
    uint32_t max_batch;   /* queries per GPU batch; 0 = 4096 */
    uint32_t max_wait_us; /* how long an idle worker lets a batch fill before launching it (default 100) */
} fpx_batcher_config;
typedef struct fpx_batcher_stats {
    uint64_t batches, queries, max_batch_seen, timeouts;
} fpx_batcher_stats;
fpx_status fpx_batcher_create(fpx_ctx *ctx, const fpx_batcher_config *config /* may be NULL */, fpx_batcher **out);
fpx_status fpx_batcher_set_snapshot(fpx_batcher *b, fpx_snapshot *snapshot /* NULL = none */);
/* One query, blocking, thread-safe.  timeout_ms as api.SearchRequest.timeout (api.zig:7-8): 0 = no bound;
 * FPX_TIMEOUT mirrors error.SearchTimeout (MultiIndex.zig:320).  At most `capacity` results. */
fpx_status fpx_batcher_search(fpx_batcher *b, const uint32_t *terms, uint64_t n_terms, const fpx_search_opts *opts,
                              uint32_t timeout_ms, uint32_t *out_ids, uint32_t *out_scores, uint32_t capacity,
                              uint32_t *out_count);
fpx_status fpx_batcher_get_stats(fpx_batcher *b, fpx_batcher_stats *out);
/* Answers what is still queued, then stops the worker.  No fpx_batcher_search may be running or start. */
void fpx_batcher_destroy(fpx_batcher *b);

/* ---- wire codecs of the search call (for a search-only endpoint without the Zig host) ----
 * api.SearchRequest / api.SearchResponse (api.zig:14-27, 56-72) as the HTTP server decodes and encodes them
 * (server.zig:84-142, 189-196: msgpack with one-letter keys, JSON with field names; limit clamped to [1, 100],
 * timeout to <= 10000) and the legacy line protocol's fingerprint / result formats (legacy.zig:185-210, 286-296).
 * Malformed input -> FPX_INVALID_ARGUMENT (error.BadRequest).  Buffers handed out are freed with fpx_wire_free. */
#define FPX_WIRE_JSON 0u
#define FPX_WIRE_MSGPACK 1u
typedef struct fpx_wire_search_request {
    uint32_t *query;       /* n_terms raw terms (fpx_wire_free) */
    uint64_t n_terms;
    uint32_t timeout;      /* ms, 0 = none; default 500 */
    uint32_t limit;        /* default 40 */
    uint32_t has_min_score, min_score; /* null -> fpx_default_min_score(n_terms) */
    uint32_t score_pct;    /* default 10 */
} fpx_wire_search_request;
fpx_status fpx_wire_decode_search_request(uint32_t format, const uint8_t *data, uint64_t size,
                                          fpx_wire_search_request *out);
fpx_status fpx_wire_encode_search_response(uint32_t format, const uint32_t *ids, const uint32_t *scores, uint32_t n,
                                           uint8_t **out, uint64_t *out_size);
fpx_status fpx_legacy_parse_fingerprint(const char *text, uint64_t len, uint32_t **out_terms, uint64_t *out_n);
fpx_status fpx_legacy_format_results(const uint32_t *ids, const uint32_t *scores, uint32_t n, uint8_t **out,
                                     uint64_t *out_size);
void fpx_wire_free(void *p);

/* Queries per pipelined H2D / compute / D2H chunk of fpx_search_batch (same as fpx_config.chunk_queries);
 * takes effect for calls that start afterwards. */
fpx_status fpx_set_chunk_queries(fpx_ctx *ctx, uint32_t chunk_queries);

/* ---- profiling ---- */
/* Turn event recording / device counters (FPX_FLAG_PROFILE) on or off for calls that start afterwards. */
fpx_status fpx_set_profile(fpx_ctx *ctx, int enabled);
fpx_status fpx_profile_reset(fpx_ctx *ctx);
fpx_status fpx_profile_read(fpx_ctx *ctx, fpx_profile *out);
/* Profiling only: kernel variant / ablation bits (same meaning as the FPX_DEBUG_ABLATE environment variable).
 *   ablations of the hot kernel, WRONG RESULTS, timing only: bit 0 no counting, 1 no recount, 3 no row copies, 4 no sketch
 *     clear
 *   measurements, results unchanged: bit 9 the hot kernel's instance with phase timers (printed by fpx_profile_read),
 *     bit 11 per-chunk timeline of every host-buffer call on stderr, bit 12 packed result path even for pinned output
 *     arrays, bits 16..18 chunk schedule of the host-buffer call, bits 24..27 warp split of the hot kernel
 *     (csrc/fpx_kernels.cu FPX_FIND_CONFIGS) */
fpx_status fpx_debug_set(fpx_ctx *ctx, uint32_t bits);

#ifdef __cplusplus
}
#endif
#endif /* FPX_H */
