/*
 * fpindex_oracle.h — C interface of the CPU ORACLE.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  This library is a CPU restatement of
 * the acoustid-index (fpindex) `_search` hot path, written from the reference's
 * Zig sources (cited per function in fpindex_oracle.cpp).  It exists only so
 * that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs can check and time the reference algorithm.  Nothing
 * under acoustid-index_b200/ links, loads or calls it.
 *
 * Parity pinning: the reference (Zig 0.16 + four network-fetched deps) cannot be
 * built in this environment, so there is no oracle/_ref.  The oracle is pinned
 * against every known-answer vector the reference's own tests hold for this path
 * (tests/test_oracle_kat.py lists them with file:line).  Behaviours no reference
 * test pins (scan caps, max_results cut, min_score_pct anchor, tie-breaks beyond
 * a 2-way tie, score*pct overflow) are restated from the code and marked
 * "parity unpinned" in DESIGN.md.
 */
#ifndef FPINDEX_ORACLE_H
#define FPINDEX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_index orc_index;

/* change kinds for orc_update (src/change.zig: insert / delete) */
#define ORC_INSERT 0u
#define ORC_DELETE 1u

typedef struct orc_file_segment_view {
    uint64_t commit_id;
    uint64_t merges;
    uint32_t min_doc_id;
    uint32_t max_doc_id;
    uint32_t block_size;
    uint32_t _pad;
    uint64_t num_blocks;
    uint64_t num_items;
    const uint8_t *blocks;       /* num_blocks*block_size bytes + one all-zero terminator block */
    const uint32_t *block_index; /* num_blocks entries: max hash per block */
    const uint32_t *doc_ids;     /* n_docs keys of the docs map (ascending) */
    const uint8_t *doc_alive;    /* n_docs values (1 = insert, 0 = tombstone) */
    uint64_t n_docs;
} orc_file_segment_view;

typedef struct orc_memory_segment_view {
    uint64_t commit_id;
    uint64_t merges;
    uint32_t min_doc_id;
    uint32_t max_doc_id;
    const uint64_t *items; /* (hash<<32)|id, ascending */
    uint64_t n_items;
    const uint32_t *doc_ids;
    const uint8_t *doc_alive;
    uint64_t n_docs;
} orc_memory_segment_view;

orc_index *orc_index_new(uint32_t block_size /* 0 = 512 */);
void orc_index_free(orc_index *);

/* Index.update: build one memory segment from a change batch, commit_id = ++last. */
int orc_update(orc_index *, size_t n_changes, const uint8_t *kinds, const uint32_t *ids,
               const uint64_t *hash_offsets /* n_changes+1 */, const uint32_t *hashes);
/* Index.checkpoint(force): all memory segments -> one new file segment. */
int orc_checkpoint(orc_index *);
/* Index.mergeMemory / mergeFiles on an explicit adjacent range. */
int orc_merge_memory(orc_index *, size_t lo, size_t count);
int orc_merge_files(orc_index *, size_t lo, size_t count);
/* Bulk load: one file segment straight from sorted items (hash<<32|id) + docs map. */
int orc_add_file_segment_sorted(orc_index *, const uint64_t *items, size_t n_items,
                                const uint32_t *doc_ids, const uint8_t *doc_alive, size_t n_docs);
/* Adopt externally encoded segment bytes (borrowed; must stay valid; `blocks` must be
 * followed by block_size readable zero bytes = the terminator block). */
int orc_adopt_file_segment(orc_index *, uint64_t commit_id, uint64_t merges, uint32_t block_size,
                           const uint8_t *blocks, size_t n_blocks, const uint32_t *block_index,
                           const uint32_t *doc_ids, const uint8_t *doc_alive, size_t n_docs);

size_t orc_num_file_segments(const orc_index *);
size_t orc_num_memory_segments(const orc_index *);
int orc_file_segment(const orc_index *, size_t i, orc_file_segment_view *out);
int orc_memory_segment(const orc_index *, size_t i, orc_memory_segment_view *out);

/* IndexReader.search + SearchResults.finish.  Returns number of results (<= cap), <0 on error. */
int64_t orc_search(const orc_index *, const uint32_t *query, size_t n_terms, uint32_t max_results,
                   uint32_t min_score, uint32_t min_score_pct, uint32_t *out_ids,
                   uint32_t *out_scores, size_t cap);

/* Batch over n_threads host threads (one independent query stream per thread).
 * opts = per-query {max_results, min_score, min_score_pct}.  Returns seconds spent in the
 * search loop (wall), <0 on error. */
double orc_search_batch(const orc_index *, size_t n_queries, const uint32_t *terms,
                        const uint64_t *term_offsets, const uint32_t *opts3, uint32_t k_stride,
                        uint32_t *out_ids, uint32_t *out_scores, uint32_t *out_counts,
                        unsigned n_threads);

/* --- codec entry points exposed for the known-answer tests --- */
size_t orc_svb_encode_quad_0124(const uint32_t in[4], uint8_t *out_data, uint8_t *out_control);
size_t orc_svb_encode_quad_1234(const uint32_t in[4], uint8_t *out_data, uint8_t *out_control);
/* variant: 0 = 0124, 1 = 1234, 2 = 0124_minus1.  `in` needs 16 readable bytes. */
size_t orc_svb_decode_quad(int variant, uint8_t control, const uint8_t *in, uint32_t out[4]);
size_t orc_svb_decode_quad_delta(int variant, uint8_t control, const uint8_t *in, uint32_t out[4],
                                 uint32_t carry);
void orc_svb_delta_decode_in_place(uint32_t *data, size_t n, uint32_t first_value);
/* streamvbyte.decodeValues; `in` must have 16 readable bytes past the last quad start. */
void orc_svb_decode_values(size_t total_items, size_t start_item, size_t end_item,
                           const uint8_t *in, uint32_t *out, int variant, int delta,
                           uint32_t first_value);
/* BlockEncoder.encodeBlock: items = (hash<<32)|id.  Returns items consumed. */
size_t orc_encode_block(const uint64_t *items, size_t n_items, uint32_t min_doc_id, uint8_t *out,
                        size_t block_size);
/* BlockReader: decode a whole block (needs 16 readable bytes after it).  Returns num_items. */
size_t orc_decode_block(const uint8_t *block, size_t block_size, uint32_t min_doc_id,
                        uint32_t *out_hashes, uint32_t *out_docids);
/* BlockReader.searchHash: returns match count, fills [start,end) and docids. */
size_t orc_block_search_hash(const uint8_t *block, size_t block_size, uint32_t min_doc_id,
                             uint32_t hash, uint32_t *out_start, uint32_t *out_end,
                             uint32_t *out_docids);
int orc_uses_ssse3(void);

#ifdef __cplusplus
}
#endif
#endif
