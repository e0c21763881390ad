/* A C99 host of libfpx.so: what a compiled caller of include/fpx.h does, with no Python in between.
 * It writes a small file segment with the library's block writer (fpx_segment_write: the reference's BlockEncoder
 * layout, block.zig:438-567), installs it as a snapshot (the Index.swapSnapshot call site, Index.zig:469-485), and
 * answers one query (the IndexReader.search call site, Index.zig:170-177): the 20 hashes of document 8 must return
 * exactly [{id 8, score 20}] under limit 10, min_score 2, score_pct 10.
 * Exit status: 0 answered correctly, 77 no CUDA device (FPX_BACKEND_UNAVAILABLE: the caller keeps its CPU path),
 * anything else a failure.  tests/test_abi.py compiles it as C (the headers are C, not C++) and runs it. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fpx.h"
#include "fpx_segment.h"

enum { N_DOCS = 1000, HASHES = 20 };

static uint64_t splitmix(uint64_t *x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static int cmp_u64(const void *a, const void *b) {
    const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

#define CHECK(call)                                                                        \
    do {                                                                                   \
        const fpx_status st_ = (call);                                                     \
        if (st_ != FPX_OK) {                                                               \
            fprintf(stderr, "%s -> %d: %s\n", #call, (int)st_, fpx_last_error_message()); \
            return st_ == FPX_BACKEND_UNAVAILABLE ? 77 : 1;                                \
        }                                                                                  \
    } while (0)

int main(void) {
    static uint64_t items[N_DOCS * HASHES]; /* Item = (hash << 32) | id, segment.zig:87-106 */
    static uint32_t doc_ids[N_DOCS], query[HASHES];
    static uint8_t doc_alive[N_DOCS];
    uint64_t rng = 0xF1D0C0DEull;
    for (uint32_t d = 0; d < N_DOCS; ++d) {
        doc_ids[d] = d + 1;
        doc_alive[d] = 1;
        for (uint32_t j = 0; j < HASHES; ++j) {
            const uint32_t h = (uint32_t)splitmix(&rng);
            items[d * HASHES + j] = ((uint64_t)h << 32) | (d + 1);
            if (d + 1 == 8) query[j] = h;
        }
    }
    qsort(items, N_DOCS * HASHES, sizeof items[0], cmp_u64);

    if (fpx_abi_version() != FPX_ABI_VERSION) return 2;
    fpx_segment_buf *buf = NULL;
    CHECK(fpx_segment_write(items, N_DOCS * HASHES, 1, 512, 1, &buf));
    if (fpx_segment_buf_num_items(buf) != N_DOCS * HASHES) return 3;

    fpx_ctx *ctx = NULL;
    CHECK(fpx_init(NULL, &ctx)); /* 77 from here when there is no device */
    fpx_file_segment seg;
    memset(&seg, 0, sizeof seg);
    seg.commit_id = 1;
    seg.min_doc_id = 1;
    seg.block_size = fpx_segment_buf_block_size(buf);
    seg.blocks = fpx_segment_buf_blocks(buf);
    seg.num_blocks = fpx_segment_buf_num_blocks(buf);
    seg.block_index = fpx_segment_buf_block_index(buf);
    seg.doc_ids = doc_ids;
    seg.doc_alive = doc_alive;
    seg.n_docs = N_DOCS;
    fpx_snapshot_builder *b = NULL;
    fpx_snapshot *snap = NULL;
    CHECK(fpx_snapshot_begin(ctx, &b));
    CHECK(fpx_snapshot_add_file_segment(b, &seg));
    CHECK(fpx_snapshot_commit(b, &snap));
    fpx_segment_buf_free(buf); /* the snapshot holds its own copy in HBM */

    const fpx_search_opts opts = {10, 2, 10};
    uint32_t ids[10], scores[10], n = 0;
    CHECK(fpx_search(snap, query, HASHES, &opts, ids, scores, 10, &n));
    printf("results: %u, first (id %u, score %u)\n", n, n ? ids[0] : 0, n ? scores[0] : 0);
    const int ok = n == 1 && ids[0] == 8 && scores[0] == HASHES;
    CHECK(fpx_snapshot_release(snap));
    fpx_shutdown(ctx);
    return ok ? 0 : 4;
}
