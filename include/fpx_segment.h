/*
 * fpx_segment.h — writer for fpindex's block-compressed segment body.
 *
 * Mirrors src/filefmt.zig:94-138 (writeBlocks) + src/block.zig:438-567 (BlockEncoder): greedy packing of
 * (hash, id)-sorted items into fixed-size blocks plus the max-hash block index.  The search path only
 * READS this format (fpx_snapshot_add_file_segment); the writer exists so that benches, tests and tools
 * can produce byte-exact reference segments without the Zig binary.  tests/ check it byte for byte
 * against the oracle's independent restatement.
 */
#ifndef FPX_SEGMENT_H
#define FPX_SEGMENT_H
#include "fpx.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct fpx_segment_buf fpx_segment_buf;

/* items: packed Item (hash<<32)|id, ascending.  block_size 0 = 512 (filefmt.zig:29).
 * threads 0 = hardware concurrency.  The buffer holds num_blocks blocks + the all-zero terminator. */
fpx_status fpx_segment_write(const uint64_t *items, uint64_t n_items, uint32_t min_doc_id,
                             uint32_t block_size, uint32_t threads, fpx_segment_buf **out);
const uint8_t *fpx_segment_buf_blocks(const fpx_segment_buf *b);
const uint32_t *fpx_segment_buf_block_index(const fpx_segment_buf *b);
uint64_t fpx_segment_buf_num_blocks(const fpx_segment_buf *b);
uint64_t fpx_segment_buf_num_items(const fpx_segment_buf *b);
uint32_t fpx_segment_buf_block_size(const fpx_segment_buf *b);
void fpx_segment_buf_free(fpx_segment_buf *b);

/* Decode one block (block.zig:66-312 BlockReader, full decode).  Returns the item count or -1. */
int32_t fpx_block_decode(const uint8_t *block, uint32_t block_size, uint32_t min_doc_id,
                         uint32_t *out_hashes, uint32_t *out_docids);

#ifdef __cplusplus
}
#endif
#endif
