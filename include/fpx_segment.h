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

/* ---- segment files (.data) and the manifest: src/filefmt.zig:1-13, 209-285; src/manifest.zig:17-39 ----
 * A real fpindex keeps its file segments as <dir>/<name>/v<gen>/data/<commit_id:016x>-<merges:08x>.data plus a
 * `manifest` (README.md:107-109).  These calls load them straight into the layout fpx_snapshot_add_file_segment
 * takes.  Header / footer are msgpack maps keyed by field index; any valid msgpack encoding is accepted (the
 * reference's encoder is not vendored).  Errors: FPX_INVALID_SEGMENT (error.InvalidSegment and
 * error.ChecksumMismatch of filefmt.zig:235-284; fpx_last_error_message tells which). */
typedef struct fpx_segment_file fpx_segment_file;
typedef struct fpx_segment_info { /* segment.zig:23-26 */
    uint64_t commit_id, merges, version;
    uint32_t has_version, reserved;
} fpx_segment_info;

/* Parse a segment file held in memory (the bytes are copied) / read from `path`. */
fpx_status fpx_segment_file_parse(const uint8_t *data, uint64_t size, fpx_segment_file **out);
fpx_status fpx_segment_file_read(const char *path, fpx_segment_file **out);
/* The segment as fpx_snapshot_add_file_segment takes it; pointers stay valid until fpx_segment_file_close. */
fpx_status fpx_segment_file_view(const fpx_segment_file *f, fpx_file_segment *out, fpx_segment_info *info);
uint64_t fpx_segment_file_num_items(const fpx_segment_file *f);
uint64_t fpx_segment_file_metadata_count(const fpx_segment_file *f);
fpx_status fpx_segment_file_metadata_get(const fpx_segment_file *f, uint64_t i, const char **key, uint64_t *key_len,
                                         const char **value, uint64_t *value_len);
void fpx_segment_file_close(fpx_segment_file *f);
/* filefmt.zig:143-178: the bytes writeSegment would put on disk for this segment (empty metadata).
 * Free with fpx_bytes_free. */
fpx_status fpx_segment_file_serialize(const fpx_file_segment *seg, const fpx_segment_info *info, uint8_t **out,
                                      uint64_t *out_size);
void fpx_bytes_free(uint8_t *p);
/* filefmt.zig:36, 45-48: "<commit_id:016x>-<merges:08x>.data"; returns the length or -1 if buf is too small. */
int32_t fpx_segment_file_name(uint64_t commit_id, uint64_t merges, char *buf, uint64_t cap);
/* manifest.zig:17-39: msgpack array of SegmentInfo.  *n receives the number of segments (also when cap is too
 * small: FPX_INVALID_ARGUMENT then). */
fpx_status fpx_manifest_parse(const uint8_t *data, uint64_t size, fpx_segment_info *out, uint64_t cap, uint64_t *n);
/* CRC-64/XZ as the footer holds it (std.hash.crc.Crc64Xz). */
uint64_t fpx_crc64_xz(const uint8_t *data, uint64_t size);

#ifdef __cplusplus
}
#endif
#endif
