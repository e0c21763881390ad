"""ctypes binding of libfpx.so (include/fpx.h, include/fpx_segment.h).

The library is the product; this file is plumbing.  There is deliberately NO fallback: if the shared
library is missing, or no CUDA device is present when a device call is made, the call fails loudly.
"""
import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libfpx.so")

FPX_OK = 0
FPX_OUT_OF_MEMORY = 1
FPX_INVALID_ARGUMENT = 2
FPX_INVALID_SEGMENT = 3
FPX_TIMEOUT = 4
FPX_CUDA_ERROR = 5
FPX_BACKEND_UNAVAILABLE = 6
FPX_UNSUPPORTED = 7
STATUS_NAMES = {
    0: "OK", 1: "OUT_OF_MEMORY", 2: "INVALID_ARGUMENT", 3: "INVALID_SEGMENT", 4: "TIMEOUT",
    5: "CUDA_ERROR", 6: "BACKEND_UNAVAILABLE", 7: "UNSUPPORTED",
}
FPX_FLAG_PROFILE = 1
FPX_FLAG_HOST_ONLY = 2
FPX_FLAG_NO_SKETCH = 4
FPX_FLAG_HOST_BUILD = 8
FPX_MAX_QUERY_TERMS = 8192
FPX_MAX_RESULTS = 1024

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


class FpxError(RuntimeError):
    def __init__(self, status, message):
        self.status = status
        super().__init__("fpx: %s: %s" % (STATUS_NAMES.get(status, status), message))


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("host_threads", C.c_uint32), ("chunk_queries", C.c_uint32),
                ("flags", C.c_uint32)]


class FileSegmentDesc(C.Structure):
    _fields_ = [("commit_id", C.c_uint64), ("merges", C.c_uint64), ("min_doc_id", C.c_uint32),
                ("block_size", C.c_uint32), ("blocks", C.c_void_p), ("num_blocks", C.c_uint64),
                ("block_index", C.c_void_p), ("doc_ids", C.c_void_p), ("doc_alive", C.c_void_p),
                ("n_docs", C.c_uint64)]


class MemorySegmentDesc(C.Structure):
    _fields_ = [("commit_id", C.c_uint64), ("merges", C.c_uint64), ("items", C.c_void_p),
                ("n_items", C.c_uint64), ("doc_ids", C.c_void_p), ("doc_alive", C.c_void_p),
                ("n_docs", C.c_uint64)]


class WireSearchRequest(C.Structure):
    _fields_ = [("query", u32p), ("n_terms", C.c_uint64), ("timeout", C.c_uint32), ("limit", C.c_uint32),
                ("has_min_score", C.c_uint32), ("min_score", C.c_uint32), ("score_pct", C.c_uint32)]


class BatcherConfig(C.Structure):
    _fields_ = [("max_batch", C.c_uint32), ("max_wait_us", C.c_uint32)]


class BatcherStats(C.Structure):
    _fields_ = [("batches", C.c_uint64), ("queries", C.c_uint64), ("max_batch_seen", C.c_uint64),
                ("timeouts", C.c_uint64)]


class SegmentInfo(C.Structure):  # fpx_segment_info
    _fields_ = [("commit_id", C.c_uint64), ("merges", C.c_uint64), ("version", C.c_uint64),
                ("has_version", C.c_uint32), ("reserved", C.c_uint32)]


class SnapshotInfo(C.Structure):
    _fields_ = [("n_segments", C.c_uint64), ("n_terms", C.c_uint64), ("n_postings", C.c_uint64),
                ("n_postings_total", C.c_uint64), ("n_dropped_unreachable", C.c_uint64),
                ("n_dropped_superseded", C.c_uint64), ("n_dropped_out_of_range", C.c_uint64),
                ("device_bytes", C.c_uint64), ("max_row_len", C.c_uint64), ("pad_id", C.c_uint32),
                ("table_log2", C.c_uint32), ("doc_lo", C.c_uint32), ("doc_hi", C.c_uint32)]


class CsrView(C.Structure):
    _fields_ = [("n_terms", C.c_uint64), ("terms", u32p), ("row_offsets", u64p), ("docids", u32p)]


class Profile(C.Structure):
    _fields_ = [("prepare_ms", C.c_double), ("prepare_launches", C.c_uint64),
                ("sketch_ms", C.c_double), ("sketch_launches", C.c_uint64),
                ("search_ms", C.c_double), ("search_launches", C.c_uint64),
                ("wide_ms", C.c_double), ("wide_launches", C.c_uint64),
                ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
                ("queries", C.c_uint64), ("unique_terms", C.c_uint64), ("postings", C.c_uint64),
                ("results", C.c_uint64), ("sketch_queries", C.c_uint64), ("wide_queries", C.c_uint64),
                ("overflow_requeues", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


# every symbol the headers declare; tests assert the library exports all of them
EXPORTS = [
    "fpx_abi_version", "fpx_last_error_message", "fpx_init", "fpx_shutdown",
    "fpx_snapshot_begin", "fpx_snapshot_add_file_segment", "fpx_snapshot_add_memory_segment",
    "fpx_snapshot_set_doc_range", "fpx_snapshot_compile", "fpx_snapshot_csr", "fpx_snapshot_commit",
    "fpx_snapshot_abort", "fpx_snapshot_acquire", "fpx_snapshot_release", "fpx_snapshot_get_info",
    "fpx_snapshot_row_lengths", "fpx_snapshot_read_row", "fpx_default_min_score", "fpx_search", "fpx_search_batch",
    "fpx_search_batch_packed", "fpx_search_batch_timeout", "fpx_search_batch_device_async",
    "fpx_search_batch_device", "fpx_merge_shard_results", "fpx_profile_reset", "fpx_profile_read", "fpx_debug_set", "fpx_set_chunk_queries", "fpx_set_profile", "fpx_pack_results_device", "fpx_merge_packed_shards_device",
    "fpx_segment_write", "fpx_segment_buf_blocks", "fpx_segment_buf_block_index",
    "fpx_segment_buf_num_blocks", "fpx_segment_buf_num_items", "fpx_segment_buf_block_size",
    "fpx_segment_buf_free", "fpx_block_decode",
    "fpx_wire_decode_search_request", "fpx_wire_encode_search_response", "fpx_legacy_parse_fingerprint",
    "fpx_legacy_format_results", "fpx_wire_free",
    "fpx_batcher_create", "fpx_batcher_set_snapshot", "fpx_batcher_search", "fpx_batcher_get_stats",
    "fpx_batcher_destroy",
    "fpx_segment_file_parse", "fpx_segment_file_read", "fpx_segment_file_view", "fpx_segment_file_num_items",
    "fpx_segment_file_metadata_count", "fpx_segment_file_metadata_get", "fpx_segment_file_close",
    "fpx_segment_file_serialize", "fpx_bytes_free", "fpx_segment_file_name", "fpx_manifest_parse", "fpx_crc64_xz",
]

_LIB = None


def build(verbose=False):
    """Compile libfpx.so in-tree with nvcc for sm_100a (works without a GPU)."""
    out = subprocess.run(["make", "-C", _PKG, "-B"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("libfpx.so build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise FpxError(FPX_BACKEND_UNAVAILABLE,
                       "libfpx.so is not built (run __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.fpx_abi_version.restype = C.c_uint32
    L.fpx_last_error_message.restype = C.c_char_p
    L.fpx_init.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.fpx_shutdown.argtypes = [vp]
    L.fpx_shutdown.restype = None
    L.fpx_snapshot_begin.argtypes = [vp, C.POINTER(vp)]
    L.fpx_snapshot_add_file_segment.argtypes = [vp, C.POINTER(FileSegmentDesc)]
    L.fpx_snapshot_add_memory_segment.argtypes = [vp, C.POINTER(MemorySegmentDesc)]
    L.fpx_snapshot_set_doc_range.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.fpx_snapshot_compile.argtypes = [vp]
    L.fpx_snapshot_csr.argtypes = [vp, C.POINTER(CsrView)]
    L.fpx_snapshot_commit.argtypes = [vp, C.POINTER(vp)]
    L.fpx_snapshot_abort.argtypes = [vp]
    L.fpx_snapshot_abort.restype = None
    L.fpx_snapshot_acquire.argtypes = [vp]
    L.fpx_snapshot_release.argtypes = [vp]
    L.fpx_snapshot_get_info.argtypes = [vp, C.POINTER(SnapshotInfo)]
    L.fpx_snapshot_row_lengths.argtypes = [vp, vp, C.c_uint64, vp]
    L.fpx_snapshot_read_row.argtypes = [vp, C.c_uint32, vp, C.c_uint64, u64p]
    L.fpx_default_min_score.argtypes = [C.c_uint64]
    L.fpx_default_min_score.restype = C.c_uint32
    L.fpx_search.argtypes = [vp, vp, C.c_uint64, vp, vp, vp, C.c_uint32, u32p]
    L.fpx_search_batch.argtypes = [vp, C.c_uint64, vp, vp, vp, C.c_uint32, vp, vp, vp]
    L.fpx_search_batch_device.argtypes = [vp, C.c_uint64, vp, vp, vp, C.c_uint32, vp, vp, vp, vp]
    L.fpx_merge_shard_results.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, vp, vp, vp, vp, vp, vp, vp]
    L.fpx_profile_reset.argtypes = [vp]
    L.fpx_profile_read.argtypes = [vp, C.POINTER(Profile)]
    L.fpx_debug_set.argtypes = [vp, C.c_uint32]
    L.fpx_set_chunk_queries.argtypes = [vp, C.c_uint32]
    L.fpx_set_profile.argtypes = [vp, C.c_int]
    L.fpx_search_batch_timeout.argtypes = [vp, C.c_uint64, vp, vp, vp, C.c_uint32, vp, vp, vp, C.c_uint32]
    L.fpx_search_batch_device_async.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, vp, vp, vp, C.c_uint32, vp, vp, vp, vp, vp]
    L.fpx_search_batch_packed.argtypes = [vp, C.c_uint64, vp, vp, vp, C.c_uint32, vp, vp, C.c_uint64, vp]
    L.fpx_pack_results_device.argtypes = [C.c_uint64, C.c_uint32, vp, vp, vp, vp, C.c_uint32, vp]
    L.fpx_merge_packed_shards_device.argtypes = [C.c_uint32, C.c_uint64, vp, C.c_uint64, vp, C.c_uint32, vp, vp, vp, vp]
    L.fpx_wire_decode_search_request.argtypes = [C.c_uint32, C.c_char_p, C.c_uint64, C.POINTER(WireSearchRequest)]
    L.fpx_wire_encode_search_response.argtypes = [C.c_uint32, vp, vp, C.c_uint32, C.POINTER(vp), u64p]
    L.fpx_legacy_parse_fingerprint.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(u32p), u64p]
    L.fpx_legacy_format_results.argtypes = [vp, vp, C.c_uint32, C.POINTER(vp), u64p]
    L.fpx_wire_free.argtypes = [vp]
    L.fpx_wire_free.restype = None
    L.fpx_batcher_create.argtypes = [vp, C.POINTER(BatcherConfig), C.POINTER(vp)]
    L.fpx_batcher_set_snapshot.argtypes = [vp, vp]
    L.fpx_batcher_search.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint32, vp, vp, C.c_uint32, u32p]
    L.fpx_batcher_get_stats.argtypes = [vp, C.POINTER(BatcherStats)]
    L.fpx_batcher_destroy.argtypes = [vp]
    L.fpx_batcher_destroy.restype = None
    L.fpx_segment_file_parse.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
    L.fpx_segment_file_read.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.fpx_segment_file_view.argtypes = [vp, C.POINTER(FileSegmentDesc), C.POINTER(SegmentInfo)]
    L.fpx_segment_file_num_items.argtypes = [vp]
    L.fpx_segment_file_num_items.restype = C.c_uint64
    L.fpx_segment_file_metadata_count.argtypes = [vp]
    L.fpx_segment_file_metadata_count.restype = C.c_uint64
    L.fpx_segment_file_metadata_get.argtypes = [vp, C.c_uint64, C.POINTER(vp), u64p, C.POINTER(vp), u64p]
    L.fpx_segment_file_close.argtypes = [vp]
    L.fpx_segment_file_close.restype = None
    L.fpx_segment_file_serialize.argtypes = [C.POINTER(FileSegmentDesc), C.POINTER(SegmentInfo), C.POINTER(vp), u64p]
    L.fpx_bytes_free.argtypes = [vp]
    L.fpx_bytes_free.restype = None
    L.fpx_segment_file_name.argtypes = [C.c_uint64, C.c_uint64, C.c_char_p, C.c_uint64]
    L.fpx_segment_file_name.restype = C.c_int32
    L.fpx_manifest_parse.argtypes = [vp, C.c_uint64, C.POINTER(SegmentInfo), C.c_uint64, u64p]
    L.fpx_crc64_xz.argtypes = [vp, C.c_uint64]
    L.fpx_crc64_xz.restype = C.c_uint64
    L.fpx_segment_write.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp)]
    L.fpx_segment_buf_blocks.argtypes = [vp]
    L.fpx_segment_buf_blocks.restype = vp
    L.fpx_segment_buf_block_index.argtypes = [vp]
    L.fpx_segment_buf_block_index.restype = vp
    L.fpx_segment_buf_num_blocks.argtypes = [vp]
    L.fpx_segment_buf_num_blocks.restype = C.c_uint64
    L.fpx_segment_buf_num_items.argtypes = [vp]
    L.fpx_segment_buf_num_items.restype = C.c_uint64
    L.fpx_segment_buf_block_size.argtypes = [vp]
    L.fpx_segment_buf_block_size.restype = C.c_uint32
    L.fpx_segment_buf_free.argtypes = [vp]
    L.fpx_segment_buf_free.restype = None
    L.fpx_block_decode.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp]
    L.fpx_block_decode.restype = C.c_int32
    _LIB = L
    return L


def check(status):
    if status != FPX_OK:
        msg = lib().fpx_last_error_message()
        raise FpxError(status, msg.decode("utf-8", "replace") if msg else "")
