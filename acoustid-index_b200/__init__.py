"""fpx — B200-native `_search` path for acoustid-index (fpindex).

The product is libfpx.so (csrc/: hand-written sm_100a kernels behind the C ABI in include/fpx.h).
This Python package is the host-side mirror of the reference interface used by tests and bench.py.
The directory name has a hyphen (it mirrors the reference repo name); import it through
`__graft_entry__.load_package()` which registers it as `acoustid_index_b200`.
"""
from . import _ffi, index, synth  # noqa: F401  (multi_gpu is imported on demand: it needs torch.distributed)
from ._ffi import FpxError, build as build_library, lib  # noqa: F401
from .index import (WIRE_JSON, WIRE_MSGPACK, Batcher, Context, FileSegment, decode_search_request,
                    encode_search_response, legacy_format_results, legacy_parse_fingerprint, IndexReader, MemorySegment, SearchOptions, SearchRequest,  # noqa: F401
                    SearchResult, Snapshot, SnapshotBuilder, merge_packed_shards_device, merge_shard_results, multi_index_search,
                    open_index_dir, pack_results_device, parse_manifest, segment_file_bytes, segment_file_name,
                    SegmentFile, swap_snapshot, unpack_results)
