"""Host-side mirror of the reference's interface for the `_search` path, over the C ABI.

Names and argument meaning follow the reference (paths relative to the reference root):
  SearchOptions / SearchResult      src/common.zig:45-54
  SearchRequest                     src/api.zig:14-27
  FileSegment / MemorySegment       src/FileSegment.zig:33-53, src/MemorySegment.zig:21-28 (data only)
  Index.swap_snapshot               src/Index.zig:469-485   -> builds an immutable GPU snapshot
  IndexReader.search                src/Index.zig:170-177 + src/common.zig:131-167 (finish)
  multi_index_search                src/MultiIndex.zig:287-330 option mapping (+ server.zig:192 clamp)

Everything here is plumbing around libfpx.so; no search arithmetic happens in Python.
"""
import ctypes as C
from typing import List, NamedTuple, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import FpxError, check, lib

# src/api.zig:7-11
default_search_timeout = 500
max_search_timeout = 10000
default_search_limit = 40
min_search_limit = 1
max_search_limit = 100


class SearchResult(NamedTuple):
    id: int
    score: int


class SearchOptions(NamedTuple):  # common.zig:50-54
    max_results: int = 10
    min_score: int = 1
    min_score_pct: int = 10


class SearchRequest:  # api.zig:14-27
    def __init__(self, query, timeout=default_search_timeout, limit=default_search_limit, min_score=None,
                 score_pct=10):
        self.query = list(query)
        self.timeout = timeout
        self.limit = limit
        self.min_score = min_score
        self.score_pct = score_pct


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


class FileSegment:
    """An immutable block-compressed segment resident in host memory (FileSegment.zig:33-53)."""

    def __init__(self, commit_id, merges, min_doc_id, block_size, blocks, num_blocks, block_index, doc_ids, doc_alive,
                 _owner=None):
        self.commit_id, self.merges = int(commit_id), int(merges)
        self.min_doc_id, self.block_size = int(min_doc_id), int(block_size)
        self.blocks, self.num_blocks, self.block_index = blocks, int(num_blocks), block_index
        self.doc_ids, self.doc_alive = _u32(doc_ids), np.ascontiguousarray(doc_alive, dtype=np.uint8)
        self._owner = _owner

    @staticmethod
    def from_items(items, doc_ids, doc_alive, commit_id, merges=0, block_size=512, threads=0):
        """Write sorted packed items ((hash<<32)|id) as blocks: filefmt.zig:94-138 via fpx_segment_write."""
        items = np.ascontiguousarray(items, dtype=np.uint64)
        doc_ids = _u32(doc_ids)
        min_doc_id = int(doc_ids.min()) if len(doc_ids) else 0  # filefmt.zig:244-250
        buf = C.c_void_p()
        check(lib().fpx_segment_write(items.ctypes.data, len(items), min_doc_id, block_size, threads, C.byref(buf)))
        owner = _SegmentBuf(buf)
        nb = lib().fpx_segment_buf_num_blocks(buf)
        bs = lib().fpx_segment_buf_block_size(buf)
        blocks = np.ctypeslib.as_array(C.cast(lib().fpx_segment_buf_blocks(buf), _ffi.u8p), shape=((nb + 1) * bs,))
        if nb:
            index = np.ctypeslib.as_array(C.cast(lib().fpx_segment_buf_block_index(buf), _ffi.u32p), shape=(nb,))
        else:
            index = np.zeros(0, dtype=np.uint32)
        assert lib().fpx_segment_buf_num_items(buf) == len(items)
        return FileSegment(commit_id, merges, min_doc_id, bs, blocks, nb, index, doc_ids, doc_alive, _owner=owner)

    def _desc(self):
        d = _ffi.FileSegmentDesc()
        d.commit_id, d.merges = self.commit_id, self.merges
        d.min_doc_id, d.block_size = self.min_doc_id, self.block_size
        d.blocks = self.blocks.ctypes.data if self.num_blocks else None
        d.num_blocks = self.num_blocks
        d.block_index = self.block_index.ctypes.data if self.num_blocks else None
        d.doc_ids = self.doc_ids.ctypes.data if len(self.doc_ids) else None
        d.doc_alive = self.doc_alive.ctypes.data if len(self.doc_alive) else None
        d.n_docs = len(self.doc_ids)
        return d


class SegmentFile:
    """A segment file (.data) loaded into memory: src/filefmt.zig:209-285 readSegment.  `.segment` is the FileSegment
    view (valid while this object lives), `.info` = (commit_id, merges, version or None)."""

    def __init__(self, h):
        self.h = h
        d, si = _ffi.FileSegmentDesc(), _ffi.SegmentInfo()
        check(lib().fpx_segment_file_view(h, C.byref(d), C.byref(si)))
        self.info = (si.commit_id, si.merges, si.version if si.has_version else None)
        self.num_items = lib().fpx_segment_file_num_items(h)
        nb, bs, nd = int(d.num_blocks), int(d.block_size), int(d.n_docs)
        blocks = np.ctypeslib.as_array(C.cast(d.blocks, _ffi.u8p), shape=((nb + 1) * bs,))
        index = np.ctypeslib.as_array(C.cast(d.block_index, _ffi.u32p), shape=(nb,)) if nb else np.zeros(0, np.uint32)
        ids = np.ctypeslib.as_array(C.cast(d.doc_ids, _ffi.u32p), shape=(nd,)) if nd else np.zeros(0, np.uint32)
        alive = np.ctypeslib.as_array(C.cast(d.doc_alive, _ffi.u8p), shape=(nd,)) if nd else np.zeros(0, np.uint8)
        self.segment = FileSegment(d.commit_id, d.merges, d.min_doc_id, bs, blocks, nb, index, ids.copy(), alive.copy(),
                                   _owner=self)
        self.metadata = {}
        for i in range(lib().fpx_segment_file_metadata_count(h)):
            k, v, kl, vl = C.c_void_p(), C.c_void_p(), C.c_uint64(), C.c_uint64()
            check(lib().fpx_segment_file_metadata_get(h, i, C.byref(k), C.byref(kl), C.byref(v), C.byref(vl)))
            self.metadata[C.string_at(k, kl.value)] = C.string_at(v, vl.value)

    @staticmethod
    def parse(data: bytes):
        h = C.c_void_p()
        buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data or b"\0")
        check(lib().fpx_segment_file_parse(C.cast(buf, C.c_void_p), len(data), C.byref(h)))
        return SegmentFile(h)

    @staticmethod
    def read(path):
        h = C.c_void_p()
        check(lib().fpx_segment_file_read(str(path).encode(), C.byref(h)))
        return SegmentFile(h)

    def __del__(self):
        try:
            if self.h:
                lib().fpx_segment_file_close(self.h)
                self.h = None
        except Exception:
            pass


def segment_file_bytes(seg: "FileSegment", version=None) -> bytes:
    """The bytes filefmt.writeSegment (filefmt.zig:143-178) puts on disk for `seg` (empty metadata)."""
    d = seg._desc()
    si = _ffi.SegmentInfo(seg.commit_id, seg.merges, version or 0, 1 if version is not None else 0, 0)
    out, n = C.c_void_p(), C.c_uint64()
    check(lib().fpx_segment_file_serialize(C.byref(d), C.byref(si), C.byref(out), C.byref(n)))
    try:
        return C.string_at(out, n.value)
    finally:
        lib().fpx_bytes_free(out)


def segment_file_name(commit_id, merges=0) -> str:
    buf = C.create_string_buffer(64)
    n = lib().fpx_segment_file_name(commit_id, merges, buf, 64)
    assert n > 0
    return buf.value.decode()


def parse_manifest(data: bytes):
    """manifest.zig:17-39 -> list of (commit_id, merges, version or None)."""
    arr = (_ffi.SegmentInfo * 4096)()
    n = C.c_uint64()
    buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data or b"\0")
    check(lib().fpx_manifest_parse(C.cast(buf, C.c_void_p), len(data), arr, 4096, C.byref(n)))
    return [(arr[i].commit_id, arr[i].merges, arr[i].version if arr[i].has_version else None) for i in range(n.value)]


def open_index_dir(path):
    """Load every file segment the manifest in `path` lists (Index.open, Index.zig:271-315): SegmentFile objects,
    oldest first."""
    import os
    mp = os.path.join(path, "manifest")
    infos = parse_manifest(open(mp, "rb").read()) if os.path.exists(mp) else []
    return [SegmentFile.read(os.path.join(path, segment_file_name(c, m))) for c, m, _ in infos]


class _SegmentBuf:
    def __init__(self, h):
        self.h = h

    def __del__(self):
        try:
            if self.h:
                lib().fpx_segment_buf_free(self.h)
                self.h = None
        except Exception:
            pass


class MemorySegment:
    """Sorted packed items + docs map of an in-memory segment (MemorySegment.zig:21-28)."""

    def __init__(self, commit_id, merges, items, doc_ids, doc_alive):
        self.commit_id, self.merges = int(commit_id), int(merges)
        self.items = np.ascontiguousarray(items, dtype=np.uint64)
        self.doc_ids, self.doc_alive = _u32(doc_ids), np.ascontiguousarray(doc_alive, dtype=np.uint8)

    def _desc(self):
        d = _ffi.MemorySegmentDesc()
        d.commit_id, d.merges = self.commit_id, self.merges
        d.items = self.items.ctypes.data if len(self.items) else None
        d.n_items = len(self.items)
        d.doc_ids = self.doc_ids.ctypes.data if len(self.doc_ids) else None
        d.doc_alive = self.doc_alive.ctypes.data if len(self.doc_alive) else None
        d.n_docs = len(self.doc_ids)
        return d


class Context:
    def __init__(self, device=-1, profile=False, host_only=False, host_threads=0, chunk_queries=0, no_sketch=False,
                 host_build=False):
        cfg = _ffi.Config(device, host_threads, chunk_queries,
                          (_ffi.FPX_FLAG_PROFILE if profile else 0) | (_ffi.FPX_FLAG_HOST_ONLY if host_only else 0) |
                          (_ffi.FPX_FLAG_NO_SKETCH if no_sketch else 0) | (_ffi.FPX_FLAG_HOST_BUILD if host_build else 0))
        self.h = C.c_void_p()
        check(lib().fpx_init(C.byref(cfg), C.byref(self.h)))

    def close(self):
        if self.h:
            lib().fpx_shutdown(self.h)
            self.h = None

    def set_profile(self, enabled):
        check(lib().fpx_set_profile(self.h, 1 if enabled else 0))

    def set_chunk_queries(self, n):
        check(lib().fpx_set_chunk_queries(self.h, int(n)))

    def debug_set(self, bits):
        """Profiling only: kernel variant / ablation bits (results are wrong while ablation bits are set)."""
        check(lib().fpx_debug_set(self.h, int(bits)))

    def profile_reset(self):
        check(lib().fpx_profile_reset(self.h))

    def profile(self):
        p = _ffi.Profile()
        check(lib().fpx_profile_read(self.h, C.byref(p)))
        return {k: getattr(p, k) for k, _ in _ffi.Profile._fields_}


class SnapshotBuilder:
    """Index.swapSnapshot hook: segments oldest -> newest, file before memory (Index.zig:33-41)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.h = C.c_void_p()
        check(lib().fpx_snapshot_begin(ctx.h, C.byref(self.h)))

    def add_file_segment(self, seg: FileSegment):
        d = seg._desc()
        check(lib().fpx_snapshot_add_file_segment(self.h, C.byref(d)))

    def add_memory_segment(self, seg: MemorySegment):
        d = seg._desc()
        check(lib().fpx_snapshot_add_memory_segment(self.h, C.byref(d)))

    def set_doc_range(self, lo, hi):
        check(lib().fpx_snapshot_set_doc_range(self.h, lo, hi))

    def csr(self):
        """Host view of the compiled CSR as numpy copies: (terms, row_offsets, docids)."""
        v = _ffi.CsrView()
        check(lib().fpx_snapshot_csr(self.h, C.byref(v)))
        n = v.n_terms
        if n == 0:
            return np.zeros(0, np.uint32), np.zeros(1, np.uint64), np.zeros(0, np.uint32)
        terms = np.ctypeslib.as_array(v.terms, shape=(n,)).copy()
        offs = np.ctypeslib.as_array(v.row_offsets, shape=(n + 1,)).copy()
        total = int(offs[-1])
        docids = np.ctypeslib.as_array(v.docids, shape=(total,)).copy() if total else np.zeros(0, np.uint32)
        return terms, offs, docids

    def commit(self):
        s = C.c_void_p()
        check(lib().fpx_snapshot_commit(self.h, C.byref(s)))
        self.h = None
        return Snapshot(self.ctx, s)

    def abort(self):
        if self.h:
            lib().fpx_snapshot_abort(self.h)
            self.h = None

    def __del__(self):
        try:
            self.abort()
        except Exception:
            pass


class Snapshot:
    """Immutable, refcounted GPU mirror of a Segments snapshot (Index.zig:36-63)."""

    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    def info(self):
        i = _ffi.SnapshotInfo()
        check(lib().fpx_snapshot_get_info(self.h, C.byref(i)))
        return {k: getattr(i, k) for k, _ in _ffi.SnapshotInfo._fields_}

    def row_lengths(self, terms):
        terms = _u32(terms)
        out = np.zeros(len(terms), dtype=np.uint32)
        if len(terms):
            check(lib().fpx_snapshot_row_lengths(self.h, terms.ctypes.data, len(terms), out.ctypes.data))
        return out

    def read_row(self, term, capacity=1 << 20):
        """The row of `term` as it lies in HBM (debug / tests)."""
        out = np.zeros(capacity, dtype=np.uint32)
        n = C.c_uint64(0)
        check(lib().fpx_snapshot_read_row(self.h, int(term) & 0xFFFFFFFF, out.ctypes.data, capacity, C.byref(n)))
        return out[:n.value].copy()

    def acquire(self):
        check(lib().fpx_snapshot_acquire(self.h))

    def release(self):
        if self.h:
            check(lib().fpx_snapshot_release(self.h))
            self.h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def swap_snapshot(ctx: Context, file_segments: Sequence[FileSegment], memory_segments: Sequence[MemorySegment] = (),
                  doc_range=None) -> Snapshot:
    """What a host would do inside Index.swapSnapshot (Index.zig:469-485): mirror the new Segments in HBM."""
    b = SnapshotBuilder(ctx)
    for s in file_segments:
        b.add_file_segment(s)
    for s in memory_segments:
        b.add_memory_segment(s)
    if doc_range is not None:
        b.set_doc_range(*doc_range)
    return b.commit()


class IndexReader:
    """A held snapshot; search works on it without any lock (Index.zig:152-177)."""

    def __init__(self, snapshot: Snapshot):
        self.snapshot = snapshot

    def search(self, hashes, options: SearchOptions = SearchOptions()) -> List[SearchResult]:
        q = _u32([int(x) & 0xFFFFFFFF for x in hashes])
        opts = np.array([options.max_results, options.min_score, options.min_score_pct], dtype=np.uint32)
        cap = max(1, min(int(options.max_results), _ffi.FPX_MAX_RESULTS))
        ids = np.zeros(cap, dtype=np.uint32)
        sc = np.zeros(cap, dtype=np.uint32)
        n = C.c_uint32(0)
        check(lib().fpx_search(self.snapshot.h, q.ctypes.data if len(q) else None, len(q), opts.ctypes.data,
                               ids.ctypes.data, sc.ctypes.data, cap, C.byref(n)))
        return [SearchResult(int(ids[i]), int(sc[i])) for i in range(n.value)]

    def search_batch(self, terms, term_offsets, opts, k_stride, out=None):
        """Host buffers (numpy; pinned torch memory can be passed via .numpy()).  Returns (ids, scores, counts)."""
        terms = _u32(terms)
        term_offsets = np.ascontiguousarray(term_offsets, dtype=np.uint64)
        opts = _u32(opts)
        nq = len(term_offsets) - 1
        assert opts.size == 3 * nq
        if out is None:
            out = (np.zeros((nq, k_stride), np.uint32), np.zeros((nq, k_stride), np.uint32), np.zeros(nq, np.uint32))
        ids, sc, cnt = out
        check(lib().fpx_search_batch(self.snapshot.h, nq, terms.ctypes.data if len(terms) else None,
                                     term_offsets.ctypes.data, opts.ctypes.data, k_stride, ids.ctypes.data,
                                     sc.ctypes.data, cnt.ctypes.data))
        return ids, sc, cnt

    def search_batch_ptr(self, nq, terms_ptr, offsets_ptr, opts_ptr, k_stride, ids_ptr, scores_ptr, counts_ptr):
        """Raw host pointers (e.g. pinned torch tensors' data_ptr())."""
        check(lib().fpx_search_batch(self.snapshot.h, nq, terms_ptr, offsets_ptr, opts_ptr, k_stride, ids_ptr,
                                     scores_ptr, counts_ptr))

    def search_batch_packed_ptr(self, nq, terms_ptr, offsets_ptr, opts_ptr, k_stride, counts_ptr, pairs_ptr,
                                capacity_pairs):
        """Raw host pointers; results as {count per query, (id, score) pairs back to back}.  Returns the number of pairs."""
        n = C.c_uint64(0)
        check(lib().fpx_search_batch_packed(self.snapshot.h, nq, terms_ptr, offsets_ptr, opts_ptr, k_stride, counts_ptr,
                                            pairs_ptr, capacity_pairs, C.byref(n)))
        return n.value

    def search_batch_packed(self, terms, offsets, opts, k_stride, capacity_pairs=None):
        """numpy in, (counts [nq], pairs [n, 2]) out: the reference's list-per-query result layout."""
        terms, offsets = _u32(terms), np.ascontiguousarray(offsets, dtype=np.uint64)
        opts = np.ascontiguousarray(opts, dtype=np.uint32).reshape(-1, 3)
        nq = len(offsets) - 1
        cap = int(capacity_pairs if capacity_pairs is not None else nq * k_stride)
        counts = np.zeros(nq, np.uint32)
        pairs = np.zeros((max(cap, 1), 2), np.uint32)
        n = self.search_batch_packed_ptr(nq, terms.ctypes.data if len(terms) else None, offsets.ctypes.data, opts.ctypes.data,
                                         k_stride, counts.ctypes.data, pairs.ctypes.data, cap)
        return counts, pairs[:n]

    def search_batch_device_async(self, nq, term_base, n_terms_total, d_terms, d_offsets, d_opts, k_stride, d_ids, d_scores,
                                  d_counts, d_status=None, stream=0):
        """Device pointers (ints); never waits for the device (see fpx.h).  d_status: device word for what the kernels raise."""
        check(lib().fpx_search_batch_device_async(self.snapshot.h, nq, term_base, n_terms_total, d_terms, d_offsets, d_opts,
                                                  k_stride, d_ids, d_scores, d_counts, d_status, stream))

    def search_batch_timeout(self, terms, offsets, opts, k_stride, timeout_ms):
        """search_batch with a deadline: raises FpxError(FPX_TIMEOUT) like error.SearchTimeout."""
        terms, offsets = _u32(terms), np.ascontiguousarray(offsets, dtype=np.uint64)
        opts = np.ascontiguousarray(opts, dtype=np.uint32).reshape(-1, 3)
        nq = len(offsets) - 1
        ids, sc, cnt = np.zeros((nq, k_stride), np.uint32), np.zeros((nq, k_stride), np.uint32), np.zeros(nq, np.uint32)
        check(lib().fpx_search_batch_timeout(self.snapshot.h, nq, terms.ctypes.data if len(terms) else None, offsets.ctypes.data,
                                             opts.ctypes.data, k_stride, ids.ctypes.data, sc.ctypes.data, cnt.ctypes.data,
                                             int(timeout_ms)))
        return ids, sc, cnt

    def search_batch_device(self, nq, d_terms, d_offsets, d_opts, k_stride, d_ids, d_scores, d_counts, stream=0):
        """Device pointers (ints), asynchronous on `stream` (a cudaStream_t handle as int)."""
        check(lib().fpx_search_batch_device(self.snapshot.h, nq, d_terms, d_offsets, d_opts, k_stride, d_ids, d_scores,
                                            d_counts, stream))


WIRE_JSON, WIRE_MSGPACK = 0, 1


def decode_search_request(data: bytes, fmt=WIRE_JSON) -> SearchRequest:
    """server.handleSearch's decode + sanitize (server.zig:189-194) -> SearchRequest."""
    r = _ffi.WireSearchRequest()
    check(lib().fpx_wire_decode_search_request(fmt, data, len(data), C.byref(r)))
    try:
        q = [int(r.query[i]) for i in range(r.n_terms)]
    finally:
        lib().fpx_wire_free(C.cast(r.query, C.c_void_p))
    return SearchRequest(q, timeout=r.timeout, limit=r.limit, min_score=r.min_score if r.has_min_score else None,
                         score_pct=r.score_pct)


def encode_search_response(results, fmt=WIRE_JSON) -> bytes:
    ids = _u32([r[0] for r in results])
    sc = _u32([r[1] for r in results])
    out, n = C.c_void_p(), C.c_uint64()
    check(lib().fpx_wire_encode_search_response(fmt, ids.ctypes.data if len(ids) else None,
                                                sc.ctypes.data if len(sc) else None, len(ids), C.byref(out), C.byref(n)))
    try:
        return C.string_at(out, n.value)
    finally:
        lib().fpx_wire_free(out)


def legacy_parse_fingerprint(text: str):
    """legacy.zig:286-296: comma-separated signed decimals -> u32 hashes."""
    b = text.encode()
    out, n = _ffi.u32p(), C.c_uint64()
    check(lib().fpx_legacy_parse_fingerprint(b, len(b), C.byref(out), C.byref(n)))
    try:
        return [int(out[i]) for i in range(n.value)]
    finally:
        lib().fpx_wire_free(C.cast(out, C.c_void_p))


def legacy_format_results(results) -> str:
    ids = _u32([r[0] for r in results])
    sc = _u32([r[1] for r in results])
    out, n = C.c_void_p(), C.c_uint64()
    check(lib().fpx_legacy_format_results(ids.ctypes.data if len(ids) else None, sc.ctypes.data if len(sc) else None,
                                          len(ids), C.byref(out), C.byref(n)))
    try:
        return C.string_at(out, n.value).decode()
    finally:
        lib().fpx_wire_free(out)


class Batcher:
    """Request micro-batcher: what a host would put behind MultiIndex.search (MultiIndex.zig:287-330).  search() is
    thread-safe and blocking; concurrent callers are answered by shared GPU batches."""

    def __init__(self, ctx: Context, max_batch=0, max_wait_us=100):
        cfg = _ffi.BatcherConfig(max_batch, max_wait_us)
        self.h = C.c_void_p()
        check(lib().fpx_batcher_create(ctx.h, C.byref(cfg), C.byref(self.h)))

    def set_snapshot(self, snapshot: Optional[Snapshot]):
        """Index.swapSnapshot hook (Index.zig:469-485)."""
        check(lib().fpx_batcher_set_snapshot(self.h, snapshot.h if snapshot is not None else None))

    def search(self, request: SearchRequest, clamp_http=True) -> List[SearchResult]:
        limit = request.limit
        if clamp_http:
            limit = max(min(limit, max_search_limit), min_search_limit)   # server.zig:192
        min_score = request.min_score
        if min_score is None:
            min_score = lib().fpx_default_min_score(len(request.query))   # MultiIndex.zig:304, RAW length
        q = _u32([int(x) & 0xFFFFFFFF for x in request.query])
        opts = np.array([limit, min_score, request.score_pct], dtype=np.uint32)
        cap = max(1, min(int(limit), _ffi.FPX_MAX_RESULTS))
        ids, sc, n = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32), C.c_uint32(0)
        check(lib().fpx_batcher_search(self.h, q.ctypes.data if len(q) else None, len(q), opts.ctypes.data,
                                       int(request.timeout), ids.ctypes.data, sc.ctypes.data, cap, C.byref(n)))
        return [SearchResult(int(ids[i]), int(sc[i])) for i in range(n.value)]

    def stats(self):
        s = _ffi.BatcherStats()
        check(lib().fpx_batcher_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in _ffi.BatcherStats._fields_}

    def close(self):
        if self.h:
            lib().fpx_batcher_destroy(self.h)
            self.h = None


def multi_index_search(reader: IndexReader, request: SearchRequest, clamp_http=True) -> List[SearchResult]:
    """MultiIndex.search option mapping (MultiIndex.zig:302-306); clamp_http applies server.zig:192-193."""
    limit = request.limit
    if clamp_http:
        limit = max(min(limit, max_search_limit), min_search_limit)
    min_score = request.min_score
    if min_score is None:
        min_score = lib().fpx_default_min_score(len(request.query))  # RAW length, before de-duplication
    return reader.search(request.query, SearchOptions(limit, min_score, request.score_pct))


def pack_results_device(nq, k_stride, d_ids, d_scores, d_counts, d_packed, capacity_pairs, stream=0):
    """Device pointers (ints).  Packs a batch's result arrays for an exchange between GPUs (see fpx.h)."""
    check(lib().fpx_pack_results_device(nq, k_stride, d_ids, d_scores, d_counts, d_packed, capacity_pairs, stream))


def merge_packed_shards_device(n_shards, nq, d_packed, shard_stride_words, d_opts, k_stride, d_ids, d_scores, d_counts,
                               stream=0):
    """Device pointers (ints).  Merges n_shards packed result blocks into k_stride-wide arrays (see fpx.h)."""
    check(lib().fpx_merge_packed_shards_device(n_shards, nq, d_packed, shard_stride_words, d_opts, k_stride, d_ids,
                                               d_scores, d_counts, stream))


def unpack_results(packed, nq, k_stride, capacity_pairs):
    """numpy view of one rank's packed results -> (ids, scores, counts) in the k_stride-wide layout."""
    packed = np.asarray(packed).view(np.uint32)
    counts = packed[:nq].copy()
    offs = packed[nq:2 * nq + 1].astype(np.int64)
    if int(offs[nq]) > capacity_pairs:
        raise FpxError(_ffi.FPX_UNSUPPORTED, "packed results exceed the exchange capacity")
    pairs = packed[2 * nq + 2:2 * nq + 2 + 2 * int(offs[nq])].reshape(-1, 2)
    ids = np.zeros((nq, k_stride), np.uint32)
    sc = np.zeros((nq, k_stride), np.uint32)
    rows = np.repeat(np.arange(nq), counts)
    cols = np.arange(len(rows)) - np.repeat(offs[:nq], counts)
    ids[rows, cols] = pairs[:, 0]
    sc[rows, cols] = pairs[:, 1]
    return ids, sc, counts


def merge_shard_results(ids, scores, counts, opts, k_stride):
    """ids/scores: [n_shards, nq, k_stride]; counts: [n_shards, nq]; opts: [nq, 3] -> merged (ids, scores, counts)."""
    ids, scores, counts, opts = _u32(ids), _u32(scores), _u32(counts), _u32(opts)
    g, nq = counts.shape
    out = (np.zeros((nq, k_stride), np.uint32), np.zeros((nq, k_stride), np.uint32), np.zeros(nq, np.uint32))
    check(lib().fpx_merge_shard_results(g, nq, k_stride, ids.ctypes.data, scores.ctypes.data, counts.ctypes.data,
                                        opts.ctypes.data, out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data))
    return out
