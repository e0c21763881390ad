"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference's `_search` path
(oracle/fpindex_oracle.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


class FileSegmentView(C.Structure):
    _fields_ = [
        ("commit_id", C.c_uint64), ("merges", C.c_uint64),
        ("min_doc_id", C.c_uint32), ("max_doc_id", C.c_uint32),
        ("block_size", C.c_uint32), ("_pad", C.c_uint32),
        ("num_blocks", C.c_uint64), ("num_items", C.c_uint64),
        ("blocks", u8p), ("block_index", u32p),
        ("doc_ids", u32p), ("doc_alive", u8p), ("n_docs", C.c_uint64),
    ]


class MemorySegmentView(C.Structure):
    _fields_ = [
        ("commit_id", C.c_uint64), ("merges", C.c_uint64),
        ("min_doc_id", C.c_uint32), ("max_doc_id", C.c_uint32),
        ("items", u64p), ("n_items", C.c_uint64),
        ("doc_ids", u32p), ("doc_alive", u8p), ("n_docs", C.c_uint64),
    ]


def build(force=False):
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "fpindex_oracle.cpp")
    hdr = os.path.join(ORACLE_DIR, "fpindex_oracle.h")
    stale = (not os.path.exists(so)) or any(
        os.path.getmtime(p) > os.path.getmtime(so) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(so):
        so = build()
    L = C.CDLL(so)
    L.orc_index_new.restype = C.c_void_p
    L.orc_index_new.argtypes = [C.c_uint32]
    L.orc_index_free.argtypes = [C.c_void_p]
    L.orc_update.argtypes = [C.c_void_p, C.c_size_t, u8p, u32p, u64p, u32p]
    L.orc_checkpoint.argtypes = [C.c_void_p]
    L.orc_merge_memory.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
    L.orc_merge_files.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
    L.orc_add_file_segment_sorted.argtypes = [C.c_void_p, u64p, C.c_size_t, u32p, u8p, C.c_size_t]
    L.orc_adopt_file_segment.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p,
                                         C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.orc_num_file_segments.restype = C.c_size_t
    L.orc_num_file_segments.argtypes = [C.c_void_p]
    L.orc_num_memory_segments.restype = C.c_size_t
    L.orc_num_memory_segments.argtypes = [C.c_void_p]
    L.orc_file_segment.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(FileSegmentView)]
    L.orc_memory_segment.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(MemorySegmentView)]
    L.orc_search.restype = C.c_int64
    L.orc_search.argtypes = [C.c_void_p, u32p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32,
                             u32p, u32p, C.c_size_t]
    L.orc_search_batch.restype = C.c_double
    L.orc_search_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]
    L.orc_svb_encode_quad_0124.restype = C.c_size_t
    L.orc_svb_encode_quad_0124.argtypes = [u32p, u8p, u8p]
    L.orc_svb_encode_quad_1234.restype = C.c_size_t
    L.orc_svb_encode_quad_1234.argtypes = [u32p, u8p, u8p]
    L.orc_svb_decode_quad.restype = C.c_size_t
    L.orc_svb_decode_quad.argtypes = [C.c_int, C.c_uint8, u8p, u32p]
    L.orc_svb_decode_quad_delta.restype = C.c_size_t
    L.orc_svb_decode_quad_delta.argtypes = [C.c_int, C.c_uint8, u8p, u32p, C.c_uint32]
    L.orc_svb_delta_decode_in_place.argtypes = [u32p, C.c_size_t, C.c_uint32]
    L.orc_svb_decode_values.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, u8p, u32p, C.c_int,
                                        C.c_int, C.c_uint32]
    L.orc_encode_block.restype = C.c_size_t
    L.orc_encode_block.argtypes = [u64p, C.c_size_t, C.c_uint32, u8p, C.c_size_t]
    L.orc_decode_block.restype = C.c_size_t
    L.orc_decode_block.argtypes = [u8p, C.c_size_t, C.c_uint32, u32p, u32p]
    L.orc_block_search_hash.restype = C.c_size_t
    L.orc_block_search_hash.argtypes = [u8p, C.c_size_t, C.c_uint32, C.c_uint32, u32p, u32p, u32p]
    L.orc_uses_ssse3.restype = C.c_int
    _LIB = L
    return L


def _p(a, t):
    return a.ctypes.data_as(t)


def items_u64(pairs):
    """[(hash, id), ...] -> sorted u64 array (hash<<32 | id), segment.zig:87-106."""
    a = np.array([(int(h) << 32) | int(i) for h, i in pairs], dtype=np.uint64)
    a.sort()
    return a


class OracleIndex:
    """Mirror of the reference's Index (update / checkpoint / merge) + IndexReader.search."""

    def __init__(self, block_size=0):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_index_new(block_size))
        self._keep = []

    def close(self):
        if self.h:
            self.L.orc_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # changes: list of ("insert", id, [hashes]) / ("delete", id)
    def update(self, changes):
        n = len(changes)
        kinds = np.zeros(n, dtype=np.uint8)
        ids = np.zeros(n, dtype=np.uint32)
        offs = np.zeros(n + 1, dtype=np.uint64)
        hs = []
        for i, ch in enumerate(changes):
            if ch[0] == "insert":
                kinds[i] = 0
                ids[i] = ch[1]
                hs.extend(int(x) & 0xFFFFFFFF for x in ch[2])
            elif ch[0] == "delete":
                kinds[i] = 1
                ids[i] = ch[1]
            else:
                raise ValueError(ch[0])
            offs[i + 1] = len(hs)
        hashes = np.array(hs if hs else [0], dtype=np.uint32)
        rc = self.L.orc_update(self.h, n, _p(kinds, u8p), _p(ids, u32p), _p(offs, u64p), _p(hashes, u32p))
        assert rc == 0, rc

    def update_arrays(self, kinds, ids, offs, hashes):
        kinds = np.ascontiguousarray(kinds, dtype=np.uint8)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        hashes = np.ascontiguousarray(hashes, dtype=np.uint32)
        rc = self.L.orc_update(self.h, len(kinds), _p(kinds, u8p), _p(ids, u32p), _p(offs, u64p), _p(hashes, u32p))
        assert rc == 0, rc

    def checkpoint(self):
        return self.L.orc_checkpoint(self.h)

    def merge_memory(self, lo, count):
        assert self.L.orc_merge_memory(self.h, lo, count) == 0

    def merge_files(self, lo, count):
        assert self.L.orc_merge_files(self.h, lo, count) == 0

    def add_file_segment_sorted(self, items, doc_ids, doc_alive):
        items = np.ascontiguousarray(items, dtype=np.uint64)
        doc_ids = np.ascontiguousarray(doc_ids, dtype=np.uint32)
        doc_alive = np.ascontiguousarray(doc_alive, dtype=np.uint8)
        rc = self.L.orc_add_file_segment_sorted(self.h, _p(items, u64p), len(items), _p(doc_ids, u32p),
                                                _p(doc_alive, u8p), len(doc_ids))
        assert rc == 0, rc

    def adopt_file_segment(self, commit_id, merges, block_size, blocks, n_blocks, block_index, doc_ids, doc_alive):
        """blocks/block_index/doc arrays are borrowed: numpy arrays kept alive here."""
        self._keep.append((blocks, block_index, doc_ids, doc_alive))
        rc = self.L.orc_adopt_file_segment(self.h, commit_id, merges, block_size, blocks.ctypes.data, n_blocks,
                                           block_index.ctypes.data, doc_ids.ctypes.data, doc_alive.ctypes.data,
                                           len(doc_ids))
        assert rc == 0, rc

    @property
    def num_file_segments(self):
        return self.L.orc_num_file_segments(self.h)

    @property
    def num_memory_segments(self):
        return self.L.orc_num_memory_segments(self.h)

    def file_segment(self, i):
        v = FileSegmentView()
        assert self.L.orc_file_segment(self.h, i, C.byref(v)) == 0
        return v

    def memory_segment(self, i):
        v = MemorySegmentView()
        assert self.L.orc_memory_segment(self.h, i, C.byref(v)) == 0
        return v

    def search(self, query, max_results=10, min_score=1, min_score_pct=10):
        q = np.array([int(x) & 0xFFFFFFFF for x in query], dtype=np.uint32)
        cap = max(int(max_results), 1)
        cap = min(cap, 1 << 20)
        ids = np.zeros(cap, dtype=np.uint32)
        sc = np.zeros(cap, dtype=np.uint32)
        qp = _p(q, u32p) if len(q) else None
        n = self.L.orc_search(self.h, qp, len(q), max_results, min_score, min_score_pct, _p(ids, u32p), _p(sc, u32p), cap)
        assert n >= 0
        n = min(n, cap)
        return [(int(ids[i]), int(sc[i])) for i in range(n)]

    def search_http(self, query, limit=40, min_score=None, score_pct=10):
        """MultiIndex.search option mapping (MultiIndex.zig:302-306) + HTTP clamp (server.zig:192)."""
        limit = max(min(limit, 100), 1)
        if min_score is None:
            min_score = (len(query) + 19) // 20
        return self.search(query, limit, min_score, score_pct)

    def search_batch(self, terms, offsets, opts3, k_stride, n_threads=1):
        terms = np.ascontiguousarray(terms, dtype=np.uint32)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        opts3 = np.ascontiguousarray(opts3, dtype=np.uint32)
        nq = len(offsets) - 1
        ids = np.zeros((nq, k_stride), dtype=np.uint32)
        sc = np.zeros((nq, k_stride), dtype=np.uint32)
        cnt = np.zeros(nq, dtype=np.uint32)
        secs = self.L.orc_search_batch(self.h, nq, terms.ctypes.data, offsets.ctypes.data, opts3.ctypes.data,
                                       k_stride, ids.ctypes.data, sc.ctypes.data, cnt.ctypes.data, n_threads)
        assert secs >= 0
        return ids, sc, cnt, secs
