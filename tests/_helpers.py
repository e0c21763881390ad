"""Shared test helpers: package loading, oracle -> product segment hand-over, a numpy ranking of CSR counts."""
import ctypes as C
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)


def segments_from_oracle(ix):
    """Copy the oracle index's segments (its Segments snapshot) into product FileSegment / MemorySegment objects,
    i.e. exactly the bytes the reference holds in RAM."""
    files, mems = [], []
    for i in range(ix.num_file_segments):
        v = ix.file_segment(i)
        nb, bs = int(v.num_blocks), int(v.block_size)
        blocks = _arr(v.blocks, (nb + 1) * bs, np.uint8)
        index = _arr(v.block_index, nb, np.uint32)
        files.append(pkg.FileSegment(v.commit_id, v.merges, v.min_doc_id, bs, blocks, nb, index,
                                     _arr(v.doc_ids, v.n_docs, np.uint32), _arr(v.doc_alive, v.n_docs, np.uint8)))
    for i in range(ix.num_memory_segments):
        v = ix.memory_segment(i)
        mems.append(pkg.MemorySegment(v.commit_id, v.merges, _arr(v.items, v.n_items, np.uint64),
                                      _arr(v.doc_ids, v.n_docs, np.uint32), _arr(v.doc_alive, v.n_docs, np.uint8)))
    return files, mems


def csr_rank(terms, offs, docids, query, max_results, min_score, min_score_pct):
    """score(id) = sum over unique query terms of the id's multiplicity in CSR[term]; then the ranking of
    common.zig:131-167 with the hasNewerCommit test removed (dead postings are already gone)."""
    q = np.unique(np.asarray(query, dtype=np.uint32))
    pos = np.searchsorted(terms, q)
    rows = []
    for p, t in zip(pos, q):
        if p < len(terms) and terms[p] == t:
            rows.append(docids[int(offs[p]):int(offs[p + 1])])
    if not rows:
        return []
    ids, counts = np.unique(np.concatenate(rows), return_counts=True)
    keep = counts >= min_score
    ids, counts = ids[keep], counts[keep]
    order = np.lexsort((ids, -counts.astype(np.int64)))
    out, ms = [], min_score
    for i in order:
        if len(out) == max_results:
            break
        s = int(counts[i])
        if s < ms:
            break
        if not out:
            ms = max(ms, ((s * min_score_pct) & 0xFFFFFFFF) // 100)
        out.append((int(ids[i]), s))
    return out


def flat_queries(queries):
    """list of term lists -> (terms u32, offsets u64)."""
    offs = np.zeros(len(queries) + 1, dtype=np.uint64)
    for i, q in enumerate(queries):
        offs[i + 1] = offs[i] + len(q)
    terms = np.array([int(t) & 0xFFFFFFFF for q in queries for t in q] or [0], dtype=np.uint32)
    return terms, offs
