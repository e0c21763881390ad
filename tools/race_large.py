#!/usr/bin/env python
"""Small workload of queries that almost fill a stage of the hot kernel, for compute-sanitizer runs: 105-term queries on 100 K docs x 120 hashes (rows of ~92
postings, ~2450 of a stage's 2552 granules per query), checked against the oracle.   python tools/race_large.py [n_queries] [repeats]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft
from _oracle import OracleIndex
pkg = graft.load_package()
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 600
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = pkg.synth.SynthConfig(n_docs=100_000, hashes_per_doc=120, vocab_log2=17, seed=0xF1D00001 + 9)
syn = pkg.synth.Synth(cfg, device="cuda:0")
items, doc_ids, doc_alive = syn.corpus_items()
seg = pkg.FileSegment.from_items(items, doc_ids, doc_alive, commit_id=1)
ctx = pkg.Context(device=0, profile=True)
snap = pkg.swap_snapshot(ctx, [seg])
ix = OracleIndex()
ix.adopt_file_segment(1, 0, seg.block_size, seg.blocks, seg.num_blocks, seg.block_index, seg.doc_ids, seg.doc_alive)
terms, _ = syn.queries(nq, 105, seed=77)
offs = np.arange(nq + 1, dtype=np.uint64) * 105
opts = pkg.synth.http_opts(nq, 105)
oi, os_, oc, _ = ix.search_batch(terms.reshape(-1), offs, opts, 40, n_threads=8)
mask = np.arange(40)[None, :] < oc[:, None]
bad = 0
for r in range(reps):
    ctx.profile_reset()
    ids, sc, cnt = pkg.IndexReader(snap).search_batch(terms.reshape(-1), offs, opts, 40)
    p = ctx.profile()
    ok = np.array_equal(cnt, oc) and np.array_equal(ids[mask], oi[mask]) and np.array_equal(sc[mask], os_[mask])
    if not ok:
        bad += 1
        d = np.nonzero((np.where(mask, sc, 0) != np.where(mask, os_, 0)).any(axis=1) | (cnt != oc))[0]
        print("rep", r, "MISMATCH at", d[:5], "gpu", sc[d[0], :3], "oracle", os_[d[0], :3], flush=True)
print("sketch_queries", p["sketch_queries"], "requeues", p["overflow_requeues"], "reps", reps, "bad", bad)
