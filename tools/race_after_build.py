#!/usr/bin/env python
"""Does a search right after a snapshot build differ from the same search at rest?  C2 corpus, 10 K x 100-term queries;
every iteration builds (and drops) another snapshot first.   python tools/race_after_build.py [iterations] [mode]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
pkg = graft.load_package()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
mode = sys.argv[2] if len(sys.argv) > 2 else "build"
cfg = pkg.synth.SynthConfig(n_docs=1_000_000, hashes_per_doc=100, vocab_log2=20, seed=0xF1D00001 + 2)
syn = pkg.synth.Synth(cfg, device="cuda:0")
items, doc_ids, doc_alive = syn.corpus_items()
seg = pkg.FileSegment.from_items(items, doc_ids, doc_alive, commit_id=1)
del items
ctx = pkg.Context(device=0, profile=True)
snap = pkg.swap_snapshot(ctx, [seg])
reader = pkg.IndexReader(snap)
terms, _ = syn.queries(10_000, 100, seed=0xF1D01001 + 2)
offs = np.arange(10_001, dtype=np.uint64) * 100
opts = pkg.synth.http_opts(10_000, 100)
ref = reader.search_batch(terms.reshape(-1), offs, opts, 40)
for _ in range(3):
    again = reader.search_batch(terms.reshape(-1), offs, opts, 40)
    assert all(np.array_equal(a, b) for a, b in zip(ref, again))
bad = 0
for it in range(iters):
    if mode == "build":
        other = pkg.swap_snapshot(ctx, [seg], doc_range=(1 + 1000 * it, 600_000))
    got = reader.search_batch(terms.reshape(-1), offs, opts, 40)
    if mode == "build":
        other.release()
    if not all(np.array_equal(a, b) for a, b in zip(ref, got)):
        bad += 1
        d = np.nonzero((ref[1] != got[1]).any(axis=1) | (ref[2] != got[2]))[0]
        print("iteration", it, "differs at queries", d[:8], "ref scores", ref[1][d[0], :3], "got", got[1][d[0], :3], "ids", ref[0][d[0], :2], got[0][d[0], :2], flush=True)
print("iterations", iters, "mode", mode, "bad", bad, ctx.profile()["overflow_requeues"])
