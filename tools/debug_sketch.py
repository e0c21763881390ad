import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import __graft_entry__ as g
pkg = g.load_package()
cfg = pkg.synth.SynthConfig(n_docs=200_000, hashes_per_doc=100, vocab_log2=18, seed=5)
syn = pkg.synth.Synth(cfg, device="cuda:0")
items, ids, alive = syn.corpus_items()
seg = pkg.FileSegment.from_items(items, ids, alive, 1)
ctx = pkg.Context(device=0, profile=True)
snap = pkg.swap_snapshot(ctx, [seg])
r = pkg.IndexReader(snap)
for T in (20, 40, 60, 100):
    terms, _ = syn.queries(3000, T, seed=321)
    nq = terms.shape[0]
    offs = np.arange(nq + 1, dtype=np.uint64) * T
    for opt in ((40, 5, 10), (40, 2, 0)):
        opts = np.tile(np.array(opt, dtype=np.uint32), (nq, 1))
        ctx.profile_reset()
        a = r.search_batch(terms.reshape(-1), offs, opts, 40)
        p = ctx.profile()
        print(T, opt, {k: p[k] for k in ("queries", "postings", "sketch_queries", "wide_queries", "overflow_requeues", "results", "sketch_ms", "search_ms")})
