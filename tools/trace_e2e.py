#!/usr/bin/env python
"""Timeline of one fpx_search_batch call (FPX debug bit 11): python tools/trace_e2e.py [chunk]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
import __graft_entry__ as graft
pkg = graft.load_package()
wl = "c3"
syn, items, doc_ids, doc_alive = bench.build_corpus(pkg, wl, "cuda:0")
seg = pkg.FileSegment.from_items(items, doc_ids, doc_alive, commit_id=1, threads=os.cpu_count())
del items
ctx = pkg.Context(device=0, profile=False, host_threads=os.cpu_count())
snap = pkg.swap_snapshot(ctx, [seg])
reader = pkg.IndexReader(snap)
terms, offs, nq, T = bench.make_queries(syn, wl, 0)
opts = pkg.synth.http_opts(nq, T)
K = bench.K_STRIDE
h = [torch.from_numpy(terms.reshape(-1).view(np.int32).copy()).pin_memory(), torch.from_numpy(offs.view(np.int64).copy()).pin_memory(),
     torch.from_numpy(opts.view(np.int32).copy()).pin_memory(), torch.zeros((nq, K), dtype=torch.int32).pin_memory(),
     torch.zeros((nq, K), dtype=torch.int32).pin_memory(), torch.zeros(nq, dtype=torch.int32).pin_memory()]
for ch in [int(x) for x in (sys.argv[1:] or ["32768"])]:
    ctx.set_chunk_queries(ch)
    for i in range(4):
        ctx.debug_set(2048 if i == 3 else 0)
        reader.search_batch_ptr(nq, h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), K, h[3].data_ptr(), h[4].data_ptr(), h[5].data_ptr())
    print("---- chunk", ch, flush=True)
