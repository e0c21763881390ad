#!/usr/bin/env python
"""Host-buffer call from one and from two host threads, several batch sizes (WL=c2|c3): ms per call and, with TRACE=1, the
per-chunk timeline of both callers' last calls.   python tools/two_callers.py [n_queries ...]"""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
import __graft_entry__ as graft
pkg = graft.load_package()
wl = os.environ.get("WL", "c2")
syn = bench.make_synth(pkg, wl, "cuda:0")
ctx = pkg.Context(device=0, profile=False, host_threads=os.cpu_count())
snap, _ = bench.build_snapshot(pkg, ctx, syn, wl, os.cpu_count(), keep_segments=False)
reader = pkg.IndexReader(snap)
terms, offs, nq_all, T = bench.make_queries(syn, wl, 0, 1, "replicated")
opts = pkg.synth.http_opts(nq_all, T)
K = bench.K_STRIDE
h_terms = torch.from_numpy(terms.reshape(-1).view(np.int32).copy()).pin_memory()
h_offs = torch.from_numpy(offs.view(np.int64).copy()).pin_memory()
h_opts = torch.from_numpy(opts.view(np.int32).copy()).pin_memory()
outs = [(torch.zeros(nq_all, dtype=torch.int32).pin_memory(), torch.zeros((nq_all * K, 2), dtype=torch.int32).pin_memory()) for _ in range(2)]


def call(k, nq):
    c, pr = outs[k]
    reader.search_batch_packed_ptr(nq, h_terms.data_ptr(), h_offs.data_ptr(), h_opts.data_ptr(), K, c.data_ptr(), pr.data_ptr(), nq * K)


for nq in [int(x) for x in (sys.argv[1:] or [str(nq_all)])]:
    nq = min(nq, nq_all)
    for _ in range(3):
        call(0, nq)
    t0 = time.perf_counter()
    for _ in range(20):
        call(0, nq)
    one = (time.perf_counter() - t0) / 20 * 1e3
    gate = threading.Barrier(3)
    lat = [[], []]

    def worker(k):
        for _ in range(3):
            call(k, nq)
        gate.wait(); gate.wait()
        for _ in range(10):
            t = time.perf_counter(); call(k, nq); lat[k].append((time.perf_counter() - t) * 1e3)
        gate.wait()
    th = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    [t.start() for t in th]
    gate.wait(); t0 = time.perf_counter(); gate.wait(); gate.wait()
    two = (time.perf_counter() - t0) / 20 * 1e3
    [t.join() for t in th]
    print("nq %6d: one caller %.3f ms/call (%.1fM q/s) | two callers %.3f ms/call overall (%.1fM q/s), a call takes %.3f / %.3f ms (median)"
          % (nq, one, nq / one / 1e3, two, nq / two / 1e3, np.median(lat[0]), np.median(lat[1])), flush=True)
    if os.environ.get("TRACE") == "1":
        ctx.debug_set(2048)
        th = [threading.Thread(target=call, args=(k, nq)) for k in range(2)]
        [t.start() for t in th]; [t.join() for t in th]
        ctx.debug_set(0)
