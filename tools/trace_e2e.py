#!/usr/bin/env python
"""The host-buffer call fpx_search_batch_packed on c3 with pinned buffers: wall time per chunk size, and the per-chunk
timeline of one call (FPX debug bit 11).   python tools/trace_e2e.py [chunk ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
import __graft_entry__ as graft
pkg = graft.load_package()
wl = os.environ.get("WL", "c3")
syn = bench.make_synth(pkg, wl, "cuda:0")
ctx = pkg.Context(device=0, profile=False, host_threads=os.cpu_count())
snap, _ = bench.build_snapshot(pkg, ctx, syn, wl, os.cpu_count(), keep_segments=False)
reader = pkg.IndexReader(snap)
terms, offs, nq, T = bench.make_queries(syn, wl, 0, 1, "replicated")
opts = pkg.synth.http_opts(nq, T)
K = bench.K_STRIDE
h = [torch.from_numpy(terms.reshape(-1).view(np.int32).copy()).pin_memory(), torch.from_numpy(offs.view(np.int64).copy()).pin_memory(),
     torch.from_numpy(opts.view(np.int32).copy()).pin_memory(), torch.zeros(nq, dtype=torch.int32).pin_memory(),
     torch.zeros((nq * K, 2), dtype=torch.int32).pin_memory()]


def step():
    reader.search_batch_packed_ptr(nq, h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), K, h[3].data_ptr(), h[4].data_ptr(), nq * K)


TRACE = os.environ.get("TRACE", "1") == "1"
for arg in (sys.argv[1:] or ["131072"]):  # chunk[:schedule variant] (FPX debug bits 16..18, see search_batch_host)
    ch, _, sv = arg.partition(":")
    ch, sv = int(ch), int(sv or 0) << 16
    ctx.set_chunk_queries(ch)
    ctx.debug_set(sv)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 20 * 1e3
    print("chunk %6d sched %d: %.3f ms per call, %.1fM q/s" % (ch, sv >> 16, ms, nq / ms / 1e3), flush=True)
    if TRACE:
        ctx.debug_set(2048 | sv)
        step()
    ctx.debug_set(0)
    print("---- chunk", arg, flush=True)
