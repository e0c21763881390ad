#!/usr/bin/env python
"""Build one workload once, then time the host-buffer call fpx_search_batch (pinned buffers) per chunk size."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--chunks", default="8192,16384,32768,65536")
    ap.add_argument("--variant", default="0")
    args = ap.parse_args()
    import torch
    import __graft_entry__ as graft
    pkg = graft.load_package()
    dev = torch.device("cuda", 0)
    wl = args.workload
    syn, items, doc_ids, doc_alive = bench.build_corpus(pkg, wl, str(dev))
    seg = pkg.FileSegment.from_items(items, doc_ids, doc_alive, commit_id=1, threads=os.cpu_count())
    del items
    ctx = pkg.Context(device=0, profile=True, host_threads=os.cpu_count())
    ctx.debug_set(int(args.variant, 0))
    snap = pkg.swap_snapshot(ctx, [seg])
    reader = pkg.IndexReader(snap)
    terms, offs, nq, T = bench.make_queries(syn, wl, 0)
    opts = pkg.synth.http_opts(nq, T)
    K = bench.K_STRIDE
    h_terms = torch.from_numpy(terms.reshape(-1).view(np.int32).copy()).pin_memory()
    h_offs = torch.from_numpy(offs.view(np.int64).copy()).pin_memory()
    h_opts = torch.from_numpy(opts.view(np.int32).copy()).pin_memory()
    h_ids = torch.zeros((nq, K), dtype=torch.int32).pin_memory()
    h_sc = torch.zeros((nq, K), dtype=torch.int32).pin_memory()
    h_cnt = torch.zeros(nq, dtype=torch.int32).pin_memory()

    def step():
        reader.search_batch_ptr(nq, h_terms.data_ptr(), h_offs.data_ptr(), h_opts.data_ptr(), K,
                                h_ids.data_ptr(), h_sc.data_ptr(), h_cnt.data_ptr())

    for ch in [int(x) for x in args.chunks.split(",")]:
        ctx.set_chunk_queries(ch)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ctx.profile_reset()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        p = ctx.profile()
        print("chunk %6d: e2e %.3f ms/step (%.1fM q/s) | h2d %.3f pack+d2h %.3f kernels %.3f ms (event time) | "
              "h2d %.1f MB d2h %.2f MB | results %d" % (
                  ch, dt * 1e3, nq / dt / 1e6, p["h2d_ms"] / args.steps, p["d2h_ms"] / args.steps,
                  (p["prepare_ms"] + p["sketch_ms"] + p["search_ms"] + p["wide_ms"]) / args.steps,
                  p["h2d_bytes"] / args.steps / 1e6, p["d2h_bytes"] / args.steps / 1e6, int(h_cnt.sum())), flush=True)
    snap.release()
    ctx.close()


if __name__ == "__main__":
    main()
