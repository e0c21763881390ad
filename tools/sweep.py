#!/usr/bin/env python
"""Build one workload once, then time the device-resident batch under several FPX debug/variant settings.

  python tools/sweep.py --workload c3 --steps 5 --variants 0,512,513,514,520,528

Each variant value is passed to fpx_debug_set (see the bit table in csrc/fpx_kernels.cu).  Variants with
ablation bits give wrong results; variants marked 'v:' in the table are real kernel variants and are checked
against variant 0's results.  Prints one line per variant (CUDA-event times from fpx_profile).
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--variants", default="0,512")
    ap.add_argument("--check", default="", help="comma list of variants whose results must equal variant 0's")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import __graft_entry__ as graft
    pkg = graft.load_package()
    dev = torch.device("cuda", 0)
    wl = args.workload
    syn = bench.make_synth(pkg, wl, str(dev))
    ctx = pkg.Context(device=0, profile=True, host_threads=os.cpu_count())
    snap, _ = bench.build_snapshot(pkg, ctx, syn, wl, os.cpu_count(), keep_segments=False)
    reader = pkg.IndexReader(snap)
    terms, offs, nq, T = bench.make_queries(syn, wl, 0, 1, "replicated")
    opts = pkg.synth.http_opts(nq, T)
    torch.cuda.empty_cache()
    d_terms = torch.from_numpy(terms.reshape(-1).view(np.int32)).to(dev)
    d_offs = torch.from_numpy(offs.view(np.int64)).to(dev)
    d_opts = torch.from_numpy(opts.view(np.int32)).to(dev)
    K = bench.K_STRIDE
    d_ids = torch.zeros((nq, K), dtype=torch.int32, device=dev)
    d_sc = torch.zeros((nq, K), dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        reader.search_batch_device(nq, d_terms.data_ptr(), d_offs.data_ptr(), d_opts.data_ptr(), K,
                                   d_ids.data_ptr(), d_sc.data_ptr(), d_cnt.data_ptr(), stream.cuda_stream)

    ref = None
    check = set(int(x, 0) for x in args.check.split(",") if x)
    rows = []
    for v in [int(x, 0) for x in args.variants.split(",")]:
        ctx.debug_set(v)
        d_ids.zero_(); d_sc.zero_(); d_cnt.zero_()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ctx.profile_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        p = ctx.profile()
        res = (d_cnt.cpu().numpy().copy(), d_ids.cpu().numpy().copy(), d_sc.cpu().numpy().copy())
        ok = ""
        if v == 0:
            ref = res
        elif v in check and ref is not None:
            mask = np.arange(K)[None, :] < ref[0][:, None]
            same = (np.array_equal(res[0], ref[0]) and np.array_equal(res[1][mask], ref[1][mask]) and
                    np.array_equal(res[2][mask], ref[2][mask]))
            ok = " parity_vs_v0=%s" % same
        line = ("variant %5d: total %.3f ms/step (%.1fM q/s) | prepare %.3f sketch %.3f exact %.3f wide %.3f | "
                "sketch_q %d requeues %d results %d%s" % (
                    v, ms, nq / ms / 1e3, p["prepare_ms"] / args.steps, p["sketch_ms"] / args.steps,
                    p["search_ms"] / args.steps, p["wide_ms"] / args.steps, p["sketch_queries"] / args.steps,
                    p["overflow_requeues"] / args.steps, p["results"] / args.steps, ok))
        print(line, flush=True)
        rows.append(line)
    if args.out:
        open(args.out, "w").write("\n".join(rows) + "\n")
    snap.release()
    ctx.close()


if __name__ == "__main__":
    main()
