#!/usr/bin/env python
"""bench.py — queries/sec of the batched `_search` path on synthetic fingerprint corpora.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl fpx|reference] [--workload c3|c2|c4|c5|tiny]
                  [--mode replicated|sharded]

One step = one pass of the hot path over one batch of synthetic queries.  Workload c3 (default) is the
configuration BASELINE.json's metric is quoted on: 10 M fingerprints x 120 hashes, 100 K-query batch of
100-term queries, HTTP default options (limit 40, min_score (T+19)/20, score_pct 10).

  value     whole-job queries/s with the batch already resident in HBM (fpx_search_batch_device), CUDA-event
            timed over exactly K steps, max over ranks
  e2e       the same batch through the host-buffer C-ABI call fpx_search_batch_packed (pinned host memory; H2D of
            the queries and D2H of the results — a list per query, as the reference returns them — inside the timed
            region)
  roofline  the dominant kernel (search_find_kernel: TMA row gather + count sketch + exact resolve + top-k) —
            algorithmic bytes / its CUDA-event time, against the measured HBM peak in MEASURED_PEAKS.json
  cpu_baseline  the C++ restatement of the reference CPU path (oracle/), all host threads, bounded sample

N > 1: one process per GPU (torchrun).  --mode replicated (default): every rank holds the whole corpus and answers
its own queries, no data-path collective (c3/c2/c5: one batch per rank, weak scaling; c4: the 1 M-query batch is split
over the ranks, strong scaling).  --mode sharded: rank g holds the postings of its docid range, every rank answers the
whole batch with the absolute floor only, the packed top-k lists are all-gathered over NCCL (sized by content) and
merged on the device (strong scaling).  c4 = 50 M fingerprints in five 10 M-doc file segments.
`--impl reference` times the CPU restatement alone (the Zig reference cannot be built here: no zig, no network).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_docs, hashes/doc, vocab_log2, zipf_s, n_queries, terms/query, corpus seed, query seed, file segments)
    "tiny": (50_000, 60, 16, 0.0, 4096, 100, 0xF1D00001, 0xF1D01001, 1),
    "tiny4": (200_000, 60, 18, 0.0, 8192, 100, 0xF1D00001, 0xF1D01001, 4),   # c4-shaped plumbing check
    "c2": (1_000_000, 100, 20, 0.0, 10_000, 100, 0xF1D00001 + 2, 0xF1D01001 + 2, 1),
    "c3": (10_000_000, 120, 24, 0.0, 100_000, 100, 0xF1D00001 + 3, 0xF1D01001 + 3, 1),
    "c4": (50_000_000, 120, 26, 0.0, 1_000_000, 100, 0xF1D00001 + 4, 0xF1D01001 + 4, 5),
    "c5": (10_000_000, 120, 20, 1.0, 100_000, 100, 0xF1D00001 + 5, 0xF1D01001 + 5, 1),
}
STRONG = {"c4", "tiny4"}   # replicated mode: one batch split over the ranks (others: one batch per rank)
K_STRIDE = 40
METRIC = "queries/sec at 10M fingerprints, 100-term queries"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def bind_near_gpu(torch, dev):
    """Multi-rank runs: keep this rank's threads on the CPUs of its GPU's NUMA node, so that its pinned staging memory
    (first touch) and the library's helper threads are local to the PCIe root the copies go through.  Round 1's e2e scaling
    (0.48 at 8 GPUs) was eight ranks moving their 42 MB per step across sockets.  Best effort: any failure leaves the
    affinity as it was."""
    try:
        pr = torch.cuda.get_device_properties(dev)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return "numa node unknown"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "no allowed cpu on node %d" % node
        os.sched_setaffinity(0, cpus)
        return "node %d, %d cpus" % (node, len(cpus))
    except Exception as e:   # noqa: BLE001 - diagnostics only
        return "not bound (%s)" % e


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_synth(pkg, wl, device):
    n_docs, H, vlog, zipf, *_ = WORKLOADS[wl]
    cfg = pkg.synth.SynthConfig(n_docs=n_docs, hashes_per_doc=H, vocab_log2=vlog, seed=WORKLOADS[wl][6], zipf_s=zipf)
    return pkg.synth.Synth(cfg, device=device)


def corpus_segments(pkg, syn, wl, threads, writer=None):
    """The workload's file segments, oldest first: consecutive doc ranges of the synthetic corpus, one at a time
    (generator; the caller decides what to keep).  writer: items -> segment (default: the product's block writer)."""
    n_docs, n_seg = WORKLOADS[wl][0], WORKLOADS[wl][8]
    per = (n_docs + n_seg - 1) // n_seg
    for k in range(n_seg):
        t = time.time()
        items, doc_ids, doc_alive = syn.corpus_items(doc_lo=k * per, doc_hi=min(n_docs, (k + 1) * per))
        t1 = time.time()
        if writer is None:
            seg = pkg.FileSegment.from_items(items, doc_ids, doc_alive, commit_id=k + 1, threads=threads)
            log("[bench] segment %d/%d: %d postings generated+sorted in %.1fs, %d blocks written in %.1fs"
                % (k + 1, n_seg, len(items), t1 - t, seg.num_blocks, time.time() - t1))
        else:
            seg = writer(items, doc_ids, doc_alive, k + 1)
            log("[bench] segment %d/%d: %d postings generated+sorted in %.1fs, written in %.1fs"
                % (k + 1, n_seg, len(items), t1 - t, time.time() - t1))
        del items
        yield seg


def build_snapshot(pkg, ctx, syn, wl, threads, keep_segments, doc_range=None):
    """Synthetic corpus -> reference-format file segments -> GPU snapshot (segments are handed to the builder as they
    are written and dropped unless the caller needs them for the CPU baseline)."""
    b = pkg.SnapshotBuilder(ctx)
    if doc_range is not None and doc_range != (0, 0):
        b.set_doc_range(*doc_range)
    kept = []
    for seg in corpus_segments(pkg, syn, wl, threads):
        b.add_file_segment(seg)
        if keep_segments:
            kept.append(seg)
        del seg
    t = time.time()
    snap = b.commit()
    info = snap.info()
    log("[bench] snapshot: %d segments, %d terms, %d postings, %.2f GB in HBM, committed in %.1fs"
        % (info["n_segments"], info["n_terms"], info["n_postings"], info["device_bytes"] / 1e9, time.time() - t))
    return snap, kept


def make_queries(syn, wl, rank, world, mode):
    """This rank's queries.  Weak workloads: its own batch (stream qseed + 1000 * rank).  Strong ones (c4) in
    replicated mode: its contiguous slice of the one batch.  Sharded mode: the whole batch on every rank."""
    nq, T, qseed = WORKLOADS[wl][4], WORKLOADS[wl][5], WORKLOADS[wl][7]
    first = 0
    if mode == "sharded":
        pass
    elif wl in STRONG:
        per = (nq + world - 1) // world
        first, nq = min(nq, rank * per), max(0, min(nq, (rank + 1) * per) - min(nq, rank * per))
    else:
        qseed += 1000 * rank
    terms, _ = syn.queries(nq, T, seed=qseed, first=first)
    offs = np.arange(nq + 1, dtype=np.uint64) * T
    return terms, offs, nq, T


def oracle_with(segments, OracleIndex):
    orc = OracleIndex()
    for seg in segments:
        orc.adopt_file_segment(seg.commit_id, 0, seg.block_size, seg.blocks, seg.num_blocks, seg.block_index, seg.doc_ids,
                               seg.doc_alive)
    return orc


def check_against_oracle(orc, terms, offs, opts, n_chk, threads, g_ids, g_sc, g_cnt):
    """Bit-exact comparison of the first n_chk queries: counts, ids AND scores."""
    oi, os_, oc, _ = orc.search_batch(terms, offs[:n_chk + 1], opts[:n_chk], K_STRIDE, n_threads=threads)
    mask = np.arange(K_STRIDE)[None, :] < oc[:, None]
    return bool(np.array_equal(g_cnt[:n_chk], oc) and np.array_equal(g_ids[:n_chk][mask], oi[mask]) and
                np.array_equal(g_sc[:n_chk][mask], os_[mask]))


def cpu_baseline(orc_index, terms, offs, opts, threads, budget_s=12.0):
    """Oracle (C++ restatement of the reference CPU path) on a bounded sample of the same batch."""
    nq = len(offs) - 1
    probe = min(nq, 64 * threads)
    secs = orc_index.search_batch(terms, offs[:probe + 1], opts[:probe], K_STRIDE, n_threads=threads)[3]
    rate = probe / max(secs, 1e-9)
    sample = int(min(nq, max(probe, rate * budget_s)))
    secs = orc_index.search_batch(terms, offs[:sample + 1], opts[:sample], K_STRIDE, n_threads=threads)[3]
    return sample / secs, sample


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port), all host threads, bounded sample per step."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import __graft_entry__ as graft
    from _oracle import OracleIndex, build as build_oracle
    build_oracle()
    pkg = graft.load_package()  # synthetic data generator only; no libfpx call on this arm
    wl = args.workload
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    syn = make_synth(pkg, wl, dev)
    orc = OracleIndex()

    def oracle_writer(items, doc_ids, doc_alive, commit_id):   # the oracle's own block writer
        orc.add_file_segment_sorted(items, doc_ids, doc_alive)
        return None

    for _ in corpus_segments(pkg, syn, wl, os.cpu_count() or 1, writer=oracle_writer):
        pass
    terms, offs, nq, T = make_queries(syn, wl, 0, 1, "replicated")
    opts = pkg.synth.http_opts(nq, T)
    threads = os.cpu_count() or 1
    flat = terms.reshape(-1)
    probe = min(nq, 64 * threads)
    secs = orc.search_batch(flat, offs[:probe + 1], opts[:probe], K_STRIDE, n_threads=threads)[3]
    per_step = int(min(nq, max(probe, probe / secs * 6.0)))   # ~6 s of CPU work per step
    for _ in range(args.warmup):
        orc.search_batch(flat, offs[:per_step + 1], opts[:per_step], K_STRIDE, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.search_batch(flat, offs[:per_step + 1], opts[:per_step], K_STRIDE, n_threads=threads)
    dt = time.perf_counter() - t0
    qps = per_step * args.steps / dt
    sample = "%d queries/step of the %s batch" % (per_step, wl)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": scaling_of(wl, args.mode), "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": workload_config(wl),
        "parallelism": "cpu threads x%d" % threads,
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "C++ restatement of the reference CPU path (Zig toolchain unavailable); not the reference binary",
    }), flush=True)


def scaling_of(wl, mode):
    return "strong" if (mode == "sharded" or wl in STRONG) else "weak"


def workload_config(wl):
    """Identical for both arms (the driver compares the strings); how the work is spread is reported beside it."""
    n_docs, H, vlog, zipf, nq, T, _, _, n_seg = WORKLOADS[wl]
    segs = "one merged file segment" if n_seg == 1 else "%d file segments of %d fingerprints" % (n_seg, n_docs // n_seg)
    return {"workload": "%s: %d fingerprints x %d hashes, vocab 2^%d%s, %d-query batch x %d terms, limit 40, "
                        "min_score (T+19)/20, score_pct 10, %s (512-byte blocks)"
                        % (wl, n_docs, H, vlog, ", Zipf s=%.1f" % zipf if zipf else "", nq, T, segs),
            "l2": "inputs larger than L2: CSR rows touched per step (~GBs) >> 126 MB L2; no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fpx", choices=["fpx", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="replicated", choices=["replicated", "sharded"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sketch", action="store_true", help="A/B: exact count-table kernels only")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer pass")
    ap.add_argument("--parity-queries", type=int, default=0,
                    help="check this many queries of rank 0 against the oracle (default: the CPU baseline's sample, <= 2000)")
    ap.add_argument("--chunk", type=int, default=0, help="queries per pipelined chunk of the host-buffer call (0 = library default)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import importlib
    import torch
    import __graft_entry__ as graft
    pkg = graft.load_package()
    multi_gpu = importlib.import_module("acoustid_index_b200.multi_gpu")   # needs torch.distributed: imported on demand
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the fpx search path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    wl, mode = args.workload, args.mode
    sharded = mode == "sharded"
    numa = bind_near_gpu(torch, dev) if world > 1 else "single rank: not bound"
    host_threads = max(1, (os.cpu_count() or 1) // max(1, world))
    n_docs = WORKLOADS[wl][0]
    want_oracle = rank == 0 and (args.parity_queries > 0 or (world == 1 and not args.no_cpu_baseline))

    # ---- setup (untimed): synthetic corpus -> reference-format segments -> GPU snapshot
    syn = make_synth(pkg, wl, str(dev))
    ctx = pkg.Context(device=local_rank, profile=True, host_threads=host_threads, no_sketch=args.no_sketch,
                      chunk_queries=args.chunk)
    doc_range = multi_gpu.doc_ranges(1, n_docs, world)[rank] if sharded else None
    snap, segments = build_snapshot(pkg, ctx, syn, wl, host_threads, keep_segments=want_oracle, doc_range=doc_range)
    reader = pkg.IndexReader(snap)
    terms, offs, nq, T = make_queries(syn, wl, rank, world, mode)
    nq_total = WORKLOADS[wl][4] if (sharded or wl in STRONG) else world * nq
    opts = pkg.synth.http_opts(nq, T)
    torch.cuda.empty_cache()

    # device-resident inputs/outputs
    d_terms = torch.from_numpy(terms.reshape(-1).view(np.int32)).to(dev)
    d_offs = torch.from_numpy(offs.view(np.int64)).to(dev)
    d_opts = torch.from_numpy(opts.view(np.int32)).to(dev)
    stream = torch.cuda.current_stream()
    if sharded:   # local searches use the absolute floor only; the merge applies the relative cutoff
        local_opts = opts.copy()
        local_opts[:, 2] = 0
        d_local_opts = torch.from_numpy(local_opts.view(np.int32)).to(dev)
        ss = multi_gpu.ShardedSearch(reader, nq, K_STRIDE, dev)
        d_ids, d_sc, d_cnt = ss.out_ids, ss.out_sc, ss.out_cnt

        def step():
            ss.step(d_terms, d_offs, d_local_opts, d_opts)
    else:
        d_ids = torch.zeros((nq, K_STRIDE), dtype=torch.int32, device=dev)
        d_sc = torch.zeros((nq, K_STRIDE), dtype=torch.int32, device=dev)
        d_cnt = torch.zeros(nq, dtype=torch.int32, device=dev)

        def step():   # replicated: every rank answers its own queries; nothing to exchange
            reader.search_batch_device(nq, d_terms.data_ptr(), d_offs.data_ptr(), d_opts.data_ptr(), K_STRIDE,
                                       d_ids.data_ptr(), d_sc.data_ptr(), d_cnt.data_ptr(), stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ctx.profile_reset()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    prof = ctx.profile()
    if world > 1:
        tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms_per_step = ms_total / args.steps
    value = nq_total / (ms_per_step * 1e-3)
    r_cnt = d_cnt.cpu().numpy().view(np.uint32)
    r_ids = d_ids.cpu().numpy().view(np.uint32)
    r_sc = d_sc.cpu().numpy().view(np.uint32)

    # ---- e2e through the host-buffer C-ABI call, pinned host memory, copies inside the timed region
    e2e = None
    if not args.no_e2e and not sharded:
        h_terms = torch.from_numpy(terms.reshape(-1).view(np.int32).copy()).pin_memory()
        h_offs = torch.from_numpy(offs.view(np.int64).copy()).pin_memory()
        h_opts = torch.from_numpy(opts.view(np.int32).copy()).pin_memory()
        # results as the reference hands them out: a list per query (fpx_search_batch_packed: counts + pairs back to back)
        h_cnt = torch.zeros(nq, dtype=torch.int32).pin_memory()
        h_pairs = torch.zeros((nq * K_STRIDE, 2), dtype=torch.int32).pin_memory()
        n_pairs = [0]

        def e2e_step():
            n_pairs[0] = reader.search_batch_packed_ptr(nq, h_terms.data_ptr(), h_offs.data_ptr(), h_opts.data_ptr(), K_STRIDE,
                                                        h_cnt.data_ptr(), h_pairs.data_ptr(), nq * K_STRIDE)

        for _ in range(3):
            e2e_step()
        barrier()
        ctx.set_profile(False)     # the timed e2e steps run without the library's own event recording
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        ctx.set_profile(True)      # ... and the same steps again, recorded, for the breakdown below
        ctx.profile_reset()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_s = float(tt.item())
        prof_e2e = ctx.profile()
        h2d = int(h_terms.numel() * 4 + h_offs.numel() * 8 + h_opts.numel() * 4)
        d2h = int(prof_e2e["d2h_bytes"] // args.steps)   # counted by the library from what it moved
        # the e2e results must equal the device-resident ones
        e_cnt = h_cnt.numpy().view(np.uint32)
        assert np.array_equal(e_cnt, r_cnt) and n_pairs[0] == int(r_cnt.sum()), "e2e and device-resident results differ"
        _m = np.arange(K_STRIDE)[None, :] < r_cnt[:, None]
        e_pairs = h_pairs.numpy().view(np.uint32)[:n_pairs[0]]
        assert np.array_equal(e_pairs[:, 0], r_ids[_m]) and np.array_equal(e_pairs[:, 1], r_sc[_m]), \
            "e2e and device-resident results differ"
        e2e = {"value": nq_total * args.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3,
               "h2d_ms_per_step": prof_e2e["h2d_ms"] / args.steps, "d2h_ms_per_step": prof_e2e["d2h_ms"] / args.steps,
               "kernel_ms_per_step": (prof_e2e["prepare_ms"] + prof_e2e["sketch_ms"] + prof_e2e["search_ms"] +
                                      prof_e2e["wide_ms"]) / args.steps}

    # ---- roofline of the dominant kernel (its launches of the timed region)
    peak, peak_src = measured_peak()
    steps = args.steps
    rows = prof["unique_terms"] / steps        # upper bound of row descriptors read (present terms <= unique terms)
    postings = prof["postings"] / steps
    results = prof["results"] / steps
    search_bytes = 8.0 * rows + 4.0 * postings + 8.0 * results + 4.0 * nq
    path_bytes = 20.0 * rows + 4.0 * postings + 8.0 * results          # SURVEY.md §8d per-query formula
    sketch_share = prof["sketch_queries"] / max(1, prof["queries"])
    if sketch_share >= 0.5:   # the TMA/sketch kernel answers (nearly) all queries of this workload
        kernel_name = "search_find_kernel (TMA bulk row gather + u8 count sketch + exact key-range resolve + top-k)"
        search_ms = prof["sketch_ms"] / steps
        search_bytes *= sketch_share
    else:
        kernel_name = "search_smem_kernel<13|14|15> (gather + exact count table + top-k)"
        search_ms = prof["search_ms"] / steps
    achieved = search_bytes / (search_ms * 1e-3) / 1e9 if search_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_%s.json" % wl)
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel_name,
                "kernel_ms_per_step": search_ms, "algorithmic_bytes_per_step": search_bytes,
                "sketch_ms_per_step": prof["sketch_ms"] / steps, "exact_ms_per_step": prof["search_ms"] / steps,
                "sketch_queries_per_step": prof["sketch_queries"] / steps,
                "overflow_requeues_per_step": prof["overflow_requeues"] / steps,
                "whole_path_bytes_per_step": path_bytes, "postings_per_step": postings,
                "whole_path_frac": path_bytes / (ms_per_step * 1e-3) / 1e9 / peak if not sharded else None,
                "prepare_ms_per_step": prof["prepare_ms"] / steps, "wide_ms_per_step": prof["wide_ms"] / steps,
                "wide_queries_per_step": prof["wide_queries"] / steps, "peak_source": peak_src}

    if sharded:
        par = "docid-range sharded x%d, whole batch on every rank, packed NCCL all-gather + device merge" % world
        launches = 11
    elif wl in STRONG:
        par = "replicated corpus, the batch split over %d GPU(s)" % world
        launches = 7
    else:
        par = "replicated corpus, one batch per GPU x%d" % world
        launches = 7
    out = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": scaling_of(wl, mode), "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(wl), "parallelism": par, "mode": mode,
        "clocks": clocks,
        "gpu_launches": launches * args.steps,  # prepare, prepare_long, hot kernel, 3 exact classes, wide (+ pack x3, merge)
        "host_affinity": numa,
        "roofline": roofline,
    }
    if e2e is not None:
        out["e2e"] = e2e
    if sharded:
        out["exchange_bytes_per_rank_per_step"] = int(ss.last_words) * 4

    # ---- CPU baseline beside it (rank 0, N=1 only) and the bit-exact check against the oracle (rank 0)
    if want_oracle:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from _oracle import OracleIndex, build as build_oracle
        build_oracle()
        orc = oracle_with(segments, OracleIndex)
        threads = os.cpu_count() or 1
        n_chk = args.parity_queries
        if world == 1 and not args.no_cpu_baseline:
            qps, sample = cpu_baseline(orc, terms.reshape(-1), offs, opts, threads)
            out["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                                   "sample": "%d queries of the same batch" % sample}
            n_chk = n_chk or min(sample, 2000)
        n_chk = min(n_chk, nq)
        out["parity_checked_queries"] = n_chk
        out["parity_bit_exact"] = check_against_oracle(orc, terms.reshape(-1), offs, opts, n_chk, threads, r_ids, r_sc, r_cnt)
        out["parity_compares"] = "counts, ids and scores"
    if rank == 0:
        print(json.dumps(out), flush=True)
    snap.release()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
