"""N > 1 host logic on CPU: two gloo ranks, docid-range sharded snapshots compiled with FPX_FLAG_HOST_ONLY,
local top-k from the compiled CSR, all-gather, exact merge through the C ABI — against the oracle's answer."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    import torch.distributed as dist
    from _helpers import csr_rank, pkg, segments_from_oracle
    from _oracle import OracleIndex
    import importlib
    mg = importlib.import_module("acoustid_index_b200.multi_gpu")

    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)           # same corpus on every rank (replicated input, sharded snapshot)
    ix = OracleIndex()
    ids_all = []
    for r in range(6):
        ch = []
        for i in range(150):
            did = r * 150 + i + 1
            ids_all.append(did)
            ch.append(("insert", did, rng.integers(0, 600, size=12).tolist()))
        if r >= 2:
            ch.append(("delete", int(rng.integers(1, r * 150))))
            ch.append(("insert", int(rng.integers(1, r * 150)), rng.integers(0, 600, size=12).tolist()))
        ix.update(ch)
        if r % 2 == 1:
            ix.checkpoint()
    queries = [rng.integers(0, 600, size=int(rng.integers(1, 30))).tolist() for _ in range(120)]
    k = 16
    opts = np.tile(np.array((k, 1, 30), dtype=np.uint32), (len(queries), 1))

    lo, hi = mg.doc_ranges(min(ids_all), max(ids_all), world)[rank]
    ctx = pkg.Context(host_only=True, host_threads=2)
    files, mems = segments_from_oracle(ix)
    b = pkg.SnapshotBuilder(ctx)
    for s in files:
        b.add_file_segment(s)
    for s in mems:
        b.add_memory_segment(s)
    b.set_doc_range(lo, hi)
    terms, offs, docids = b.csr()
    b.abort()
    assert len(docids) == 0 or ((docids >= lo).all() and (hi == 0 or (docids < hi).all()))   # hi == 0: open-ended last shard
    # local top-k with the absolute floor only (this is what a GPU shard returns with min_score_pct = 0)
    l_ids = np.zeros((len(queries), k), np.uint32)
    l_sc = np.zeros((len(queries), k), np.uint32)
    l_cnt = np.zeros(len(queries), np.uint32)
    for qi, q in enumerate(queries):
        res = csr_rank(terms, offs, docids, q, k, 1, 0)
        l_cnt[qi] = len(res)
        for j, (d, s) in enumerate(res):
            l_ids[qi, j], l_sc[qi, j] = d, s
    g_ids, g_sc, g_cnt = mg.all_gather_results(l_ids, l_sc, l_cnt)
    m_ids, m_sc, m_cnt = pkg.merge_shard_results(g_ids, g_sc, g_cnt, opts, k)
    ok = True
    for qi, q in enumerate(queries):
        want = ix.search(q, k, 1, 30)
        got = [(int(m_ids[qi, j]), int(m_sc[qi, j])) for j in range(int(m_cnt[qi]))]
        ok = ok and got == want
    lo_q, hi_q = mg.query_slice(len(queries), rank, world)
    out_q.put((rank, ok, int(g_cnt.sum()), (lo_q, hi_q)))
    ctx.close()
    dist.destroy_process_group()


def test_two_rank_sharded_merge_matches_oracle():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res), res
    assert res[0][2] == res[1][2] > 0                    # both ranks saw the same gathered candidate lists
    assert res[0][3] == (0, 60) and res[1][3] == (60, 120)  # replicated mode: contiguous query slices
