"""GPU parity: the CUDA path (through the C ABI) vs the CPU oracle on identical inputs, bit-exact.
Covers BASELINE.json configs C1, C2 (full size) and a Zipf stress, plus the edge cases the reference tests."""
import numpy as np
import pytest

from _helpers import flat_queries, have_gpu, pkg, segments_from_oracle
from _oracle import OracleIndex

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    if not have_gpu():
        pytest.skip("no CUDA device")
    c = pkg.Context(device=0, profile=True)
    yield c
    c.close()


def _snapshot_of(ctx, ix, doc_range=None):
    files, mems = segments_from_oracle(ix)
    return pkg.swap_snapshot(ctx, files, mems, doc_range=doc_range)


def _compare_batch(reader, ix, terms, offs, opts, k_stride, threads=8):
    ids, sc, cnt = reader.search_batch(terms, offs, opts, k_stride)
    oi, os_, oc, _ = ix.search_batch(terms, offs, opts, k_stride, n_threads=threads)
    bad = np.nonzero(cnt != oc)[0]
    assert len(bad) == 0, "count mismatch at queries %s: gpu %s oracle %s" % (bad[:5], cnt[bad[:5]], oc[bad[:5]])
    mask = np.arange(k_stride)[None, :] < cnt[:, None]
    for name, got, want in (("ids", ids, oi), ("scores", sc, os_)):
        diff = np.nonzero((np.where(mask, got, 0) != np.where(mask, want, 0)).any(axis=1))[0]
        assert len(diff) == 0, "%s differ at %d queries, first %d: gpu ids %s scores %s | oracle ids %s scores %s" % (
            name, len(diff), diff[0], ids[diff[0], :cnt[diff[0]]][:8], sc[diff[0], :cnt[diff[0]]][:8],
            oi[diff[0], :cnt[diff[0]]][:8], os_[diff[0], :cnt[diff[0]]][:8])
    return ids, sc, cnt


def _build_c1():
    """C1: 1K fingerprints x 50 hashes: 1 file segment (first 800 docs) + 2 memory segments with
    20 re-inserts and 10 deletes."""
    cfg = pkg.synth.SynthConfig(n_docs=1000, hashes_per_doc=50, vocab_log2=10, seed=0xF1D00001 + 1)
    syn = pkg.synth.Synth(cfg)
    fps = syn.doc_hashes(np.arange(1000))
    ix = OracleIndex()
    ix.update([("insert", i + 1, fps[i].tolist()) for i in range(800)])
    ix.checkpoint()
    ix.update([("insert", i + 1, fps[i].tolist()) for i in range(800, 900)] +
              [("insert", i + 1, fps[999 - i].tolist()) for i in range(20)])          # 20 re-inserts
    ix.update([("insert", i + 1, fps[i].tolist()) for i in range(900, 1000)] +
              [("delete", i + 1) for i in range(100, 110)])                          # 10 deletes
    return ix, syn, fps


def test_c1_single_term_queries(ctx):
    ix, syn, fps = _build_c1()
    snap = _snapshot_of(ctx, ix)
    reader = pkg.IndexReader(snap)
    terms, _ = syn.queries(100, 1, single_term=True)
    offs = np.arange(101, dtype=np.uint64)
    opts = pkg.synth.http_opts(100, 1)
    _, _, cnt = _compare_batch(reader, ix, terms.reshape(-1), offs, opts, 40)
    assert cnt.sum() > 0
    # reference-interface mirror, one query at a time, HTTP option mapping
    for q in range(10):
        got = pkg.multi_index_search(reader, pkg.SearchRequest([int(terms[q, 0])]))
        assert [tuple(r) for r in got] == ix.search_http([int(terms[q, 0])])
    # full fingerprints as queries: updated / deleted docs must follow the supersession rules
    for d in (0, 5, 19, 100, 109, 500, 850, 999):
        q = fps[d].tolist()
        got = pkg.multi_index_search(reader, pkg.SearchRequest(q))
        assert [tuple(r) for r in got] == ix.search_http(q)
    snap.release()


def test_reference_kats_through_the_gpu(ctx):
    """The reference's own e2e vectors (tests/test_fingerprint_api.py:5-52, 102-260; Index.zig:1056-1096)."""
    ix = OracleIndex()
    ix.update([("insert", 1, [101, 201, 301]), ("insert", 2, [102, 202, 302])])
    r = pkg.IndexReader(_snapshot_of(ctx, ix))
    assert pkg.multi_index_search(r, pkg.SearchRequest([101, 201, 301])) == [(1, 3)]
    assert pkg.multi_index_search(r, pkg.SearchRequest([101, 201, 301, 102, 202, 302])) == [(1, 3), (2, 3)]
    assert r.search([101, 101], pkg.SearchOptions(10, 1, 10)) == [(1, 1)]          # duplicate query hashes
    ix.checkpoint()
    r = pkg.IndexReader(_snapshot_of(ctx, ix))
    assert r.search([101, 101], pkg.SearchOptions(10, 1, 10)) == [(1, 1)]
    ix.update([("insert", 1, [101, 201, 999])])                                      # partial update
    r = pkg.IndexReader(_snapshot_of(ctx, ix))
    assert pkg.multi_index_search(r, pkg.SearchRequest([101, 201, 301])) == [(1, 2)]
    assert pkg.multi_index_search(r, pkg.SearchRequest([101, 201, 999])) == [(1, 3)]
    ix.update([("delete", 1), ("delete", 2)])
    r = pkg.IndexReader(_snapshot_of(ctx, ix))
    assert pkg.multi_index_search(r, pkg.SearchRequest([101, 201, 301, 102, 202, 302])) == []
    assert r.search([], pkg.SearchOptions(10, 0, 10)) == []
    ix2 = OracleIndex()
    ix2.update([("insert", 1001, [11000, 12000, 13000]), ("insert", 1002, [11000, 12000, 19000])])
    r = pkg.IndexReader(_snapshot_of(ctx, ix2))                                      # tests/test_legacy.py:61-69
    assert r.search([11000, 12000, 13000], pkg.SearchOptions(500, 1, 10)) == [(1001, 3), (1002, 2)]
    assert r.search([11000, 12000, 19000], pkg.SearchOptions(500, 1, 10)) == [(1002, 3), (1001, 2)]


def _random_index(rng, rounds=9, n_docs=400, H=30, vocab=4000, hot=(7, 8)):
    ix = OracleIndex()
    next_id, all_ids = 1, []
    for r in range(rounds):
        ch = []
        for _ in range(n_docs):
            roll = rng.random()
            if all_ids and roll < 0.1:
                did = int(rng.choice(all_ids))
            elif all_ids and roll < 0.15:
                ch.append(("delete", int(rng.choice(all_ids))))
                continue
            else:
                did = next_id
                next_id += 1
                all_ids.append(did)
            hs = rng.integers(0, vocab, size=H)
            hs[rng.random(H) < 0.2] = rng.choice(hot)
            hs = hs.tolist()
            if rng.random() < 0.3:
                hs[1] = hs[0]
            ch.append(("insert", did, hs))
        ix.update(ch)
        if r % 3 == 2:
            ix.checkpoint()
    return ix, all_ids


def test_random_multi_segment_varied_options(ctx):
    rng = np.random.default_rng(42)
    ix, _ = _random_index(rng)
    assert ix.num_file_segments == 3 and ix.num_memory_segments == 0
    ix.update([("insert", 5, [1, 2, 3])])
    snap = _snapshot_of(ctx, ix)
    info = snap.info()
    assert info["n_dropped_superseded"] > 0 and info["n_dropped_unreachable"] > 0
    reader = pkg.IndexReader(snap)
    queries = [rng.integers(0, 4000, size=int(rng.integers(0, 120))).tolist() for _ in range(600)]
    queries += [[7, 8] + rng.integers(0, 4000, size=30).tolist() for _ in range(100)]
    terms, offs = flat_queries(queries)
    nq = len(queries)
    for k_stride, opt in ((40, (40, 1, 10)), (64, (10, 2, 0)), (8, (100, 1, 100)), (600, (500, 1, 10)),
                          (1024, (1024, 0, 3)), (40, (0, 1, 10)), (40, (40, 3, 250)), (1, (5, 1, 10))):
        opts = np.tile(np.array(opt, dtype=np.uint32), (nq, 1))
        _compare_batch(reader, ix, terms, offs, opts, k_stride)
    # per-query mixed options in one batch
    opts = np.stack([rng.integers(0, 60, nq), rng.integers(0, 6, nq), rng.integers(0, 120, nq)], 1).astype(np.uint32)
    _compare_batch(reader, ix, terms, offs, opts, 64)
    snap.release()


def test_count_overflow_and_duplicates_fall_back_exactly(ctx):
    """Docs with hundreds of duplicate hashes push per-doc counts past the packed table's counter: the
    shared-memory path must notice and the global-memory path must still give the reference's answer."""
    ix = OracleIndex()
    ch = [("insert", 1, [5] * 700 + [6] * 300), ("insert", 2, [5] * 600), ("insert", 3, [5, 6, 7])]
    ch += [("insert", 10 + i, [5, 1000 + i]) for i in range(300)]
    ix.update(ch)
    ix.update([("insert", 4, [6] * 2000)])
    for flush in (False, True):
        if flush:
            ix.checkpoint()
        snap = _snapshot_of(ctx, ix)
        reader = pkg.IndexReader(snap)
        queries = [[5], [5, 6], [5, 6, 7], [6], [7] + list(range(1000, 1100))]
        terms, offs = flat_queries(queries)
        for opt in ((40, 1, 10), (500, 1, 0), (3, 2, 0)):
            opts = np.tile(np.array(opt, dtype=np.uint32), (len(queries), 1))
            _compare_batch(reader, ix, terms, offs, opts, 512)
        ctx.profile_reset()
        reader.search([5, 6], pkg.SearchOptions(40, 1, 10))
        assert ctx.profile()["wide_queries"] >= 1
        snap.release()


def test_sketch_counter_saturation_hands_over_exactly(ctx):
    """The sketch kernel counts in 8-bit counters.  A docid that arrives more than 128 times (here: every hash
    stored twice, 100 query terms -> 200 arrivals; and three docs of 60 arrivals that may share nothing) must
    send the query to the exact path, with the reference's scores; min_score 128 / 129 sit on the class limit."""
    ix = OracleIndex()
    hs = list(range(5000, 5100))
    ch = [("insert", 1, hs + hs), ("insert", 2, hs[:60]), ("insert", 3, hs[40:]), ("insert", 4, hs[::2] + hs[::2])]
    ch += [("insert", 100 + i, [hs[i % 100], 9000 + i]) for i in range(400)]
    ix.update(ch)
    ix.checkpoint()
    snap = _snapshot_of(ctx, ix)
    reader = pkg.IndexReader(snap)
    queries = [hs, hs[:70], hs[30:], hs + [9000 + i for i in range(20)]]
    terms, offs = flat_queries(queries)
    ctx.profile_reset()
    for opt in ((40, 5, 10), (40, 2, 0), (40, 100, 0), (40, 128, 0), (40, 129, 0), (40, 200, 0), (40, 201, 0)):
        opts = np.tile(np.array(opt, dtype=np.uint32), (len(queries), 1))
        _compare_batch(reader, ix, terms, offs, opts, 40)
    prof = ctx.profile()
    assert prof["overflow_requeues"] >= 1, prof
    snap.release()


def test_long_queries(ctx):
    rng = np.random.default_rng(3)
    ix, _ = _random_index(rng, rounds=6, vocab=20000, hot=(7,))
    snap = _snapshot_of(ctx, ix)
    reader = pkg.IndexReader(snap)
    queries = [rng.integers(0, 20000, size=n).tolist() for n in (128, 129, 130, 255, 256, 257, 1000, 4096, 8192, 1, 0)]
    queries.append([3] * 500 + [4] * 500)            # heavy duplication in a long query
    terms, offs = flat_queries(queries)
    opts = np.array([[40, pkg.lib().fpx_default_min_score(len(q)), 10] for q in queries], dtype=np.uint32)
    _compare_batch(reader, ix, terms, offs, opts, 40)
    opts[:, 1] = 1
    _compare_batch(reader, ix, terms, offs, opts, 40)
    with pytest.raises(pkg.FpxError) as e:
        reader.search(list(range(8193)), pkg.SearchOptions())
    assert e.value.status == pkg._ffi.FPX_UNSUPPORTED
    snap.release()


def test_doc_range_shards_merge_to_the_unsharded_answer(ctx):
    rng = np.random.default_rng(9)
    ix, ids = _random_index(rng, rounds=6)
    queries = [rng.integers(0, 4000, size=40).tolist() for _ in range(200)]
    terms, offs = flat_queries(queries)
    nq, k = len(queries), 40
    opts = np.tile(np.array((40, 1, 10), dtype=np.uint32), (nq, 1))
    whole = pkg.IndexReader(_snapshot_of(ctx, ix))
    want = whole.search_batch(terms, offs, opts, k)
    cuts = [0, int(np.quantile(ids, 0.3)), int(np.quantile(ids, 0.7)), 0xFFFFFFFF]
    shard_opts = opts.copy()
    shard_opts[:, 2] = 0                                 # absolute floor only on the shards
    parts = [pkg.IndexReader(_snapshot_of(ctx, ix, (cuts[i], cuts[i + 1]))).search_batch(terms, offs, shard_opts, k)
             for i in range(3)]
    got = pkg.merge_shard_results(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]),
                                  np.stack([p[2] for p in parts]), opts, k)
    assert np.array_equal(got[2], want[2])
    mask = np.arange(k)[None, :] < want[2][:, None]
    assert np.array_equal(np.where(mask, got[0], 0), np.where(mask, want[0], 0))
    assert np.array_equal(np.where(mask, got[1], 0), np.where(mask, want[1], 0))
    _compare_batch(whole, ix, terms, offs, opts, k)


@pytest.fixture(scope="module")
def c2(ctx):
    """C2: 1M fingerprints x 100 hashes, one fully merged file segment; 10K queries x 100 terms."""
    cfg = pkg.synth.SynthConfig(n_docs=1_000_000, hashes_per_doc=100, vocab_log2=20, seed=0xF1D00001 + 2)
    syn = pkg.synth.Synth(cfg, device="cuda:0")
    items, doc_ids, doc_alive = syn.corpus_items()
    seg = pkg.FileSegment.from_items(items, doc_ids, doc_alive, commit_id=1)
    del items
    snap = pkg.swap_snapshot(ctx, [seg])
    ix = OracleIndex()
    ix.adopt_file_segment(1, 0, seg.block_size, seg.blocks, seg.num_blocks, seg.block_index, seg.doc_ids, seg.doc_alive)
    yield syn, seg, snap, ix
    snap.release()


def test_c2_full_size_bit_exact(ctx, c2):
    syn, seg, snap, ix = c2
    info = snap.info()
    assert info["n_postings_total"] == 100_000_000
    reader = pkg.IndexReader(snap)
    terms, src = syn.queries(10_000, 100, seed=0xF1D01001 + 2)
    nq, T = terms.shape
    offs = np.arange(nq + 1, dtype=np.uint64) * T
    opts = pkg.synth.http_opts(nq, T)
    ids, sc, cnt = _compare_batch(reader, ix, terms.reshape(-1), offs, opts, 40, threads=16)
    # domain property: a noisy copy of doc d finds d first
    hit = src >= 0
    assert (cnt[hit] >= 1).all()
    assert (ids[hit, 0] == (src[hit] + 1)).mean() > 0.99
    # legacy options (limit 500, min_score 1) on a slice: exercises the many-candidates path
    opts2 = np.tile(np.array((500, 1, 10), dtype=np.uint32), (500, 1))
    _compare_batch(reader, ix, terms[:500].reshape(-1), offs[:501], opts2, 500, threads=16)


def test_c2_size_independent_properties(ctx, c2):
    syn, seg, snap, ix = c2
    reader = pkg.IndexReader(snap)
    rng = np.random.default_rng(1)
    terms, _ = syn.queries(4000, 100, seed=77)
    nq, T = terms.shape
    offs = np.arange(nq + 1, dtype=np.uint64) * T
    opts = pkg.synth.http_opts(nq, T)
    a = reader.search_batch(terms.reshape(-1), offs, opts, 40)
    b = reader.search_batch(terms.reshape(-1), offs, opts, 40)                       # idempotence
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    perm = np.stack([rng.permutation(T) for _ in range(nq)])                          # the query is a set:
    shuffled = np.take_along_axis(terms, perm, 1)                                    # order and repeats do not matter
    doubled = np.concatenate([shuffled, terms[:, :37]], 1)
    offs2 = np.arange(nq + 1, dtype=np.uint64) * doubled.shape[1]
    c = reader.search_batch(doubled.reshape(-1), offs2, opts, 40)
    assert all(np.array_equal(x, y) for x, y in zip(a, c))
    ids, sc, cnt = a
    mask = np.arange(40)[None, :] < cnt[:, None]
    s = np.where(mask, sc, 0).astype(np.int64)
    assert (np.diff(s, axis=1) <= 0).all()                                           # score descending
    tie = mask[:, 1:] & (sc[:, 1:] == sc[:, :-1])
    assert (ids[:, 1:][tie] > ids[:, :-1][tie]).all()                                # id ascending on ties
    floor = np.maximum(opts[:, 1], (sc[:, 0] * opts[:, 2]) // 100)
    assert (np.where(mask, sc, 1 << 30) >= floor[:, None]).all()                     # both cutoffs hold
    # splitting the batch (other chunking, other workspaces) changes nothing
    h = nq // 3
    p1 = reader.search_batch(terms[:h].reshape(-1), offs[:h + 1], opts[:h], 40)
    p2 = reader.search_batch(terms[h:].reshape(-1), offs[:nq - h + 1], opts[h:], 40)
    assert np.array_equal(np.concatenate([p1[2], p2[2]]), cnt)
    assert np.array_equal(np.concatenate([p1[0], p2[0]])[mask], ids[mask])


def test_zipf_hot_postings(ctx):
    """C5-shaped stress at test size: Zipf vocabulary, rows cut by the scan caps, multi-pass queries."""
    cfg = pkg.synth.SynthConfig(n_docs=300_000, hashes_per_doc=60, vocab_log2=14, seed=0xF1D00001 + 5, zipf_s=1.0)
    syn = pkg.synth.Synth(cfg, device="cuda:0")
    items, doc_ids, doc_alive = syn.corpus_items()
    seg = pkg.FileSegment.from_items(items, doc_ids, doc_alive, commit_id=1)
    snap = pkg.swap_snapshot(ctx, [seg])
    info = snap.info()
    assert info["n_dropped_unreachable"] > 0 and info["max_row_len"] > 1000
    ix = OracleIndex()
    ix.adopt_file_segment(1, 0, seg.block_size, seg.blocks, seg.num_blocks, seg.block_index, seg.doc_ids, seg.doc_alive)
    reader = pkg.IndexReader(snap)
    terms, _ = syn.queries(3000, 100, seed=5)
    nq, T = terms.shape
    offs = np.arange(nq + 1, dtype=np.uint64) * T
    ctx.profile_reset()
    _compare_batch(reader, ix, terms.reshape(-1), offs, pkg.synth.http_opts(nq, T), 40, threads=16)
    opts = np.tile(np.array((100, 2, 0), dtype=np.uint32), (nq, 1))
    _compare_batch(reader, ix, terms.reshape(-1), offs, opts, 100, threads=16)
    snap.release()


def test_sketch_and_exact_kernels_agree(ctx, c2):
    """The TMA/sketch kernel and the exact count-table kernels must give identical answers."""
    syn, seg, snap, ix = c2
    terms, _ = syn.queries(3000, 60, seed=321)     # 60 x ~95 postings fits the 32 KB stage of the sketch path
    nq, T = terms.shape
    offs = np.arange(nq + 1, dtype=np.uint64) * T
    ctx2 = pkg.Context(device=0, profile=True, no_sketch=True)
    snap2 = pkg.swap_snapshot(ctx2, [seg])
    for opt in ((40, 5, 10), (40, 2, 0), (100, 3, 50), (512, 2, 10)):
        opts = np.tile(np.array(opt, dtype=np.uint32), (nq, 1))
        ctx.profile_reset()
        a = pkg.IndexReader(snap).search_batch(terms.reshape(-1), offs, opts, opt[0])
        prof = ctx.profile()
        if opt[1] >= 4:
            assert prof["sketch_queries"] > 0.9 * nq, prof
        ctx2.profile_reset()
        b = pkg.IndexReader(snap2).search_batch(terms.reshape(-1), offs, opts, opt[0])
        assert ctx2.profile()["sketch_queries"] == 0
        assert np.array_equal(a[2], b[2])
        mask = np.arange(opt[0])[None, :] < a[2][:, None]
        assert np.array_equal(a[0][mask], b[0][mask]) and np.array_equal(a[1][mask], b[1][mask])
    snap2.release()
    ctx2.close()


def test_device_resident_api_matches_host_api(ctx, c2):
    import torch
    syn, seg, snap, ix = c2
    reader = pkg.IndexReader(snap)
    terms, _ = syn.queries(3000, 100, seed=123)
    nq, T = terms.shape
    offs = np.arange(nq + 1, dtype=np.uint64) * T
    opts = pkg.synth.http_opts(nq, T)
    want = reader.search_batch(terms.reshape(-1), offs, opts, 40)
    dev = torch.device("cuda:0")
    d_terms = torch.from_numpy(terms.reshape(-1).view(np.int32)).to(dev)
    d_offs = torch.from_numpy(offs.view(np.int64)).to(dev)
    d_opts = torch.from_numpy(opts.view(np.int32)).to(dev)
    d_ids = torch.zeros((nq, 40), dtype=torch.int32, device=dev)
    d_sc = torch.zeros((nq, 40), dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    reader.search_batch_device(nq, d_terms.data_ptr(), d_offs.data_ptr(), d_opts.data_ptr(), 40, d_ids.data_ptr(),
                               d_sc.data_ptr(), d_cnt.data_ptr(), st.cuda_stream)
    st.synchronize()
    cnt = d_cnt.cpu().numpy().view(np.uint32)
    assert np.array_equal(cnt, want[2])
    mask = np.arange(40)[None, :] < cnt[:, None]
    assert np.array_equal(d_ids.cpu().numpy().view(np.uint32)[mask], want[0][mask])
    assert np.array_equal(d_sc.cpu().numpy().view(np.uint32)[mask], want[1][mask])


def test_pinned_dma_and_pageable_packed_paths_agree(ctx, c2):
    """fpx_search_batch returns results by DMA into pinned caller memory, or — for pageable memory — packed by the
    GPU into the library's pinned buffers and scattered by the host.  Both must give the same arrays, also when the
    batch is cut into many chunks (three streams, rotating workspace slots)."""
    import torch
    syn, seg, snap, ix = c2
    reader = pkg.IndexReader(snap)
    terms, _ = syn.queries(6000, 60, seed=777)
    nq, T = terms.shape
    offs = np.arange(nq + 1, dtype=np.uint64) * T
    opts = np.tile(np.array((40, 3, 10), dtype=np.uint32), (nq, 1))
    want = _compare_batch(reader, ix, terms.reshape(-1), offs, opts, 40, threads=16)     # pageable numpy: packed path
    h = [torch.from_numpy(terms.reshape(-1).view(np.int32).copy()).pin_memory(),
         torch.from_numpy(offs.view(np.int64).copy()).pin_memory(),
         torch.from_numpy(opts.view(np.int32).copy()).pin_memory()]
    for chunk in (1024, 8192):       # schedules of 1024.. and 1024-sized chunks: > 3 chunks in flight
        ctx.set_chunk_queries(chunk)
        o = [torch.full((nq, 40), -1, dtype=torch.int32).pin_memory(), torch.full((nq, 40), -1, dtype=torch.int32).pin_memory(),
             torch.full((nq,), -1, dtype=torch.int32).pin_memory()]
        reader.search_batch_ptr(nq, h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), 40, o[0].data_ptr(),
                                o[1].data_ptr(), o[2].data_ptr())                        # pinned: DMA path
        cnt = o[2].numpy().view(np.uint32)
        assert np.array_equal(cnt, want[2])
        mask = np.arange(40)[None, :] < cnt[:, None]
        assert np.array_equal(o[0].numpy().view(np.uint32)[mask], want[0][mask])
        assert np.array_equal(o[1].numpy().view(np.uint32)[mask], want[1][mask])
        got = reader.search_batch(terms.reshape(-1), offs, opts, 40)                    # packed path, same chunking
        assert np.array_equal(got[2], want[2]) and np.array_equal(got[0][mask], want[0][mask]) \
            and np.array_equal(got[1][mask], want[1][mask])
    ctx.set_chunk_queries(131072)


def test_async_device_api_queues_batches_without_a_sync(ctx, c2):
    """fpx_search_batch_device_async: 20 batches enqueued back to back on one stream with no host synchronisation in
    between (each into its own output arrays), then one sync: every batch must carry the oracle's answers.  The status
    word reports what the kernels raised: 0 for good batches, FPX_UNSUPPORTED for a batch holding a query with more than
    FPX_MAX_QUERY_TERMS terms (count 0 for that query, the others answered), FPX_INVALID_ARGUMENT for offsets outside
    the stated window."""
    import torch
    syn, seg, snap, ix = c2
    reader = pkg.IndexReader(snap)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    nb, nq, T = 20, 700, 100
    terms, _ = syn.queries(nb * nq, T, seed=2468)
    opts = pkg.synth.http_opts(nb * nq, T)
    d_terms = torch.from_numpy(terms.reshape(-1).view(np.int32)).to(dev)
    d_offs = torch.from_numpy((np.arange(nb * nq + 1, dtype=np.uint64) * T).view(np.int64)).to(dev)
    d_opts = torch.from_numpy(opts.view(np.int32)).to(dev)
    outs = [(torch.zeros((nq, 40), dtype=torch.int32, device=dev), torch.zeros((nq, 40), dtype=torch.int32, device=dev),
             torch.zeros(nq, dtype=torch.int32, device=dev), torch.full((1,), 99, dtype=torch.int32, device=dev)) for _ in range(nb)]
    torch.cuda.synchronize()
    for b in range(nb):                                  # the offsets of batch b start at element b * nq of d_offs
        o = outs[b]
        reader.search_batch_device_async(nq, b * nq * T, nq * T, d_terms.data_ptr(), d_offs.data_ptr() + 8 * b * nq,
                                         d_opts.data_ptr() + 12 * b * nq, 40, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                                         o[3].data_ptr(), stream.cuda_stream)
    stream.synchronize()
    oi, os_, oc, _ = ix.search_batch(terms.reshape(-1), np.arange(nb * nq + 1, dtype=np.uint64) * T, opts, 40, n_threads=16)
    for b in range(nb):
        ids, sc, cnt, status = [x.cpu().numpy().view(np.uint32) for x in outs[b]]
        assert status[0] == 0
        sl = slice(b * nq, (b + 1) * nq)
        assert np.array_equal(cnt, oc[sl])
        mask = np.arange(40)[None, :] < cnt[:, None]
        assert np.array_equal(ids[mask], oi[sl][mask]) and np.array_equal(sc[mask], os_[sl][mask])
    # a batch with one oversize query (8193 terms): status FPX_UNSUPPORTED, count 0 there, the neighbours answered
    big = np.concatenate([terms[0], np.arange(8193, dtype=np.uint32), terms[1]])
    offs2 = np.array([0, T, T + 8193, 2 * T + 8193], dtype=np.uint64)
    d_t2 = torch.from_numpy(big.view(np.int32)).to(dev)
    d_o2 = torch.from_numpy(offs2.view(np.int64)).to(dev)
    o = [torch.zeros((3, 40), dtype=torch.int32, device=dev), torch.zeros((3, 40), dtype=torch.int32, device=dev),
         torch.full((3,), 7, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)]
    reader.search_batch_device_async(3, 0, len(big), d_t2.data_ptr(), d_o2.data_ptr(), d_opts.data_ptr(), 40, o[0].data_ptr(),
                                     o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr(), stream.cuda_stream)
    stream.synchronize()
    cnt = o[2].cpu().numpy().view(np.uint32)
    assert int(o[3].item()) == pkg._ffi.FPX_UNSUPPORTED and cnt[1] == 0 and cnt[0] == oc[0] and cnt[2] == oc[1]
    # offsets that leave the stated window: FPX_INVALID_ARGUMENT on the device, nothing read out of bounds
    reader.search_batch_device_async(3, 0, 50, d_t2.data_ptr(), d_o2.data_ptr(), d_opts.data_ptr(), 40, o[0].data_ptr(),
                                     o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr(), stream.cuda_stream)
    stream.synchronize()
    assert int(o[3].item()) == pkg._ffi.FPX_INVALID_ARGUMENT and (o[2].cpu().numpy() == 0).all()


def test_batch_deadline_mirrors_search_timeout(ctx, c2):
    """fpx_search_batch_timeout: a generous deadline gives the normal answer, an impossible one FPX_TIMEOUT
    (error.SearchTimeout, MultiIndex.zig:311-322) — and the context keeps working afterwards."""
    syn, seg, snap, ix = c2
    reader = pkg.IndexReader(snap)
    terms, _ = syn.queries(40000, 100, seed=1357)
    nq, T = terms.shape
    offs = np.arange(nq + 1, dtype=np.uint64) * T
    opts = np.tile(np.array((100, 1, 0), dtype=np.uint32), (nq, 1))     # floor 1: the exact count-table path, several ms
    want = reader.search_batch(terms.reshape(-1), offs, opts, 100)
    got = reader.search_batch_timeout(terms.reshape(-1), offs, opts, 100, 60_000)
    assert np.array_equal(got[2], want[2]) and np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    with pytest.raises(pkg.FpxError) as err:
        reader.search_batch_timeout(terms.reshape(-1), offs, opts, 100, 1)
    assert err.value.status == pkg._ffi.FPX_TIMEOUT
    again = reader.search_batch(terms.reshape(-1), offs, opts, 100)
    assert np.array_equal(again[2], want[2]) and np.array_equal(again[0], want[0])


def test_packed_result_api_matches_the_strided_one(ctx, c2):
    """fpx_search_batch_packed (a list per query, as the reference returns results) against fpx_search_batch: same
    counts, same (id, score) pairs in the same order, under fine chunking too; a too small pair buffer is reported
    with the size needed and complete counts."""
    syn, seg, snap, ix = c2
    reader = pkg.IndexReader(snap)
    terms, _ = syn.queries(6000, 100, seed=5150)
    nq, T = terms.shape
    offs = np.arange(nq + 1, dtype=np.uint64) * T
    opts = pkg.synth.http_opts(nq, T)
    opts[::3] = (40, 1, 0)                              # floor 1: dozens of results for a third of the queries
    want = reader.search_batch(terms.reshape(-1), offs, opts, 40)
    mask = np.arange(40)[None, :] < want[2][:, None]
    for chunk in (131072, 1024):
        ctx.set_chunk_queries(chunk)
        cnt, pairs = reader.search_batch_packed(terms.reshape(-1), offs, opts, 40)
        assert np.array_equal(cnt, want[2]) and len(pairs) == int(want[2].sum())
        assert np.array_equal(pairs[:, 0], want[0][mask]) and np.array_equal(pairs[:, 1], want[1][mask])
    ctx.set_chunk_queries(131072)
    with pytest.raises(pkg.FpxError):
        reader.search_batch_packed(terms.reshape(-1), offs, opts, 40, capacity_pairs=10)
    e_cnt, e_pairs = reader.search_batch_packed(np.zeros(0, np.uint32), np.zeros(1, np.uint64), np.zeros((0, 3), np.uint32), 40)
    assert len(e_cnt) == 0 and len(e_pairs) == 0


def test_hot_kernel_variants_agree_with_the_oracle(ctx, c2):
    """search_find_kernel under its alternative warp splits and stage counts (FPX_DEBUG_ABLATE bits 24..27) against the
    oracle: C2 documents as queries under several option sets, and a multi-segment snapshot with duplicate hashes
    and supersession."""
    syn, seg, snap, ix = c2
    reader = pkg.IndexReader(snap)
    terms, _ = syn.queries(3000, 60, seed=4321)
    nq, T = terms.shape
    offs = np.arange(nq + 1, dtype=np.uint64) * T
    rng = np.random.default_rng(7)
    ix2, _ = _random_index(rng)
    snap2 = _snapshot_of(ctx, ix2)
    queries = [rng.integers(0, 4000, size=int(rng.integers(0, 120))).tolist() for _ in range(500)]
    t2, o2 = flat_queries(queries)
    try:
        for variant in (1 << 24, 2 << 24, 3 << 24):
            ctx.debug_set(variant)
            for opt in ((40, 5, 10), (40, 2, 0), (100, 3, 50), (512, 7, 10), (3, 4, 100)):
                opts = np.tile(np.array(opt, dtype=np.uint32), (nq, 1))
                ctx.profile_reset()
                _compare_batch(reader, ix, terms.reshape(-1), offs, opts, min(opt[0], 64), threads=16)
                if opt[1] >= 5:
                    assert ctx.profile()["sketch_queries"] > 0.9 * nq
            for opt in ((40, 2, 10), (40, 3, 0), (64, 5, 10)):
                _compare_batch(pkg.IndexReader(snap2), ix2, t2, o2, np.tile(np.array(opt, dtype=np.uint32), (len(queries), 1)), 64)
    finally:
        ctx.debug_set(0)
        snap2.release()


def test_c2_queries_stay_on_the_sketch_path(ctx, c2):
    """100-term C2 queries (~9.7 K padded postings, 39 KB) fit a 40832-byte stage of the hot kernel and are answered
    there, not by the exact count-table kernels."""
    syn, seg, snap, ix = c2
    terms, _ = syn.queries(4000, 100, seed=99)
    nq, T = terms.shape
    offs = np.arange(nq + 1, dtype=np.uint64) * T
    opts = pkg.synth.http_opts(nq, T)
    ctx.profile_reset()
    _compare_batch(pkg.IndexReader(snap), ix, terms.reshape(-1), offs, opts, 40, threads=16)
    prof = ctx.profile()
    assert prof["sketch_queries"] > 0.95 * nq, prof


def test_sketch_admission_keeps_requeues_rare(ctx, c2):
    """The admission limits of the sketch path (csrc/fpx_kernels.cu make_item: postings <= 512 / 2900 / 7600 for
    min_score 2 / 3 / 4) are derived for <= 4 expected chance-hot counters per query; kHotCap is 32.  Measured here:
    queries right below each limit must be answered by the sketch kernel with (almost) no re-queues."""
    syn, seg, snap, ix = c2
    reader = pkg.IndexReader(snap)
    for ms, T in ((2, 5), (3, 29), (4, 76), (5, 100)):    # ~95 postings per term
        terms, _ = syn.queries(2000, T, seed=1000 + ms)
        nq = terms.shape[0]
        offs = np.arange(nq + 1, dtype=np.uint64) * T
        opts = np.tile(np.array((40, ms, 0), dtype=np.uint32), (nq, 1))
        ctx.profile_reset()
        _compare_batch(reader, ix, terms.reshape(-1), offs, opts, 40, threads=16)
        prof = ctx.profile()
        assert prof["sketch_queries"] + prof["overflow_requeues"] > 0.8 * nq, (ms, prof)
        assert prof["overflow_requeues"] <= 0.02 * nq, (ms, prof)


def test_doc_range_shards_merge_on_the_device(ctx, c2):
    """The sharded mode's device side: three docid-range shards of C2 searched with the absolute floor only, their
    results packed (fpx_pack_results_device), the three blocks laid out as an all-gather delivers them (cut to the
    largest block), fpx_merge_packed_shards_device -> must equal the unsharded answer and the oracle; several
    option sets incl. limit 3 (the global top-k cuts across shards) and a 100 % relative cutoff."""
    import torch
    syn, seg, snap, ix = c2
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    terms, _ = syn.queries(4000, 100, seed=777)
    nq, T = terms.shape
    offs = np.arange(nq + 1, dtype=np.uint64) * T
    import importlib
    ranges = importlib.import_module("acoustid_index_b200.multi_gpu").doc_ranges(1, 1_000_000, 3)
    assert ranges[0][0] == 0 and ranges[-1][1] == 0
    shards = [pkg.swap_snapshot(ctx, [seg], doc_range=r) for r in ranges]
    assert sum(s.info()["n_postings"] for s in shards) == snap.info()["n_postings"]
    d_terms = torch.from_numpy(terms.reshape(-1).view(np.int32)).to(dev)
    d_offs = torch.from_numpy(offs.view(np.int64)).to(dev)
    for opt, k in (((40, 5, 10), 40), ((3, 2, 0), 8), ((40, 1, 100), 40), ((100, 3, 50), 64)):
        opts = np.tile(np.array(opt, dtype=np.uint32), (nq, 1))
        local = opts.copy()
        local[:, 2] = 0
        d_opts = torch.from_numpy(opts.view(np.int32)).to(dev)
        d_local = torch.from_numpy(local.view(np.int32)).to(dev)
        cap = nq * k
        blocks = []
        for sh in shards:
            ids = torch.zeros((nq, k), dtype=torch.int32, device=dev)
            sc = torch.zeros((nq, k), dtype=torch.int32, device=dev)
            cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
            pkg.IndexReader(sh).search_batch_device(nq, d_terms.data_ptr(), d_offs.data_ptr(), d_local.data_ptr(), k,
                                                    ids.data_ptr(), sc.data_ptr(), cnt.data_ptr(), st)
            packed = torch.zeros(2 * nq + 2 + 2 * cap, dtype=torch.int32, device=dev)
            pkg.pack_results_device(nq, k, ids.data_ptr(), sc.data_ptr(), cnt.data_ptr(), packed.data_ptr(), cap, st)
            blocks.append(packed)
        torch.cuda.synchronize()
        words = 2 * nq + 2 + 2 * max(int(b[2 * nq].item()) for b in blocks)     # phase 1 of the exchange
        recv = torch.cat([b[:words] for b in blocks]).contiguous()               # phase 2
        out = [torch.zeros((nq, k), dtype=torch.int32, device=dev), torch.zeros((nq, k), dtype=torch.int32, device=dev),
               torch.zeros(nq, dtype=torch.int32, device=dev)]
        pkg.merge_packed_shards_device(3, nq, recv.data_ptr(), words, d_opts.data_ptr(), k, out[0].data_ptr(),
                                       out[1].data_ptr(), out[2].data_ptr(), st)
        torch.cuda.synchronize()
        got = [o.cpu().numpy().view(np.uint32) for o in out]
        want = _compare_batch(pkg.IndexReader(snap), ix, terms.reshape(-1), offs, opts, k, threads=16)
        assert np.array_equal(got[2], want[2])
        mask = np.arange(k)[None, :] < want[2][:, None]
        assert np.array_equal(got[0][mask], want[0][mask]) and np.array_equal(got[1][mask], want[1][mask])
    for sh in shards:
        sh.release()


def test_pack_results_for_exchange(ctx, c2):
    """fpx_pack_results_device + unpack_results round-trip the k_stride-wide result arrays (multi-GPU exchange)."""
    import torch
    syn, seg, snap, ix = c2
    reader = pkg.IndexReader(snap)
    terms, _ = syn.queries(5000, 100, seed=99)
    nq, T = terms.shape
    offs = np.arange(nq + 1, dtype=np.uint64) * T
    opts = np.tile(np.array((40, 1, 0), dtype=np.uint32), (nq, 1))     # floor 1: many results per query
    opts[::2] = (40, 5, 10)
    want = reader.search_batch(terms.reshape(-1), offs, opts, 40)
    dev = torch.device("cuda:0")
    d = [torch.from_numpy(np.ascontiguousarray(x).view(np.int32)).to(dev) for x in want]
    cap = int(want[2].sum()) + 7
    packed = torch.zeros(2 * nq + 2 + 2 * cap, dtype=torch.int32, device=dev)
    pkg.pack_results_device(nq, 40, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), packed.data_ptr(), cap,
                            torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ids, sc, cnt = pkg.unpack_results(packed.cpu().numpy(), nq, 40, cap)
    mask = np.arange(40)[None, :] < want[2][:, None]
    assert np.array_equal(cnt, want[2]) and np.array_equal(ids[mask], want[0][mask]) and np.array_equal(sc[mask], want[1][mask])
    small = torch.zeros(2 * nq + 2 + 2 * 10, dtype=torch.int32, device=dev)           # too small: must say so
    pkg.pack_results_device(nq, 40, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), small.data_ptr(), 10,
                            torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    with pytest.raises(pkg.FpxError):
        pkg.unpack_results(small.cpu().numpy(), nq, 40, 10)


def test_micro_batcher_concurrent_single_queries(ctx, c2):
    """fpx_batcher: 16 threads issue single queries (the MultiIndex.search seam); every answer must equal the
    oracle's HTTP-default answer, the worker must actually batch, a snapshot swap mid-stream must not disturb
    running searches, and an already expired deadline gives FPX_TIMEOUT (error.SearchTimeout)."""
    import threading
    syn, seg, snap, ix = c2
    terms, _ = syn.queries(1500, 100, seed=555)
    want = [ix.search_http(terms[q].tolist()) for q in range(200)]
    b = pkg.Batcher(ctx, max_batch=256, max_wait_us=200)
    b.set_snapshot(snap)
    got, errs = [None] * len(terms), []

    def client(t):
        try:
            for q in range(t, len(terms), 16):
                got[q] = [tuple(r) for r in b.search(pkg.SearchRequest(terms[q].tolist(), timeout=0))]
                if q == 700:
                    b.set_snapshot(snap)      # swap (to an equal snapshot) while others are searching
        except Exception as e:                # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=client, args=(t,)) for t in range(16)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs[:1]
    assert got[:200] == want
    batch_ids, batch_sc, batch_cnt = pkg.IndexReader(snap).search_batch(terms.reshape(-1), np.arange(len(terms) + 1, dtype=np.uint64) * 100,
                                                                        pkg.synth.http_opts(len(terms), 100), 40)
    for q in range(len(terms)):
        n = int(batch_cnt[q])
        assert got[q] == list(zip(batch_ids[q, :n].tolist(), batch_sc[q, :n].tolist()))
    st = b.stats()
    assert st["queries"] == len(terms) and st["batches"] < len(terms) and st["max_batch_seen"] > 1, st
    # limits and errors travel per request
    assert b.search(pkg.SearchRequest([], timeout=0)) == []
    with pytest.raises(pkg.FpxError) as e:
        b.search(pkg.SearchRequest(list(range(9000)), timeout=0))
    assert e.value.status == 7                                                   # FPX_UNSUPPORTED
    b.set_snapshot(None)
    with pytest.raises(pkg.FpxError):
        b.search(pkg.SearchRequest([1, 2, 3], timeout=0))
    b.set_snapshot(snap)
    b.close()
    slow = pkg.Batcher(ctx, max_batch=4096, max_wait_us=300000)                  # an idle worker waits 0.3 s for company
    slow.set_snapshot(snap)
    with pytest.raises(pkg.FpxError) as e:
        slow.search(pkg.SearchRequest(terms[0].tolist(), timeout=20))
    assert e.value.status == 4 and slow.stats()["timeouts"] == 1                 # FPX_TIMEOUT
    slow.close()


def test_device_built_snapshot_equals_host_built(ctx, c2):
    """fpx_gpu_build.cu (StreamVByte block decode, scan caps, liveness, row assembly on the GPU — the default for a
    device context) against the host compiler: same statistics, same term directory, same rows, on a multi-segment
    index with updates, deletes, duplicate hashes and hot terms that hit both scan caps, with and without a docid
    range, and on C2."""
    host_ctx = pkg.Context(device=0, host_build=True)
    rng = np.random.default_rng(2024)
    ix, ids = _random_index(rng, rounds=9, n_docs=700, H=40, vocab=3000)
    ix.update([("insert", int(ids[3]), [1, 2, 3]), ("delete", int(ids[4]))])       # a memory segment on top
    files, mems = segments_from_oracle(ix)
    assert len(files) == 3 and len(mems) == 1
    for doc_range in (None, (100, 900)):
        a = pkg.swap_snapshot(ctx, files, mems, doc_range=doc_range)               # device build
        b = pkg.swap_snapshot(host_ctx, files, mems, doc_range=doc_range)          # host build
        ia, ib = a.info(), b.info()
        for k in ("n_segments", "n_terms", "n_postings", "n_postings_total", "n_dropped_unreachable",
                  "n_dropped_superseded", "n_dropped_out_of_range", "max_row_len", "pad_id", "doc_lo", "doc_hi"):
            assert ia[k] == ib[k], (k, ia[k], ib[k])
        assert ia["n_dropped_unreachable"] > 0 and ia["n_dropped_superseded"] > 0
        terms = np.arange(0, 3000, dtype=np.uint32)
        assert np.array_equal(a.row_lengths(terms), b.row_lengths(terms))
        for t in list(range(0, 3000, 37)) + [7, 8]:                               # 7, 8: the hot terms
            assert np.array_equal(a.read_row(t), b.read_row(t)), t
        a.release(), b.release()
    syn, seg, snap, _ = c2                                                          # `snap` was built on the device
    hb = pkg.swap_snapshot(host_ctx, [seg])
    ia, ib = snap.info(), hb.info()
    assert all(ia[k] == ib[k] for k in ("n_terms", "n_postings", "max_row_len", "pad_id", "device_bytes"))
    sample = syn.queries(50, 100, seed=5)[0].reshape(-1)
    assert np.array_equal(snap.row_lengths(sample), hb.row_lengths(sample))
    for t in sample[:300]:
        assert np.array_equal(snap.read_row(int(t)), hb.read_row(int(t)))
    hb.release()
    # a corrupt block is reported, not decoded
    bad = pkg.FileSegment(seg.commit_id, 0, seg.min_doc_id, seg.block_size, seg.blocks[:4 * 512].copy(), 4,
                          seg.block_index[:4].copy(), seg.doc_ids, seg.doc_alive)
    bad.blocks[512 + 6] = 0xFF
    bad.blocks[512 + 7] = 0xFF                                                       # docids_offset beyond the block
    with pytest.raises(pkg.FpxError) as e:
        pkg.swap_snapshot(ctx, [bad])
    assert e.value.status == 3
    host_ctx.close()


@pytest.fixture(scope="module")
def c3(ctx):
    """C3, the metric's configuration: 10 M fingerprints x 120 hashes, vocabulary 2^24, one merged file segment."""
    import torch
    cfg = pkg.synth.SynthConfig(n_docs=10_000_000, hashes_per_doc=120, vocab_log2=24, seed=0xF1D00001 + 3)
    syn = pkg.synth.Synth(cfg, device="cuda:0")
    items, doc_ids, doc_alive = syn.corpus_items()
    seg = pkg.FileSegment.from_items(items, doc_ids, doc_alive, commit_id=1)
    del items
    torch.cuda.empty_cache()
    snap = pkg.swap_snapshot(ctx, [seg])
    ix = OracleIndex()
    ix.adopt_file_segment(1, 0, seg.block_size, seg.blocks, seg.num_blocks, seg.block_index, seg.doc_ids, seg.doc_alive)
    yield syn, seg, snap, ix
    snap.release()


def test_c3_full_size_bit_exact(ctx, c3):
    """The headline configuration at full size: the whole 100 K-query batch through the C ABI; ids, scores and
    counts of a 6000-query sample against the oracle; size-independent properties on all of it."""
    syn, seg, snap, ix = c3
    assert snap.info()["n_postings_total"] == 1_200_000_000
    reader = pkg.IndexReader(snap)
    terms, src = syn.queries(100_000, 100, seed=0xF1D01001 + 3)
    nq, T = terms.shape
    offs = np.arange(nq + 1, dtype=np.uint64) * T
    opts = pkg.synth.http_opts(nq, T)
    ctx.profile_reset()
    ids, sc, cnt = reader.search_batch(terms.reshape(-1), offs, opts, 40)
    prof = ctx.profile()
    assert prof["sketch_queries"] > 0.99 * nq and prof["overflow_requeues"] < 0.001 * nq, prof
    # the oracle on a sample spread over the batch (every 17th query, about 6000): ids AND scores AND counts
    pick = np.arange(0, nq, 17)
    st, so = flat_queries([terms[q].tolist() for q in pick])
    oi, os_, oc, _ = ix.search_batch(st, so, opts[pick], 40, n_threads=16)
    assert np.array_equal(cnt[pick], oc)
    mask = np.arange(40)[None, :] < oc[:, None]
    assert np.array_equal(ids[pick][mask], oi[mask]) and np.array_equal(sc[pick][mask], os_[mask])
    # properties of the whole batch: a noisy copy of doc d finds d first with ~75 of its 100 terms; lists are
    # ordered (score desc, id asc), respect the floor 5 and the 10 % cutoff anchored on the best
    hit = src >= 0
    assert (cnt[hit] >= 1).all()
    assert (ids[hit, 0] == (src[hit] + 1)).mean() > 0.999
    assert 60 < sc[hit, 0].mean() < 90
    valid = np.arange(40)[None, :] < cnt[:, None]
    assert (sc[valid] >= 5).all()
    key = (0xFFFFFFFF - sc.astype(np.int64)) * (1 << 32) + ids.astype(np.int64)
    both = valid[:, 1:] & valid[:, :-1]
    assert (key[:, 1:][both] > key[:, :-1][both]).all()
    best = np.where(cnt > 0, sc[:, 0], 0).astype(np.int64)
    assert (sc.astype(np.int64)[valid] >= np.maximum(5, best[:, None] * 10 // 100).repeat(40, axis=1)[valid]).all()
    # idempotence, and independence of the batch split
    ids2, sc2, cnt2 = reader.search_batch(terms[:30_000].reshape(-1), offs[:30_001], opts[:30_000], 40)
    assert np.array_equal(cnt2, cnt[:30_000])
    m2 = np.arange(40)[None, :] < cnt2[:, None]
    assert np.array_equal(ids2[m2], ids[:30_000][m2]) and np.array_equal(sc2[m2], sc[:30_000][m2])


def test_reference_50k_fingerprint_vector_through_the_gpu(ctx):
    """The reference's largest known-answer vector, tests/test_fingerprint_api.py:67-99: 50 000 fingerprints of
    100 CPython random.Random(i).randint(0, 2**18) hashes, inserted in batches of 1000; the query is fingerprint 100;
    HTTP defaults -> exactly [{id: 100, score: 100}].  Through the CUDA path, on a snapshot of several file and
    memory segments and on the fully merged one; plus 300 other fingerprints as queries against the oracle."""
    import random
    max_hash = 2 ** 18
    ix = OracleIndex()
    batch, fps = [], {}
    for i in range(1, 50001):
        rng = random.Random(i)
        fp = [rng.randint(0, max_hash) for _ in range(100)]
        if i % 167 == 0 or i == 100:
            fps[i] = fp
        batch.append(("insert", i, fp))
        if len(batch) == 1000:
            ix.update(batch)
            batch = []
            if ix.num_memory_segments >= 16:
                ix.merge_memory(0, 10)
            if (i // 1000) % 20 == 0:
                ix.checkpoint()
    assert ix.num_file_segments >= 1 and ix.num_memory_segments >= 1
    query = fps[100]
    for merged in (False, True):
        if merged:
            ix.checkpoint()
            ix.merge_files(0, ix.num_file_segments)
            assert ix.num_file_segments == 1
        snap = _snapshot_of(ctx, ix)
        reader = pkg.IndexReader(snap)
        assert [tuple(r) for r in pkg.multi_index_search(reader, pkg.SearchRequest(query))] == [(100, 100)]
        assert ix.search_http(query) == [(100, 100)]
        qs = [fps[i] for i in sorted(fps)]
        t, o = flat_queries(qs)
        _compare_batch(reader, ix, t, o, pkg.synth.http_opts(len(qs), 100), 40, threads=8)
        _compare_batch(reader, ix, t, o, np.tile(np.array((500, 1, 10), dtype=np.uint32), (len(qs), 1)), 500, threads=8)
        snap.release()
