"""The key lemma: counting over the compiled CSR + the reference ranking == the reference's
per-query segment walk, for any snapshot shape.  Host-side only (FPX_FLAG_HOST_ONLY), oracle as checker."""
import numpy as np
import pytest

from _helpers import csr_rank, pkg, segments_from_oracle
from _oracle import OracleIndex


def _compile(ix, doc_range=None):
    ctx = pkg.Context(host_only=True, host_threads=3)
    files, mems = segments_from_oracle(ix)
    b = pkg.SnapshotBuilder(ctx)
    for s in files:
        b.add_file_segment(s)
    for s in mems:
        b.add_memory_segment(s)
    if doc_range:
        b.set_doc_range(*doc_range)
    csr = b.csr()
    b.abort()
    ctx.close()
    return csr


def _check(ix, queries, opts_list):
    terms, offs, docids = _compile(ix)
    for q in queries:
        for (k, ms, pct) in opts_list:
            want = ix.search(q, k, ms, pct)
            got = csr_rank(terms, offs, docids, q, k, ms, pct)
            assert got == want, (q[:8], k, ms, pct, got[:5], want[:5])


OPTS = [(40, 1, 10), (10, 1, 0), (3, 2, 50), (100, 1, 100), (500, 1, 10), (5, 3, 200), (40, 0, 10)]


def test_single_memory_segment():
    ix = OracleIndex()
    ix.update([("insert", 1, [100, 200, 300]), ("insert", 2, [100, 200]), ("insert", 3, [100, 100, 100])])
    _check(ix, [[100], [100, 200, 300], [100, 100], [999], []], OPTS)


def _random_index(rng, n_docs, H, vocab, n_rounds, hot=(), flush_every=3, merge_at=4):
    ix = OracleIndex()
    next_id = 1
    all_ids = []
    for r in range(n_rounds):
        changes = []
        for _ in range(n_docs):
            roll = rng.random()
            if all_ids and roll < 0.15:    # re-insert an existing id with new hashes (update)
                did = int(rng.choice(all_ids))
            elif all_ids and roll < 0.22:  # delete
                changes.append(("delete", int(rng.choice(all_ids))))
                continue
            else:
                did = next_id
                next_id += 1
                all_ids.append(did)
            hs = rng.integers(0, vocab, size=H).tolist()
            for i in range(H):
                if hot and rng.random() < 0.25:
                    hs[i] = int(rng.choice(hot))
            if rng.random() < 0.3:
                hs[1] = hs[0]              # duplicate hash inside one fingerprint
            changes.append(("insert", did, hs))
        ix.update(changes)
        if r % flush_every == flush_every - 1:
            ix.checkpoint()
            if ix.num_file_segments >= merge_at:
                ix.merge_files(0, 2)
        elif ix.num_memory_segments >= 3:
            ix.merge_memory(0, 2)
    return ix, all_ids


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_multi_segment_updates_deletes(seed):
    rng = np.random.default_rng(seed)
    ix, _ = _random_index(rng, 60, 12, 300, 11)
    assert ix.num_file_segments >= 2 and ix.num_memory_segments >= 1
    queries = [rng.integers(0, 300, size=int(rng.integers(1, 40))).tolist() for _ in range(60)]
    _check(ix, queries, OPTS)


def test_scan_caps_hot_terms():
    """Hot hashes whose runs span > 4 blocks / > 1000 docs: the reference stops scanning
    (FileSegment.zig:173-174) and the compiled rows must stop at exactly the same posting."""
    rng = np.random.default_rng(7)
    ix, _ = _random_index(rng, 1500, 8, 5000, 6, hot=(11, 22, 4000), flush_every=2, merge_at=3)
    ix.checkpoint()
    ix.merge_files(0, ix.num_file_segments)
    assert ix.num_file_segments == 1
    terms, offs, docids = _compile(ix)
    lens = {int(t): int(offs[i + 1] - offs[i]) for i, t in enumerate(terms)}
    v = ix.file_segment(0)
    assert v.num_items > sum(lens.values()), "the caps must have cut something in this corpus"
    for hot in (11, 22, 4000):
        assert 1000 < lens[hot] <= 1344
        # limit 1 << 20 > any cap, min_score 1, no relative cutoff: the oracle returns exactly the reachable docs
        want = ix.search([hot], 1 << 20, 1, 0)
        p = int(np.searchsorted(terms, hot))
        row = docids[int(offs[p]):int(offs[p + 1])]
        ids, cnt = np.unique(row, return_counts=True)
        got = sorted(zip(ids.tolist(), cnt.tolist()), key=lambda x: (-x[1], x[0]))
        assert got == want
    queries = [[11, 22, 4000] + rng.integers(0, 5000, size=20).tolist() for _ in range(10)]
    _check(ix, queries, [(40, 1, 10), (1000, 1, 0), (100, 2, 10)])


def test_caps_with_several_file_segments_and_memory():
    rng = np.random.default_rng(11)
    ix, _ = _random_index(rng, 1200, 6, 3000, 7, hot=(5,), flush_every=2, merge_at=99)
    assert ix.num_file_segments >= 3 and ix.num_memory_segments >= 1
    queries = [[5] + rng.integers(0, 3000, size=10).tolist() for _ in range(10)]
    _check(ix, queries, [(40, 1, 10), (1000, 1, 0)])


def test_tombstone_and_reinsert_chain():
    ix = OracleIndex()
    ix.update([("insert", 1, [10, 20, 30]), ("insert", 2, [10, 20])])
    ix.checkpoint()
    ix.update([("delete", 1)])
    ix.checkpoint()
    ix.update([("insert", 1, [10]), ("insert", 3, [10, 20, 30])])
    ix.update([("insert", 2, [99])])
    _check(ix, [[10, 20, 30], [10], [99], [20]], OPTS)
    terms, offs, docids = _compile(ix)
    # id 2's old postings (10, 20) are dead: a newer segment mentions id 2
    p = int(np.searchsorted(terms, 20))
    assert docids[int(offs[p]):int(offs[p + 1])].tolist() == [3]


def test_doc_range_shards_union_to_the_whole():
    rng = np.random.default_rng(5)
    ix, ids = _random_index(rng, 80, 10, 200, 8)
    full = _compile(ix)
    mid = int(np.median(ids))
    a = _compile(ix, (0, mid))
    b = _compile(ix, (mid, 0xFFFFFFFF))
    assert int(a[1][-1]) + int(b[1][-1]) == int(full[1][-1])
    assert (a[2] < mid).all() and (b[2] >= mid).all()
    for t_i, t in enumerate(full[0]):
        row = full[2][int(full[1][t_i]):int(full[1][t_i + 1])]
        parts = []
        for c in (a, b):
            p = int(np.searchsorted(c[0], t))
            if p < len(c[0]) and c[0][p] == t:
                parts.append(c[2][int(c[1][p]):int(c[1][p + 1])])
        assert np.array_equal(np.sort(np.concatenate(parts)), row)


def test_rejects_bad_input():
    ctx = pkg.Context(host_only=True)
    ix = OracleIndex()
    ix.update([("insert", 1, [1, 2, 3])])
    ix.checkpoint()
    ix.update([("insert", 2, [1])])
    files, mems = segments_from_oracle(ix)
    b = pkg.SnapshotBuilder(ctx)
    b.add_memory_segment(mems[0])
    with pytest.raises(pkg.FpxError) as e:   # file after memory / not ascending commit ids
        b.add_file_segment(files[0])
    assert e.value.status == pkg._ffi.FPX_INVALID_SEGMENT
    b.abort()
    bad = files[0]
    bad.block_index = bad.block_index.copy()
    bad.block_index[0] += 1
    b = pkg.SnapshotBuilder(ctx)
    with pytest.raises(pkg.FpxError) as e:
        b.add_file_segment(bad)
    assert e.value.status == pkg._ffi.FPX_INVALID_SEGMENT
    b.abort()
    ctx.close()
