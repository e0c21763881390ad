"""Multi-GPU plumbing for the `_search` path (one process per GPU, torch.distributed).

Two modes:
  replicated   every rank holds the whole snapshot and answers its own contiguous slice of the query batch;
               the only exchange is gathering the per-query result lists (NCCL all-gather over NVLink).
  sharded      (snapshot larger than one GPU's HBM) rank g holds the postings whose docid falls in its range
               (fpx_snapshot_set_doc_range: the reference's scan caps and supersession rules are applied on
               the whole snapshot first), every rank sees every query and returns its local top-k under the
               absolute floor only (min_score_pct = 0); the lists are all-gathered and merged with
               fpx_merge_shard_results, which applies the relative cutoff anchored on the global best
               (common.zig:153-166).
torch.distributed is plumbing here; the merge itself is the C-ABI call.
"""
import numpy as np
import torch
import torch.distributed as dist

from .index import merge_shard_results


def query_slice(n_queries, rank, world):
    """Contiguous slice [lo, hi) of the batch answered by `rank` in replicated mode."""
    per = (n_queries + world - 1) // world
    lo = min(n_queries, rank * per)
    return lo, min(n_queries, lo + per)


def doc_ranges(min_id, max_id, world):
    """Equal-width docid ranges [lo, hi) covering [min_id, max_id]; the last one is open-ended."""
    span = max(1, (max_id - min_id + world) // world)
    out = []
    for g in range(world):
        lo = min_id + g * span if g else 0
        hi = min_id + (g + 1) * span if g + 1 < world else 0xFFFFFFFF
        out.append((lo, hi))
    return out


def all_gather_results(ids, scores, counts, group=None):
    """ids/scores: [nq, k] uint32-compatible tensors or arrays, counts: [nq].  Returns stacked
    ([world, nq, k], [world, nq, k], [world, nq]) numpy arrays on every rank."""
    world = dist.get_world_size(group)

    def gather(x):
        t = torch.as_tensor(np.ascontiguousarray(x).view(np.int32)) if isinstance(x, np.ndarray) else x
        bufs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(bufs, t, group=group)
        return np.stack([b.cpu().numpy().view(np.uint32) for b in bufs])

    return gather(ids), gather(scores), gather(counts)


def sharded_search(reader, terms, offsets, opts, k_stride, group=None):
    """Docid-range sharded search: local top-k with the absolute floor, all-gather, exact merge."""
    opts = np.ascontiguousarray(opts, dtype=np.uint32).reshape(-1, 3)
    local_opts = opts.copy()
    local_opts[:, 2] = 0  # relative cutoff is applied after the merge, anchored on the global best
    ids, sc, cnt = reader.search_batch(terms, offsets, local_opts, k_stride)
    g_ids, g_sc, g_cnt = all_gather_results(ids, sc, cnt, group)
    return merge_shard_results(g_ids, g_sc, g_cnt, opts, k_stride)
