"""Multi-GPU plumbing for the `_search` path (one process per GPU, torch.distributed).

Two modes:
  replicated   every rank holds the whole snapshot and answers its own contiguous slice of the query batch;
               the only exchange is gathering the per-query result lists (NCCL all-gather over NVLink).
  sharded      (snapshot larger than one GPU's HBM) rank g holds the postings whose docid falls in its range
               (fpx_snapshot_set_doc_range: the reference's scan caps and supersession rules are applied on
               the whole snapshot first), every rank sees every query and returns its local top-k under the
               absolute floor only (min_score_pct = 0); the lists are all-gathered and merged, and the merge
               applies the relative cutoff anchored on the global best (common.zig:153-166).
               ShardedSearch keeps everything on the device: results packed to {counts, offsets, (id, score) pairs}
               (fpx_pack_results_device), a one-word all-gather of the packed sizes, an NCCL all-gather of exactly the
               largest rank's bytes, and fpx_merge_packed_shards_device on the same stream.  sharded_search is the
               host-buffer version (gloo-testable) with fpx_merge_shard_results.
torch.distributed is plumbing here; search, packing and the merge are C-ABI calls.
"""
import numpy as np
import torch
import torch.distributed as dist

from .index import merge_packed_shards_device, merge_shard_results, pack_results_device


def query_slice(n_queries, rank, world):
    """Contiguous slice [lo, hi) of the batch answered by `rank` in replicated mode."""
    per = (n_queries + world - 1) // world
    lo = min(n_queries, rank * per)
    return lo, min(n_queries, lo + per)


def doc_ranges(min_id, max_id, world):
    """Equal-width docid ranges [lo, hi) covering [min_id, max_id]; the first starts at 0, the last one is open-ended:
    its hi is 0, which fpx_snapshot_set_doc_range reads as 2^32 when lo > 0 (so the id 0xFFFFFFFF has a shard too).
    A single shard is (0, 0) = no range."""
    if world == 1:
        return [(0, 0)]
    span = max(1, (max_id - min_id + world) // world)
    out = []
    for g in range(world):
        lo = min_id + g * span if g else 0
        hi = min_id + (g + 1) * span if g + 1 < world else 0
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


class ShardedSearch:
    """Docid-range sharded search with device-resident buffers (one instance per rank and batch shape).

    step(d_terms, d_offsets, d_local_opts, d_opts) -> (d_ids, d_scores, d_counts) torch tensors holding the merged
    answer on every rank.  d_local_opts are the queries' options with min_score_pct = 0."""

    def __init__(self, reader, nq, k_stride, device, group=None):
        self.reader, self.nq, self.k, self.dev, self.group = reader, int(nq), int(k_stride), device, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cap = self.nq * self.k                       # worst case; the exchange moves what is needed only
        i32 = dict(dtype=torch.int32, device=device)
        self.ids = torch.zeros((self.nq, self.k), **i32)
        self.sc = torch.zeros((self.nq, self.k), **i32)
        self.cnt = torch.zeros(self.nq, **i32)
        self.out_ids = torch.zeros((self.nq, self.k), **i32)
        self.out_sc = torch.zeros((self.nq, self.k), **i32)
        self.out_cnt = torch.zeros(self.nq, **i32)
        self.header = 2 * self.nq + 2
        self.packed = torch.zeros(self.header + 2 * self.cap, **i32)
        self.sizes = torch.zeros(self.world, **i32)
        self.gathered = None
        self.last_words = 0

    def step(self, d_terms, d_offsets, d_local_opts, d_opts):
        st = torch.cuda.current_stream(self.dev).cuda_stream
        nq, k = self.nq, self.k
        self.reader.search_batch_device(nq, d_terms.data_ptr(), d_offsets.data_ptr(), d_local_opts.data_ptr(), k,
                                        self.ids.data_ptr(), self.sc.data_ptr(), self.cnt.data_ptr(), st)
        pack_results_device(nq, k, self.ids.data_ptr(), self.sc.data_ptr(), self.cnt.data_ptr(), self.packed.data_ptr(),
                            self.cap, st)
        if self.world > 1:
            # phase 1: how many pairs does each rank hold (word 2n of its block); phase 2: exactly that much
            dist.all_gather_into_tensor(self.sizes, self.packed[2 * nq:2 * nq + 1], group=self.group)
            words = self.header + 2 * int(self.sizes.max().item())
            if self.gathered is None or self.gathered.numel() < self.world * words:
                self.gathered = torch.empty(self.world * (words + words // 4), dtype=torch.int32, device=self.dev)
            recv = self.gathered[:self.world * words]
            dist.all_gather_into_tensor(recv, self.packed[:words], group=self.group)
        else:
            words, recv = self.packed.numel(), self.packed
        self.last_words = words
        merge_packed_shards_device(self.world, nq, recv.data_ptr(), words, d_opts.data_ptr(), k, self.out_ids.data_ptr(),
                                   self.out_sc.data_ptr(), self.out_cnt.data_ptr(), st)
        return self.out_ids, self.out_sc, self.out_cnt
