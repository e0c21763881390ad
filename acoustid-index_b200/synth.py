"""Synthetic fingerprint corpora and query batches — input data only, not the search path.

Everything is counter-based SplitMix64 in wrapping int64 arithmetic, so the same (config, seed) gives the
same corpus on CPU and on GPU (torch is used for device memory and the sort; plumbing).

  draw(stream, n)   = mix64(seed_stream + (n + 1) * GAMMA)
  corpus term (doc i, slot j) = vocab_map(rank(draw(corpus, i*H + j)))      rank uniform in [0,V) or Zipf(s)
  vocab_map(v)      = low 32 bits of mix64(v + 0xACED0000)   (spreads the vocabulary over the u32 hash space)
  query q           = 90 %: the first T terms of a random doc, each replaced w.p. 1/4 by a random vocab term
                      10 %: T random vocab terms
"""
from dataclasses import dataclass

import numpy as np
import torch

_M64 = (1 << 64) - 1
GAMMA = 0x9E3779B97F4A7C15
_C1 = 0xBF58476D1CE4E5B9
_C2 = 0x94D049BB133111EB


def _s64(x):
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(z, s):
    return (z >> s) & ((1 << (64 - s)) - 1)


def mix64(z):
    z = (z ^ _lsr(z, 30)) * _s64(_C1)
    z = (z ^ _lsr(z, 27)) * _s64(_C2)
    return z ^ _lsr(z, 31)


def draw(seed, n):
    """n: int64 tensor of counters."""
    return mix64((n + 1) * _s64(GAMMA) + _s64(seed))


def vocab_map(v):
    return mix64(v + 0xACED0000) & 0xFFFFFFFF


@dataclass
class SynthConfig:
    n_docs: int
    hashes_per_doc: int
    vocab_log2: int
    seed: int = 0xF1D00001
    zipf_s: float = 0.0      # 0 = uniform vocabulary
    first_doc_id: int = 1


class Synth:
    def __init__(self, cfg: SynthConfig, device="cpu"):
        self.cfg = cfg
        self.device = torch.device(device)
        self.V = 1 << cfg.vocab_log2
        self._zipf_thresholds = None
        if cfg.zipf_s > 0:
            # integer inverse-CDF table, built once on the host so that CPU and GPU draws agree exactly
            w = 1.0 / np.power(np.arange(1, self.V + 1, dtype=np.float64), cfg.zipf_s)
            cdf = np.cumsum(w)
            cdf /= cdf[-1]
            thr = np.minimum(np.floor(cdf * float(1 << 62)), float((1 << 62) - 1)).astype(np.int64)
            thr[-1] = (1 << 62)
            self._zipf_thresholds = torch.from_numpy(thr).to(self.device)

    def rank(self, r):
        """random 64-bit draw -> vocabulary rank in [0, V)."""
        if self._zipf_thresholds is None:
            return r & (self.V - 1)
        u = _lsr(r, 2)  # uniform in [0, 2^62)
        return torch.searchsorted(self._zipf_thresholds, u, right=True).clamp_(max=self.V - 1)

    def corpus_terms(self, doc_index, slots):
        """doc_index: int64 [n] (0-based); slots: int64 [m] -> terms int64 [n, m] (values < 2^32)."""
        H = self.cfg.hashes_per_doc
        n = doc_index[:, None] * H + slots[None, :]
        return vocab_map(self.rank(draw(self.cfg.seed, n)))

    def corpus_items(self, chunk_docs=1 << 20, doc_lo=0, doc_hi=None):
        """The postings of docs [doc_lo, doc_hi) (0-based; default: all) as sorted packed items (hash<<32)|id, plus
        their docs map.  numpy, host memory.  A corpus cut into consecutive doc ranges gives the segments of a
        multi-segment index (a segment footer counts items in u32, filefmt.zig:76-80)."""
        cfg, dev = self.cfg, self.device
        H = cfg.hashes_per_doc
        doc_hi = cfg.n_docs if doc_hi is None else doc_hi
        n = doc_hi - doc_lo
        slots = torch.arange(H, dtype=torch.int64, device=dev)
        keys = torch.empty(n * H, dtype=torch.int64, device=dev)
        for d0 in range(doc_lo, doc_hi, chunk_docs):
            d1 = min(doc_hi, d0 + chunk_docs)
            di = torch.arange(d0, d1, dtype=torch.int64, device=dev)
            t = self.corpus_terms(di, slots)
            k = (t << 32) | (di[:, None] + cfg.first_doc_id)
            keys[(d0 - doc_lo) * H:(d1 - doc_lo) * H] = (k ^ _s64(1 << 63)).reshape(-1)  # signed order == unsigned order
        keys = torch.sort(keys).values
        keys ^= _s64(1 << 63)
        items = keys.cpu().numpy().view(np.uint64)
        del keys
        doc_ids = np.arange(cfg.first_doc_id + doc_lo, cfg.first_doc_id + doc_hi, dtype=np.uint32)
        doc_alive = np.ones(n, dtype=np.uint8)
        return items, doc_ids, doc_alive

    def doc_hashes(self, doc_index):
        """Full fingerprint (H terms, insertion order) of 0-based docs: uint32 [n, H]."""
        di = torch.as_tensor(doc_index, dtype=torch.int64, device=self.device)
        slots = torch.arange(self.cfg.hashes_per_doc, dtype=torch.int64, device=self.device)
        return self.corpus_terms(di, slots).cpu().numpy().astype(np.uint32)

    def queries(self, n_queries, terms_per_query, seed=0xF1D01001, single_term=False, first=0):
        """uint32 [Q, T] query terms (host numpy) and the source doc index (-1 for random queries): queries
        first .. first + n_queries - 1 of the stream `seed` (a slice of a batch equals the batch's slice)."""
        cfg, dev, T = self.cfg, self.device, terms_per_query
        q = torch.arange(first, first + n_queries, dtype=torch.int64, device=dev)
        stride = 2 + 2 * T
        head = draw(seed, q * stride)
        is_random = (_lsr(head, 8) % 10) == 0
        doc = _lsr(draw(seed, q * stride + 1), 1) % cfg.n_docs
        slots = torch.arange(T, dtype=torch.int64, device=dev)
        base = self.corpus_terms(doc, slots % cfg.hashes_per_doc)
        n2 = q[:, None] * stride + 2 + slots[None, :]
        replace = (_lsr(draw(seed, n2), 8) & 3) == 0
        fresh = vocab_map(draw(seed, n2 + T) & (self.V - 1))
        if single_term:
            terms = base
        else:
            terms = torch.where(is_random[:, None] | replace, fresh, base)
            doc = torch.where(is_random, torch.full_like(doc, -1), doc)
        return terms.cpu().numpy().astype(np.uint32), doc.cpu().numpy()


def http_opts(n_queries, terms_per_query, limit=40, score_pct=10):
    """Per-query options as the HTTP handler resolves them (server.zig:192, MultiIndex.zig:302-306)."""
    o = np.empty((n_queries, 3), dtype=np.uint32)
    o[:, 0] = limit
    o[:, 1] = (terms_per_query + 19) // 20
    o[:, 2] = score_pct
    return o
