// fpx_kernels.cu — sm_100a kernels of the batched `_search` path.
//
// Per query the reference does (src/Index.zig:170-177, src/common.zig:121-167):
//   sort+dedup terms -> per term scan its postings -> hash-map count per docid -> filter, sort, top-k.
// Here, over a whole batch, against the CSR built by fpx_snapshot_host.h:
//   prepare_kernel       one warp per query: dedup terms, probe the term directory, emit row
//                        descriptors, bin the query by posting volume
//   search_smem_kernel   persistent CTAs, one query at a time: stream the rows with 128-bit loads,
//                        count docids in a shared-memory open-addressing table (one packed 32-bit word
//                        per doc: quotient tag | probe number | count), then scan the table, rank the
//                        candidates by (score desc, id asc) and apply the reference's cutoffs
//   search_wide_kernel   same algorithm over a global-memory table (64-bit slots); takes whatever the
//                        packed table cannot represent exactly (count/probe overflow, > 512 candidates,
//                        limit > 512, very large queries) so results stay bit-exact in every case.
// All arithmetic is u32/u64 integer; there is no floating point on the path.
#include "fpx_kernels.cuh"

namespace fpx {

namespace {

constexpr uint32_t kMult = 0x9E3779B1u;  // odd => d -> d*kMult is a bijection on u32
constexpr uint32_t kMult2 = 0x85EBCA6Bu; // independent hash for multi-pass partitioning
constexpr uint32_t inv32(uint32_t a) {
    uint32_t x = a;
    for (int i = 0; i < 5; ++i) x *= 2u - a * x;
    return x;
}
constexpr uint32_t kMultInv = inv32(kMult);
static_assert(kMult * kMultInv == 1u, "modular inverse");

constexpr uint32_t FPX_UNSUPPORTED_CODE = 7;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool directory_lookup(const SnapshotDev &s, uint32_t term, uint32_t &start4, uint32_t &len) {
    uint32_t h = (term * kMult) >> s.table_shift;
    for (;;) {
        const uint4 e = __ldg(reinterpret_cast<const uint4 *>(s.table + h));
        if (e.w == 0) return false;
        if (e.x == term) {
            len = e.y;
            start4 = e.z;
            return true;
        }
        h = (h + 1) & s.table_mask;
    }
}

// Bin a prepared query (called by one thread).
__device__ void classify_and_enqueue(const BatchArgs &a, uint32_t q, uint32_t n_rows, unsigned long long postings,
                                     uint32_t n_unique) {
    const SearchOpts o = a.opts[q];
    const uint32_t k_eff = min(o.max_results, a.k_stride);
    if (a.stats) {
        atomicAdd(&a.stats->queries, 1ull);
        atomicAdd(&a.stats->unique_terms, (unsigned long long)n_unique);
        atomicAdd(&a.stats->postings, postings);
    }
    if (n_rows == 0 || k_eff == 0) {
        a.out_counts[q] = 0;
        return;
    }
    uint32_t cls, passes = 1;
    if (k_eff > kFastKbuf) {
        cls = kWideClass;
    } else if (postings <= 4096) {
        cls = 0;
    } else if (postings <= 8192) {
        cls = 1;
    } else if (postings <= 16384) {
        cls = 2;
    } else {
        cls = 2;
        const unsigned long long per_pass = 12288; // load <= 0.375 per pass on average
        while ((unsigned long long)passes * per_pass < postings && passes <= 64) passes <<= 1;
        if (passes > 64) cls = kWideClass;
    }
    QueryInfo qi;
    qi.n_rows = n_rows;
    qi.postings = postings > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)postings;
    qi.passes = passes;
    qi.reserved = 0;
    a.qinfo[q] = qi;
    const uint32_t pos = atomicAdd(&a.counters->qcount[cls], 1u);
    a.queues[(size_t)cls * a.n_queries + pos] = q;
}

// ------------------------------------------------------------------------------------------------
// Term directory build (snapshot commit)
// ------------------------------------------------------------------------------------------------
__global__ void build_table_kernel(TermEntry *table, uint32_t mask, uint32_t shift, const uint32_t *terms,
                                   const uint32_t *lens, const uint32_t *start4, unsigned long long n) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t t = terms[i];
        uint32_t h = (t * kMult) >> shift;
        for (;;) {
            if (atomicCAS(&table[h].used, 0u, 1u) == 0u) {
                table[h].term = t;
                table[h].len = lens[i];
                table[h].start4 = start4[i];
                break;
            }
            h = (h + 1) & mask;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// prepare: Index.zig:171-172 (sort + dedupSorted => the query is a SET) + term -> row lookup
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) prepare_kernel(BatchArgs a) {
    const uint32_t lane = lane_id();
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < a.n_queries; q += warps_total) {
        const unsigned long long o0 = a.term_offsets[q] - a.term_base;
        const unsigned long long T64 = a.term_offsets[q + 1] - a.term_offsets[q];
        if (T64 > kMaxQueryTerms) {
            if (lane == 0) {
                a.counters->error = FPX_UNSUPPORTED_CODE;
                a.out_counts[q] = 0;
            }
            continue;
        }
        const uint32_t T = (uint32_t)T64;
        if (T > kWarpQueryTerms) {
            if (lane == 0) a.long_queue[atomicAdd(&a.counters->long_count, 1u)] = q;
            continue;
        }
        uint32_t t[4];
        bool live[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t idx = lane + 32 * j;
            live[j] = idx < T;
            t[j] = live[j] ? __ldg(a.terms + o0 + idx) : 0u;
        }
        // drop later duplicates (the reference dedups after sorting; order is irrelevant to the result)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            if (jj * 32u >= T) break;
            for (uint32_t l = 0; l < 32; ++l) {
                const uint32_t s = jj * 32 + l;
                if (s >= T) break;
                const uint32_t v = __shfl_sync(0xFFFFFFFFu, t[jj], l);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (live[j] && s < lane + 32u * j && v == t[j]) live[j] = false;
            }
        }
        uint32_t n_unique = 0, n_rows = 0;
        unsigned long long postings = 0;
        uint32_t st[4], ln[4];
        bool found[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            st[j] = ln[j] = 0;
            found[j] = live[j] && directory_lookup(a.snap, t[j], st[j], ln[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t um = __ballot_sync(0xFFFFFFFFu, live[j]);
            const uint32_t fm = __ballot_sync(0xFFFFFFFFu, found[j]);
            n_unique += __popc(um);
            if (found[j]) {
                const uint32_t pos = n_rows + __popc(fm & ((1u << lane) - 1u));
                a.rows[o0 + pos] = make_uint2(st[j], ln[j]);
                postings += ln[j];
            }
            n_rows += __popc(fm);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) postings += __shfl_xor_sync(0xFFFFFFFFu, postings, off);
        if (lane == 0) classify_and_enqueue(a, q, n_rows, postings, n_unique);
    }
}

// Queries with 129..8192 raw terms: one CTA each, bitonic sort in shared memory.
__global__ void __launch_bounds__(kThreads) prepare_long_kernel(BatchArgs a) {
    __shared__ uint32_t s_terms[kMaxQueryTerms];
    __shared__ uint32_t s_idx, s_rows, s_unique;
    __shared__ unsigned long long s_post;
    const uint32_t tid = threadIdx.x;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_idx = atomicAdd(&a.counters->long_head, 1u);
        __syncthreads();
        if (s_idx >= a.counters->long_count) break;
        const uint32_t q = a.long_queue[s_idx];
        const unsigned long long o0 = a.term_offsets[q] - a.term_base;
        const uint32_t T = (uint32_t)(a.term_offsets[q + 1] - a.term_offsets[q]);
        uint32_t m = 256;
        while (m < T) m <<= 1;
        for (uint32_t i = tid; i < m; i += kThreads) s_terms[i] = i < T ? a.terms[o0 + i] : 0xFFFFFFFFu;
        if (tid == 0) {
            s_rows = 0;
            s_unique = 0;
            s_post = 0;
        }
        __syncthreads();
        for (uint32_t k = 2; k <= m; k <<= 1)
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t i = tid; i < m; i += kThreads) {
                    const uint32_t x = i ^ j;
                    if (x > i) {
                        const uint32_t u = s_terms[i], v = s_terms[x];
                        if ((u > v) == ((i & k) == 0)) {
                            s_terms[i] = v;
                            s_terms[x] = u;
                        }
                    }
                }
                __syncthreads();
            }
        // the first T sorted entries are exactly the query's terms (padding sorts last)
        for (uint32_t i = tid; i < T; i += kThreads) {
            const uint32_t v = s_terms[i];
            if (i > 0 && s_terms[i - 1] == v) continue;
            atomicAdd(&s_unique, 1u);
            uint32_t st, ln;
            if (directory_lookup(a.snap, v, st, ln)) {
                a.rows[o0 + atomicAdd(&s_rows, 1u)] = make_uint2(st, ln);
                atomicAdd(&s_post, (unsigned long long)ln);
            }
        }
        __syncthreads();
        if (tid == 0) classify_and_enqueue(a, q, s_rows, s_post, s_unique);
    }
}

// ------------------------------------------------------------------------------------------------
// ranking helpers: key = (0xFFFFFFFF - score) << 32 | id, ascending == (score desc, id asc)
// (common.zig:169-171 compareResults)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long rank_key(uint32_t score, uint32_t id) {
    return ((unsigned long long)(0xFFFFFFFFu - score) << 32) | id;
}

// Sort kbuf[0..n) ascending; m = padded power of two (entries n..m overwritten with ~0).
__device__ void block_sort_keys(unsigned long long *kbuf, uint32_t n, uint32_t cap) {
    const uint32_t tid = threadIdx.x;
    if (n <= 1) return;
    if (n <= 32) {
        if (tid < 32) {
            const unsigned long long key = tid < n ? kbuf[tid] : ~0ull;
            uint32_t rank = 0;
            for (uint32_t l = 0; l < n; ++l) rank += (__shfl_sync(0xFFFFFFFFu, key, l) < key) ? 1u : 0u;
            __syncwarp();
            if (tid < n) kbuf[rank] = key;
        }
        __syncthreads();
        return;
    }
    uint32_t m = 64;
    while (m < n) m <<= 1;
    if (m > cap) m = cap;
    for (uint32_t i = n + tid; i < m; i += kThreads) kbuf[i] = ~0ull;
    __syncthreads();
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = tid; i < m; i += kThreads) {
                const uint32_t x = i ^ j;
                if (x > i) {
                    const unsigned long long u = kbuf[i], v = kbuf[x];
                    if ((u > v) == ((i & k) == 0)) {
                        kbuf[i] = v;
                        kbuf[x] = u;
                    }
                }
            }
            __syncthreads();
        }
}

// common.zig:153-166: walk the ranked candidates, at most k_eff, relative cutoff anchored on the best.
// kbuf[0..n) sorted.  All threads of the CTA call this.
__device__ void emit_results(const BatchArgs &a, uint32_t q, const unsigned long long *kbuf, uint32_t n,
                             const SearchOpts &o) {
    const uint32_t tid = threadIdx.x;
    const uint32_t k_eff = min(o.max_results, a.k_stride);
    const uint32_t lim = min(n, k_eff);
    uint32_t ms = o.min_score;
    if (lim > 0) {
        const uint32_t s0 = 0xFFFFFFFFu - (uint32_t)(kbuf[0] >> 32);
        ms = max(ms, (uint32_t)(s0 * o.min_score_pct) / 100u); // u32 wrapping product, truncating division
    }
    uint32_t total = 0;
    for (uint32_t base = 0; base < lim; base += kThreads) {
        const uint32_t i = base + tid;
        bool pass = false;
        if (i < lim) {
            const unsigned long long key = kbuf[i];
            const uint32_t score = 0xFFFFFFFFu - (uint32_t)(key >> 32);
            pass = (i == 0) || score >= ms; // the best candidate is emitted before the cutoff is raised
            if (pass) {
                a.out_ids[(size_t)q * a.k_stride + i] = (uint32_t)key;
                a.out_scores[(size_t)q * a.k_stride + i] = score;
            }
        }
        total += __syncthreads_count(pass);
    }
    if (tid == 0) {
        a.out_counts[q] = total;
        if (a.stats) atomicAdd(&a.stats->results, (unsigned long long)total);
    }
}

// ------------------------------------------------------------------------------------------------
// shared-memory path
// ------------------------------------------------------------------------------------------------
template <int LOG> struct Packed {
    static constexpr int kRemBits = 32 - LOG;
    static constexpr int kProbeBits = 5;
    static constexpr int kCntBits = 32 - kRemBits - kProbeBits; // = LOG - 5
    static constexpr uint32_t kCntMask = (1u << kCntBits) - 1u;
    static constexpr uint32_t kSlotMask = (1u << LOG) - 1u;
    static constexpr uint32_t kRemMask = (1u << kRemBits) - 1u;
    static constexpr uint32_t kSlots = 1u << LOG;
};

// One packed word per distinct docid:  [ rem : 32-LOG | probe : 5 | count : LOG-5 ].
// hv = d*kMult is a bijection; its top LOG bits are the home slot, the rest is `rem`.  Probing is double
// hashing with an odd step derived from rem, so (slot, probe, rem) identifies d exactly.
template <int LOG> __device__ __forceinline__ void table_insert(uint32_t *tab, uint32_t d, uint32_t *ovf) {
    using P = Packed<LOG>;
    const uint32_t hv = d * kMult;
    const uint32_t rem = hv & P::kRemMask;
    const uint32_t step = ((rem << 1) | 1u) & P::kSlotMask;
    const uint32_t tagbase = rem << (P::kProbeBits + P::kCntBits);
    uint32_t s = hv >> P::kRemBits;
#pragma unroll 1
    for (uint32_t i = 0; i < 32; ++i) {
        const uint32_t tag = tagbase | (i << P::kCntBits);
        const uint32_t old = atomicCAS(tab + s, 0u, tag | 1u);
        if (old == 0u) return;
        if ((old & ~P::kCntMask) == tag) {
            const uint32_t prev = atomicAdd(tab + s, 1u);
            if ((prev & P::kCntMask) == P::kCntMask) *ovf = 1u; // count field wrapped: redo in the wide path
            return;
        }
        s = (s + step) & P::kSlotMask;
    }
    *ovf = 1u; // probe number does not fit
}

template <int LOG> __device__ __forceinline__ uint32_t table_docid(uint32_t word, uint32_t slot) {
    using P = Packed<LOG>;
    const uint32_t rem = word >> (P::kProbeBits + P::kCntBits);
    const uint32_t probe = (word >> P::kCntBits) & 31u;
    const uint32_t step = ((rem << 1) | 1u) & P::kSlotMask;
    const uint32_t home = (slot - probe * step) & P::kSlotMask;
    return ((home << P::kRemBits) | rem) * kMultInv;
}

template <int LOG> constexpr size_t smem_bytes_for() {
    return (size_t)Packed<LOG>::kSlots * 4 + kFastKbuf * 8 + kRowsChunk * 8;
}

template <int LOG>
__global__ void __launch_bounds__(kThreads, (LOG == 13 ? 4 : (LOG == 14 ? 3 : 1))) search_smem_kernel(BatchArgs a) {
    using P = Packed<LOG>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem_raw);
    unsigned long long *kbuf = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)P::kSlots * 4);
    uint2 *rows_s = reinterpret_cast<uint2 *>(smem_raw + (size_t)P::kSlots * 4 + kFastKbuf * 8);
    __shared__ uint32_t s_idx, s_ncand, s_ovf;

    constexpr int cls = LOG - 13;
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const uint4 *docids4 = reinterpret_cast<const uint4 *>(a.snap.docids);
    const uint32_t pad = a.snap.pad_id;
    uint4 *tab4 = reinterpret_cast<uint4 *>(tab);

    for (uint32_t i = tid; i < P::kSlots / 4; i += kThreads) tab4[i] = make_uint4(0, 0, 0, 0);
    const uint32_t qcount = a.counters->qcount[cls];

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_idx = atomicAdd(&a.counters->qhead[cls], 1u);
            s_ncand = 0;
            s_ovf = 0;
        }
        __syncthreads();
        if (s_idx >= qcount) break;
        const uint32_t q = a.queues[(size_t)cls * a.n_queries + s_idx];
        const QueryInfo qi = a.qinfo[q];
        const SearchOpts o = a.opts[q];
        const uint32_t thr = max(o.min_score, 1u); // a doc in the table has score >= 1
        const uint2 *rows = a.rows + (a.term_offsets[q] - a.term_base);
        const uint32_t pmask = qi.passes - 1u;

        for (uint32_t pass = 0; pass < qi.passes; ++pass) {
            for (uint32_t r0 = 0; r0 < qi.n_rows; r0 += kRowsChunk) {
                const uint32_t nr = min(kRowsChunk, qi.n_rows - r0);
                __syncthreads();
                for (uint32_t i = tid; i < nr; i += kThreads) rows_s[i] = rows[r0 + i];
                __syncthreads();
                // a warp per posting row, 128-bit loads, two rows in flight
                for (uint32_t r = warp * 2; r < nr; r += kWarps * 2) {
                    const uint2 ra = rows_s[r];
                    const uint2 rb = (r + 1 < nr) ? rows_s[r + 1] : make_uint2(0u, 0u);
                    const uint32_t na = (ra.y + 3) >> 2, nb = (rb.y + 3) >> 2;
                    const uint32_t nmax = max(na, nb);
                    for (uint32_t i = lane; i < nmax; i += 32) {
                        uint4 va = make_uint4(pad, pad, pad, pad), vb = va;
                        if (i < na) va = __ldg(docids4 + ra.x + i);
                        if (i < nb) vb = __ldg(docids4 + rb.x + i);
                        const uint32_t d[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            if (d[e] == pad) continue;
                            if (pmask && (((d[e] * kMult2) >> 16) & pmask) != pass) continue;
                            table_insert<LOG>(tab, d[e], &s_ovf);
                        }
                    }
                }
            }
            __syncthreads();
            // scan + clear; candidates are docs with score >= max(min_score,1)  (common.zig:140-145)
            for (uint32_t i = tid; i < P::kSlots / 4; i += kThreads) {
                const uint4 w = tab4[i];
                if ((w.x | w.y | w.z | w.w) == 0u) continue;
                tab4[i] = make_uint4(0, 0, 0, 0);
                const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t cnt = ws[e] & P::kCntMask;
                    if (ws[e] != 0u && cnt >= thr) {
                        const uint32_t pos = atomicAdd(&s_ncand, 1u);
                        if (pos < kFastKbuf) kbuf[pos] = rank_key(cnt, table_docid<LOG>(ws[e], i * 4 + e));
                    }
                }
            }
        }
        __syncthreads();
        const uint32_t n = s_ncand;
        if (s_ovf || n > kFastKbuf) {
            // not representable here: hand the query to the global-memory path (still exact)
            if (tid == 0) {
                a.queues[(size_t)kWideClass * a.n_queries + atomicAdd(&a.counters->qcount[kWideClass], 1u)] = q;
                if (a.stats) atomicAdd(&a.stats->overflow_requeues, 1ull);
            }
            continue;
        }
        block_sort_keys(kbuf, n, kFastKbuf);
        emit_results(a, q, kbuf, n, o);
    }
}

// ------------------------------------------------------------------------------------------------
// global-memory path: 64-bit slots (docid << 32 | count), linear probing, any size via hash partitions
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) search_wide_kernel(BatchArgs a) {
    __shared__ unsigned long long kbuf[kWideKbuf];
    __shared__ uint2 rows_s[kRowsChunk];
    __shared__ uint32_t s_idx, s_kn, s_new, s_fail;
    __shared__ unsigned long long s_kth;

    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    unsigned long long *table = a.wide_tables + ((size_t)blockIdx.x << a.wide_cap_log2);
    const uint4 *docids4 = reinterpret_cast<const uint4 *>(a.snap.docids);
    const uint32_t pad = a.snap.pad_id;
    const uint32_t qcount = a.counters->qcount[kWideClass];

    for (;;) {
        __syncthreads();
        if (tid == 0) s_idx = atomicAdd(&a.counters->qhead[kWideClass], 1u);
        __syncthreads();
        if (s_idx >= qcount) break;
        const uint32_t q = a.queues[(size_t)kWideClass * a.n_queries + s_idx];
        const QueryInfo qi = a.qinfo[q];
        const SearchOpts o = a.opts[q];
        const uint32_t thr = max(o.min_score, 1u);
        const uint32_t k_eff = min(min(o.max_results, a.k_stride), kMaxResults);
        const uint2 *rows = a.rows + (a.term_offsets[q] - a.term_base);
        if (tid == 0 && a.stats) atomicAdd(&a.stats->wide_queries, 1ull);

        // table size: aim for load <= 0.25; more hash partitions when one table cannot hold that
        const unsigned long long need = 4ull * qi.postings;
        uint32_t passes = 1;
        while ((need / passes) > (1ull << a.wide_cap_log2) && passes < (1u << 16)) passes <<= 1;
        bool done = false;
        while (!done) {
            uint32_t capl = 12;
            while (capl < a.wide_cap_log2 && (1ull << capl) < need / passes) ++capl;
            const uint32_t cap = 1u << capl, cmask = cap - 1u;
            __syncthreads();
            if (tid == 0) {
                s_kn = 0;
                s_kth = ~0ull;
            }
            bool failed = false;
            for (uint32_t pass = 0; pass < passes && !failed; ++pass) {
                for (uint32_t i = tid; i < cap / 2; i += kThreads)
                    reinterpret_cast<uint4 *>(table)[i] = make_uint4(0, 0, 0, 0);
                if (tid == 0) {
                    s_new = 0;
                    s_fail = 0;
                }
                __syncthreads();
                for (uint32_t r0 = 0; r0 < qi.n_rows; r0 += kRowsChunk) {
                    const uint32_t nr = min(kRowsChunk, qi.n_rows - r0);
                    __syncthreads();
                    for (uint32_t i = tid; i < nr; i += kThreads) rows_s[i] = rows[r0 + i];
                    __syncthreads();
                    for (uint32_t r = warp; r < nr; r += kWarps) {
                        const uint2 ra = rows_s[r];
                        const uint32_t na = (ra.y + 3) >> 2;
                        for (uint32_t i = lane; i < na; i += 32) {
                            const uint4 va = __ldg(docids4 + ra.x + i);
                            const uint32_t d[4] = {va.x, va.y, va.z, va.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const uint32_t id = d[e];
                                if (id == pad) continue;
                                if (passes > 1 && (((id * kMult2) >> 16) & (passes - 1u)) != pass) continue;
                                uint32_t h = (id * kMult) >> (32 - capl);
                                const unsigned long long fresh = ((unsigned long long)id << 32) | 1ull;
                                uint32_t tries = 0;
                                for (;;) {
                                    const unsigned long long old = atomicCAS(table + h, 0ull, fresh);
                                    if (old == 0ull) {
                                        atomicAdd(&s_new, 1u);
                                        break;
                                    }
                                    if ((uint32_t)(old >> 32) == id) {
                                        atomicAdd(table + h, 1ull);
                                        break;
                                    }
                                    h = (h + 1) & cmask;
                                    if (++tries >= cap) {
                                        s_fail = 1;
                                        break;
                                    }
                                }
                            }
                        }
                    }
                }
                __syncthreads();
                if (s_fail || s_new > cap - cap / 4) {
                    failed = true;
                    break;
                }
                // scan in rounds of 1024 slots so the candidate buffer can never overflow
                for (uint32_t base = 0; base < cap; base += 4 * kThreads) {
                    const uint4 *t4 = reinterpret_cast<const uint4 *>(table + base + 4 * tid);
                    const uint4 w0 = t4[0], w1 = t4[1];
                    const unsigned long long kth = s_kth;
                    const uint32_t ids[4] = {w0.y, w0.w, w1.y, w1.w};
                    const uint32_t cnts[4] = {w0.x, w0.z, w1.x, w1.z};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (cnts[e] >= thr) { // empty slots have count 0 < thr
                            const unsigned long long key = rank_key(cnts[e], ids[e]);
                            if (key < kth) kbuf[atomicAdd(&s_kn, 1u)] = key;
                        }
                    }
                    __syncthreads();
                    if (s_kn > kWideKbuf - 4 * kThreads) {
                        const uint32_t n = s_kn;
                        block_sort_keys(kbuf, n, kWideKbuf);
                        __syncthreads();
                        if (tid == 0) {
                            s_kn = min(n, k_eff);
                            if (n >= k_eff && k_eff > 0) s_kth = kbuf[k_eff - 1];
                        }
                        __syncthreads();
                    }
                }
            }
            if (failed) {
                if (passes >= (1u << 16)) { // cannot happen with < 2^32 postings; fail loudly instead of looping
                    if (tid == 0) {
                        a.counters->error = FPX_UNSUPPORTED_CODE;
                        a.out_counts[q] = 0;
                    }
                    done = true;
                    break;
                }
                passes <<= 1;
                continue;
            }
            __syncthreads();
            const uint32_t n = s_kn;
            block_sort_keys(kbuf, n, kWideKbuf);
            __syncthreads();
            emit_results(a, q, kbuf, n, o);
            done = true;
        }
    }
}

} // namespace

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
cudaError_t configure_kernels() {
    cudaError_t e;
    e = cudaFuncSetAttribute(search_smem_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_for<13>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(search_smem_kernel<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_for<14>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(search_smem_kernel<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_for<15>());
    return e;
}

void launch_build_table(TermEntry *table, uint32_t log2cap, const uint32_t *terms, const uint32_t *lens,
                        const uint32_t *start4, uint64_t n_terms, cudaStream_t st) {
    if (n_terms == 0) return;
    const uint32_t mask = (1u << log2cap) - 1u;
    unsigned long long blocks = (n_terms + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    build_table_kernel<<<(unsigned)blocks, 256, 0, st>>>(table, mask, 32 - log2cap, terms, lens, start4, n_terms);
}

void launch_prepare(const BatchArgs &a, cudaStream_t st) {
    if (a.n_queries == 0) return;
    unsigned long long blocks = ((unsigned long long)a.n_queries + kWarps - 1) / kWarps;
    if (blocks > 148 * 8) blocks = 148 * 8;
    prepare_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(a);
}

void launch_prepare_long(const BatchArgs &a, cudaStream_t st, int n_sms) {
    prepare_long_kernel<<<n_sms, kThreads, 0, st>>>(a);
}

void launch_search_class(const BatchArgs &a, int cls, cudaStream_t st, int n_sms) {
    switch (cls) {
    case 0: search_smem_kernel<13><<<n_sms * 4, kThreads, smem_bytes_for<13>(), st>>>(a); break;
    case 1: search_smem_kernel<14><<<n_sms * 3, kThreads, smem_bytes_for<14>(), st>>>(a); break;
    case 2: search_smem_kernel<15><<<n_sms * 1, kThreads, smem_bytes_for<15>(), st>>>(a); break;
    default: break;
    }
}

int wide_ctas(int n_sms) { return n_sms > 64 ? 64 : n_sms; }

void launch_search_wide(const BatchArgs &a, cudaStream_t st, int n_ctas) {
    search_wide_kernel<<<n_ctas, kThreads, 0, st>>>(a);
}

} // namespace fpx
