// fpx_kernels.cu — sm_100a kernels of the batched `_search` path.
//
// Per query the reference does (src/Index.zig:170-177, src/common.zig:121-167):
//   sort+dedup terms -> per term scan its postings -> hash-map count per docid -> filter, sort, top-k.
// Here, over a whole batch, against the CSR built by fpx_snapshot_host.h:
//   prepare_kernel       one warp per query: dedup terms, probe the term directory, emit row
//                        descriptors, bin the query by posting volume
//   search_find_kernel   the hot path (2 <= min_score <= 128, i.e. every HTTP-default query of >= 21 terms whose
//                        padded rows fit a shared-memory stage): warp-specialised persistent CTAs.  Producer warps
//                        gather the query's posting rows into a stage with TMA bulk copies (cp.async.bulk + mbarrier
//                        complete_tx); two groups of counter warps add every docid to a sketch of 8-bit counters with
//                        fire-and-forget shared atomics, then read the sketch back for the counters that reached
//                        min_score; resolver warps find those counters' postings in the staged rows (sorted by the
//                        sketch hash), count them per docid exactly and rank.  The sketch never under-counts and
//                        scores never come from it, so this is exact.
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
static_assert(kMult == kRowMult, "rows are ordered by the hash the sketch counts with");

constexpr uint32_t FPX_UNSUPPORTED_CODE = 7, FPX_INVALID_ARGUMENT_CODE = 2;
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

// Which exact shared-memory class (1..3) or the wide class holds `postings` postings.
__device__ __forceinline__ uint32_t exact_class_for(unsigned long long postings, uint32_t k_eff) {
    if (k_eff > kFastKbuf) return kWideClass;
    if (postings <= 4096) return 1;
    if (postings <= 8192) return 2;
    if (postings <= 64ull * 12288) return 3; // multi-pass above 16384 (see search_smem_kernel)
    return kWideClass;
}

__device__ __forceinline__ void enqueue(const BatchArgs &a, uint32_t cls, const WorkItem &w) {
    const uint32_t pos = atomicAdd(&a.counters->qcount[cls], 1u);
    a.items[(size_t)cls * a.n_queries + pos] = w;
}

// Decide how a prepared query is answered.  Returns the class, or kNumClasses when the answer is already
// known to be empty (no term present / limit 0).
__device__ uint32_t make_item(const BatchArgs &a, uint32_t q, uint32_t rows_off, uint32_t n_rows,
                              unsigned long long postings, unsigned long long total4, WorkItem &w) {
    const SearchOpts o = a.opts[q];
    const uint32_t k_eff = min(o.max_results, a.k_stride);
    w.q = q;
    w.rows_off = rows_off;
    w.n_rows = n_rows;
    w.total4 = total4 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)total4;
    w.postings = postings > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)postings;
    w.k_eff = k_eff;
    w.min_score = o.min_score;
    w.min_score_pct = o.min_score_pct;
    if (n_rows == 0 || k_eff == 0) return kNumClasses;
    // The sketch path needs 2 <= min_score <= 128 (bias of its 8-bit counters) and the query's padded rows must
    // fit one stage.  With a low floor and many postings too many counters reach it by chance: the expected
    // number of such counters is 32768 * P(Poisson(postings / 32768) >= min_score); the limits keep it <= 4
    // (kHotCap is 32; beyond it the query is re-queued, so this is a matter of speed only).
    bool sketch_ok = a.use_sketch && o.min_score >= 2 && o.min_score <= 128 && total4 <= kStageU4 &&
                     k_eff <= kFastKbuf && n_rows <= kSketchMaxRows;
    if (o.min_score == 2 && postings > 512) sketch_ok = false;
    if (o.min_score == 3 && postings > 2900) sketch_ok = false;
    if (o.min_score == 4 && postings > 7600) sketch_ok = false;
    if (!sketch_ok) return exact_class_for(postings, k_eff);
    return kSketchClass;
}

// Bin a prepared query (called by one thread).
__device__ void classify_and_enqueue(const BatchArgs &a, uint32_t q, uint32_t rows_off, uint32_t n_rows,
                                     unsigned long long postings, unsigned long long total4, uint32_t n_unique) {
    if (a.stats) {
        atomicAdd(&a.stats->queries, 1ull);
        atomicAdd(&a.stats->unique_terms, (unsigned long long)n_unique);
        atomicAdd(&a.stats->postings, postings);
    }
    WorkItem w;
    const uint32_t cls = make_item(a, q, rows_off, n_rows, postings, total4, w);
    if (cls == kNumClasses)
        a.out_counts[q] = 0;
    else
        enqueue(a, cls, w);
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
// 48 registers: five CTAs (40 warps) per SM keep more directory probes in flight than the unconstrained 58
__global__ void __launch_bounds__(kThreads, 5) prepare_kernel(BatchArgs a) {
    __shared__ uint32_t s_set[kWarps][256]; // per-warp hash set for de-duplication (<= 128 terms, load <= 0.5)
    const uint32_t lane = lane_id();
    uint32_t *set = s_set[threadIdx.x >> 5];
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // lane i parks the work item of the i-th query this warp prepared; they are enqueued 32 at a time
    WorkItem mine{};
    uint32_t mine_cls = kNumClasses, n_parked = 0;
    unsigned long long st_q = 0, st_unique = 0, st_post = 0; // lane 0: statistics

    auto flush = [&]() {
#pragma unroll
        for (uint32_t c = 0; c < (uint32_t)kNumClasses; ++c) {
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, mine_cls == c);
            if (m == 0u) continue;
            const uint32_t leader = __ffs(m) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(&a.counters->qcount[c], (uint32_t)__popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (mine_cls == c) a.items[(size_t)c * a.n_queries + base + __popc(m & lt_mask)] = mine;
        }
        mine_cls = kNumClasses;
        n_parked = 0;
    };

    for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < a.n_queries; q += warps_total) {
        const unsigned long long o0 = a.term_offsets[q] - a.term_base;
        const unsigned long long T64 = a.term_offsets[q + 1] - a.term_offsets[q];
        // offsets that do not lie inside the batch's terms (the device API takes them on trust from device memory)
        if (a.term_offsets[q] < a.term_base || a.term_offsets[q + 1] < a.term_offsets[q] || o0 + T64 > a.n_terms_total) {
            if (lane == 0) {
                a.counters->error = FPX_INVALID_ARGUMENT_CODE;
                a.out_counts[q] = 0;
            }
            continue;
        }
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
        // The query is a set (Index.zig:171-172 sorts and dedups): drop repeats with a small hash set.
        // Slot value 0 means empty, so the term 0 is handled by a ballot instead.
#pragma unroll
        for (int j = 0; j < 8; ++j) set[lane + 32 * j] = 0u;
        __syncwarp();
        bool zero_seen = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool is_zero = live[j] && t[j] == 0u;
            const uint32_t zm = __ballot_sync(0xFFFFFFFFu, is_zero);
            if (is_zero && (zero_seen || (zm & lt_mask))) live[j] = false;
            zero_seen = zero_seen || zm != 0u;
            if (live[j] && t[j] != 0u) {
                uint32_t h = (t[j] * kMult) >> 24;
                for (;;) {
                    const uint32_t old = atomicCAS(set + h, 0u, t[j]);
                    if (old == 0u) break;
                    if (old == t[j]) {
                        live[j] = false;
                        break;
                    }
                    h = (h + 1) & 255u;
                }
            }
        }
        __syncwarp();
        // term directory: issue the first probe of all four lookups before looking at any of them
        uint32_t h[4];
        uint4 e[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = (t[j] * kMult) >> a.snap.table_shift;
            e[j] = make_uint4(0, 0, 0, 0);
            if (live[j]) e[j] = __ldg(reinterpret_cast<const uint4 *>(a.snap.table + h[j]));
        }
        // Linear probing, an empty slot ends the search.  A lane that misses its first slot reads the next
        // three together: the warp waits for its slowest lane (max displacement over 32 lanes ~ 5 at load 0.5),
        // so a window of three turns ~5 dependent round trips into ~2.
        uint32_t cnt = 0, live_n = 0;
        unsigned long long sz = 0, postings = 0;
        bool found[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            while (e[j].w != 0u && e[j].x != t[j]) {
                const uint4 e1 = __ldg(reinterpret_cast<const uint4 *>(a.snap.table + ((h[j] + 1) & a.snap.table_mask)));
                const uint4 e2 = __ldg(reinterpret_cast<const uint4 *>(a.snap.table + ((h[j] + 2) & a.snap.table_mask)));
                const uint4 e3 = __ldg(reinterpret_cast<const uint4 *>(a.snap.table + ((h[j] + 3) & a.snap.table_mask)));
                h[j] = (h[j] + 3) & a.snap.table_mask;
                e[j] = (e1.w == 0u || e1.x == t[j]) ? e1 : (e2.w == 0u || e2.x == t[j]) ? e2 : e3;
            }
            found[j] = live[j] && e[j].w != 0u;
            live_n += live[j] ? 1u : 0u;
            if (found[j]) {
                cnt += 1;
                sz += (e[j].y + 3) >> 2;
                postings += e[j].y;
            }
        }
        // One warp scan for both prefixes: rows are laid out lane-major (lane 0's terms, then lane 1's, ...);
        // a row's place in a stage is the exclusive prefix of the padded row sizes in that order.
        unsigned long long incl = ((unsigned long long)cnt << 56) | sz; // <= 128 rows; sz < 2^34
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += y;
        }
        const unsigned long long tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
        const uint32_t n_rows = (uint32_t)(tot >> 56);
        const unsigned long long total4 = tot & 0x00FFFFFFFFFFFFFFull;
        {
            unsigned long long excl = incl - (((unsigned long long)cnt << 56) | sz);
            uint32_t r = (uint32_t)(excl >> 56);
            unsigned long long off4 = excl & 0x00FFFFFFFFFFFFFFull;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (found[j]) {
                    a.rows[o0 + r] = make_uint4(e[j].z, e[j].y, (uint32_t)off4, 0u);
                    r += 1;
                    off4 += (e[j].y + 3) >> 2;
                }
        }
        const uint32_t n_unique = __reduce_add_sync(0xFFFFFFFFu, live_n);
        {   // 64-bit sum in two 32-bit warp reductions (per-lane postings < 2^34)
            const uint32_t lo = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)(postings & 0xFFFFFFull));
            const uint32_t hi = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)(postings >> 24));
            postings = ((unsigned long long)hi << 24) + lo;
        }
        // every lane computes the (identical) work item; lane n_parked keeps it
        WorkItem w;
        const uint32_t cls = make_item(a, q, (uint32_t)o0, n_rows, postings, total4, w);
        if (lane == 0) {
            st_q += 1;
            st_unique += n_unique;
            st_post += postings;
            if (cls == kNumClasses) a.out_counts[q] = 0;
        }
        if (cls != kNumClasses) {
            if (lane == n_parked) {
                mine = w;
                mine_cls = cls;
            }
            if (++n_parked == 32) flush();
        }
    }
    flush();
    if (lane == 0 && a.stats && st_q) {
        atomicAdd(&a.stats->queries, st_q);
        atomicAdd(&a.stats->unique_terms, st_unique);
        atomicAdd(&a.stats->postings, st_post);
    }
}

// Queries with 129..8192 raw terms: one CTA each, bitonic sort in shared memory.
__global__ void __launch_bounds__(kThreads) prepare_long_kernel(BatchArgs a) {
    __shared__ uint32_t s_terms[kMaxQueryTerms];
    __shared__ uint32_t s_idx, s_rows, s_unique;
    __shared__ unsigned long long s_post, s_tot4;
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
            s_tot4 = 0;
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
                const unsigned long long off4 = atomicAdd(&s_tot4, (unsigned long long)((ln + 3) >> 2));
                a.rows[o0 + atomicAdd(&s_rows, 1u)] = make_uint4(st, ln, (uint32_t)off4, 0u);
                atomicAdd(&s_post, (unsigned long long)ln);
            }
        }
        __syncthreads();
        if (tid == 0) classify_and_enqueue(a, q, (uint32_t)o0, s_rows, s_post, s_tot4, s_unique);
    }
}

// ------------------------------------------------------------------------------------------------
// ranking helpers: key = (0xFFFFFFFF - score) << 32 | id, ascending == (score desc, id asc)
// (common.zig:169-171 compareResults).  "group" = the threads that cooperate on one query: the whole
// CTA (barrier 0) in the exact kernels, the consumer warps (barrier 1) in the sketch kernel.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long rank_key(uint32_t score, uint32_t id) {
    return ((unsigned long long)(0xFFFFFFFFu - score) << 32) | id;
}

// Named barriers (bar.sync / bar.arrive are the .aligned forms: every thread of the warp has to execute them
// together, so re-converge first; compute-sanitizer's synccheck checks exactly this).
__device__ __forceinline__ void named_sync(uint32_t id, uint32_t n_threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}
__device__ __forceinline__ void named_arrive(uint32_t id, uint32_t n_threads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

struct Group {
    uint32_t tid, size, bar; // thread index in the group, group size (multiple of 32), named barrier id
    __device__ __forceinline__ void sync() const {
        if (size == 32u)
            __syncwarp(); // a single warp needs no barrier resource
        else
            named_sync(bar, size);
    }
};

// Sort kbuf[0..n) ascending (entries n..pow2 are overwritten with ~0).  All group threads call this.
__device__ void group_sort_keys(const Group &g, unsigned long long *kbuf, uint32_t n, uint32_t cap) {
    if (n <= 1) return;
    if (n <= 32) {
        if (g.tid < 32) {
            const unsigned long long key = g.tid < n ? kbuf[g.tid] : ~0ull;
            uint32_t rank = 0;
            for (uint32_t l = 0; l < n; ++l) rank += (__shfl_sync(0xFFFFFFFFu, key, l) < key) ? 1u : 0u;
            __syncwarp();
            if (g.tid < n) kbuf[rank] = key;
        }
        g.sync();
        return;
    }
    uint32_t m = 64;
    while (m < n) m <<= 1;
    if (m > cap) m = cap;
    for (uint32_t i = n + g.tid; i < m; i += g.size) kbuf[i] = ~0ull;
    g.sync();
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = g.tid; i < m; i += g.size) {
                const uint32_t x = i ^ j;
                if (x > i) {
                    const unsigned long long u = kbuf[i], v = kbuf[x];
                    if ((u > v) == ((i & k) == 0)) {
                        kbuf[i] = v;
                        kbuf[x] = u;
                    }
                }
            }
            g.sync();
        }
}

// common.zig:153-166: walk the ranked candidates, at most k_eff, relative cutoff anchored on the best.
// kbuf[0..n) sorted and visible to the group.  All group threads call this; s_count is group-shared scratch.
__device__ void group_emit_results(const Group &g, const BatchArgs &a, const WorkItem &w,
                                   const unsigned long long *kbuf, uint32_t n, uint32_t *s_count) {
    const uint32_t lim = min(n, w.k_eff);
    uint32_t ms = w.min_score;
    if (lim > 0) {
        const uint32_t s0 = 0xFFFFFFFFu - (uint32_t)(kbuf[0] >> 32);
        ms = max(ms, (uint32_t)(s0 * w.min_score_pct) / 100u); // u32 wrapping product, truncating division
    }
    if (g.tid == 0) *s_count = 0;
    g.sync();
    for (uint32_t i = g.tid; i < lim; i += g.size) {
        const unsigned long long key = kbuf[i];
        const uint32_t score = 0xFFFFFFFFu - (uint32_t)(key >> 32);
        if (i == 0 || score >= ms) { // the best candidate is emitted before the cutoff is raised
            a.out_ids[(size_t)w.q * a.k_stride + i] = (uint32_t)key;
            a.out_scores[(size_t)w.q * a.k_stride + i] = score;
            atomicAdd(s_count, 1u);
        }
    }
    g.sync();
    if (g.tid == 0) {
        a.out_counts[w.q] = *s_count;
        if (a.stats) atomicAdd(&a.stats->results, (unsigned long long)*s_count);
    }
}

// ------------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Two flavours of waiting.  try_wait with a suspend-time hint parks the thread in hardware (no issue slots
// burnt, but the wake-up is not immediate); the plain form returns quickly and is polled.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t *bar, uint32_t parity, uint32_t ns) {
    uint32_t ok = 0;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, uint32_t mode = 0) {
    if (mode == 0) {
        while (!mbar_try_wait_hint(bar, parity, 1000000u)) {
        }
    } else if (mode == 1) {
        while (!mbar_try_wait(bar, parity)) {
        }
    } else {
        while (!mbar_try_wait(bar, parity)) __nanosleep(32);
    }
}
// global -> shared bulk copy (16-byte aligned, size multiple of 16), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int kSkResolverWarps = 4; // per resolver group
constexpr int kSkResolvers = kSkResolverWarps * 32;

// ------------------------------------------------------------------------------------------------
// sketch path, "count, then find" (the hot kernel; class kSketchClass).
// One persistent CTA per SM, warps in three roles, hand-overs by named barriers (arrive / sync pairs; the TMA
// completion is the only mbarrier), up to four queries in flight per SM:
//   producers  TMA bulk copies (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP) of the query's posting rows
//              into a ring of shared-memory stages; all producer warps fill one stage at a time.
//   counters   two groups, each with its own sketch of 32768 8-bit counters (four per 32-bit word), take the staged
//              queries in turn (even / odd).  A group adds every staged docid to its sketch with shared atomics whose
//              results are never looked at (SASS: ATOMS with RZ destination): a 128-bit load per four postings, then
//              per posting one hash multiply, four integer ops and the atomic; no branch, no dependence on the
//              shared-memory round trip.  h = docid * kRowMult; the counter is h's top 15 bits: word = h[31:19],
//              byte = h[18:17].  Then the group reads its sketch back once — all loads first: shared memory answers
//              slowly under the other group's atomics — which also clears it to the bias 128 - min_score of the
//              group's next query: "counter >= min_score" is then bit 7 of a byte, sixteen counters are tested with
//              two logic ops, and the few "hot" counters go to a list that travels with the stage.
//   resolvers  (groups of four warps, taking queries in turn) find the postings of the hot counters: rows are sorted by
//              row_key(docid) = h, whose top bits are the counter, so they are one contiguous range in every staged
//              row: a 9-ary search per (hot counter, row) finds them (a multiply and a compare per probe), the warps
//              list what they found, and one warp sums the list per docid — exactly — ranks and cuts like
//              common.zig:140-166.
// Why two counter groups: a sketch is busy from the first add until its read-back is done, and count + read-back of
// one query take longer than the other roles need per query; with one sketch per group the groups overlap.
// Exactness: a counter receives every posting whose docid maps to it, so a docid with count >= min_score makes
// its counter hot, and a hot counter's postings are all enumerated; scores never come from the sketch.  A byte
// that carries into its neighbour (128 + min_score arrivals at one counter) would corrupt the picture: the
// clearing pass also sums all bytes (IDP.4A), and the sum equals 32768 * bias + the number of staged postings
// iff no byte carried; otherwise, and when there are more hot counters or candidates than fit, the query is
// re-queued to the exact count-table kernels.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kHotCap = 32;   // hot counters per query handled here
constexpr uint32_t kFoundCap = 32; // (warp, docid) findings per query; more -> exact count-table path
// named barrier ids of search_find_kernel (0 is __syncthreads): counter groups 1..2, resolver groups 3..5, stage
// release 6..10, "counted" 11..15 (one per stage: a stage's barrier cannot be arrived at again before the resolvers
// that waited on it have released the stage)
constexpr uint32_t kFbCounters = 1, kFbGroup = 3, kFbStage = 6, kFbCounted = 11;

struct FindMeta { // one per stage: what travels with the staged query
    WorkItem item;
    uint16_t row_off4[kSketchMaxRows]; // first 16-byte granule of row r inside the stage (a stage is <= 64 KB)
    uint16_t row_len[kSketchMaxRows];  // postings in row r (without padding; the row fits the stage)
    uint32_t hot[kHotCap];            // counters that reached min_score: word * 4 + byte
    uint32_t n_hot, sum;              // sum: all bytes of the sketch after counting
};

struct FindState { // private to one resolver group
    uint32_t found_id[kFoundCap], found_cnt[kFoundCap]; // what the group's warps found: docid, number of postings
    uint32_t n_found;
};

// SKLOG: log2 of the sketch's 8-bit counters (15: 32 KB per sketch; 14: 16 KB, which leaves room for a fifth 32 KB
// stage — measured slower: the chance-hot counters of a C3 query go from 0.1 to 2.3 and each costs the resolvers)
template <int STAGES, uint32_t STAGE_U4, int SKLOG> constexpr size_t find_smem_bytes() {
    return 2 * ((size_t)1 << SKLOG) + (size_t)STAGES * STAGE_U4 * 16;
}

// TIMED: the instance whose phase timers (FPX_DEBUG_ABLATE bit 9) also see the barrier waits (settle() below).  This kernel
// is sensitive to the instruction schedule far beyond what the instruction counts suggest: the never-taken branch of a
// settle() per hand-over costs the default instance 6 %, and compiling the clock reads of the plain timers OUT of it costs
// 11 % (measured, profiles/r02/INDEX.md): they are compiler barriers that keep the loads of the next queries' work items,
// descriptors and sketch bias ahead of the hand-overs.  So the default instance keeps the clock reads and has no settle().
template <int GW, int RG, int PW, int STAGES, uint32_t STAGE_U4, int SKLOG, bool TIMED = false>
__global__ void __launch_bounds__((2 * GW + RG * kSkResolverWarps + PW) * 32, 1) search_find_kernel(BatchArgs a, uint32_t cls) {
    static_assert(RG >= 1 && RG <= 3 && STAGES >= 2 && STAGES <= 5 && SKLOG >= 13 && SKLOG <= 15, "barrier ids, sketch size");
    // Consecutive tenants of a stage must be taken by the same counter group and the same resolver group: the groups wait
    // on per-stage primitives (the parity of `full`, the "counted" barrier) that cannot tell one use of the stage from the
    // next, so a group that laps the other one (a slow copy behind cold TLBs is enough) must not meet it there.
    static_assert(STAGES % 2 == 0 && STAGES % RG == 0, "stage count: a multiple of the group counts");
    constexpr uint32_t kSketchBytes = 1u << SKLOG;            // one byte per counter
    constexpr uint32_t kSketchWords = kSketchBytes / 4;
    constexpr uint32_t kKeyShift = 32 - SKLOG;                // counter = h >> kKeyShift = word * 4 + byte
    constexpr uint32_t kWordMask = kSketchBytes - 4;          // byte offset of the counter's word
    constexpr int kFirstResolver = 2 * GW;
    constexpr int kFirstProducer = 2 * GW + RG * kSkResolverWarps;
    constexpr int kAllThreads = (kFirstProducer + PW) * 32;
    constexpr uint32_t kGroup = GW * 32; // counter threads per group
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *sketch_base = smem_raw; // one sketch per counter group
    uint4 *stage = reinterpret_cast<uint4 *>(smem_raw + 2 * (size_t)kSketchBytes);
    __shared__ uint64_t full[STAGES]; // TMA completion; every other hand-over is a named barrier
    __shared__ FindMeta meta[STAGES];
    __shared__ FindState fs[RG];
    __shared__ long long t_issued[TIMED ? STAGES : 1]; // TIMED: when producer warp 0 had issued its last copy of the stage's tenant

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t count = a.counters->qcount[cls];
    const WorkItem *items = a.items + (size_t)cls * a.n_queries;
    const uint32_t pad = a.snap.pad_id;
    const bool timed = (a.debug & 512u) && blockIdx.x == 0 && a.stats != nullptr;
    auto tick = [&](int slot, long long t0) {
        if (timed) atomicAdd(&a.stats->dbg[slot], (unsigned long long)(clock64() - t0));
    };
    // bar.sync does not block at issue but at the next instruction that consumes memory: a timer read right after it
    // would miss the wait, so the timed instance touches shared memory first
    auto settle = [&]() {
        if constexpr (TIMED) {
            if (timed) {
                uint32_t x;
                asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(x) : "r"(smem_u32(&full[0])) : "memory");
            }
        }
    };
    if (count == 0u) return;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], PW); // every producer warp arrives with its share of the bytes
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // each group's sketch starts at the bias of the group's first query
    for (uint32_t g = 0; g < 2; ++g) {
        const unsigned long long idx = blockIdx.x + (unsigned long long)g * gridDim.x;
        const uint32_t ms = idx < count ? items[idx].min_score : 2u;
        const uint32_t bias = (0x80u - ms) * 0x01010101u;
        uint4 *sk4 = reinterpret_cast<uint4 *>(sketch_base + (size_t)g * kSketchBytes);
        for (uint32_t i = tid; i < kSketchWords / 4; i += kAllThreads) sk4[i] = make_uint4(bias, bias, bias, bias);
    }
    if (tid < RG) fs[tid].n_found = 0u;
    if (tid < STAGES) meta[tid].n_hot = meta[tid].sum = 0u;
    __syncthreads();

    if (warp >= kFirstProducer) {
        // ===== producers: all producer warps fill one stage at a time, query after query (stage = it % STAGES);
        // warp p issues rows p, p + P, p + 2P, ... (lane l holds row P*l + p).  A bulk copy costs ~10 instructions
        // and ~80 cycles of a warp, so a query's ~100 copies are spread over every producer warp.
        constexpr int kP = PW, kDesc = (kSketchMaxRows + 32 * kP - 1) / (32 * kP);
        const uint32_t p = warp - kFirstProducer;
        const uint4 *docids4 = reinterpret_cast<const uint4 *>(a.snap.docids);
        auto item_at = [&](uint32_t it, WorkItem &w) -> bool {
            const unsigned long long idx = blockIdx.x + (unsigned long long)it * gridDim.x;
            if (idx >= count) return false;
            w = items[idx];
            return true;
        };
        auto rows_of = [&](const WorkItem &w, uint4 (&d)[kDesc]) {
#pragma unroll
            for (int j = 0; j < kDesc; ++j) {
                const uint32_t r = (uint32_t)kP * (lane + 32 * j) + p;
                d[j] = r < w.n_rows ? a.rows[w.rows_off + r] : make_uint4(0u, 0u, 0u, 0u);
            }
        };
        // Row descriptors and work items are read two and three queries ahead: they come from DRAM (the batch's
        // descriptors are larger than L2) behind the copies' own traffic, ~1.5 us away, and a load that is waited for at
        // the top of the loop pins the whole kernel to that latency.
        WorkItem w{}, w1{}, w2{}, w3{};
        uint4 d[kDesc], d1[kDesc], d2[kDesc];
        bool have = item_at(0, w);
        if (have) rows_of(w, d);
        bool have1 = item_at(1, w1);
        if (have1) rows_of(w1, d1);
        bool have2 = item_at(2, w2);
        for (uint32_t it = 0; have; ++it) {
            const uint32_t s = it % STAGES;
            uint4 *dst = stage + (size_t)s * STAGE_U4;
            if (have2) rows_of(w2, d2);
            const bool have3 = item_at(it + 3, w3);
            const long long tp0 = clock64();
            if (it >= (uint32_t)STAGES) // the resolvers released the previous tenant of this stage: us + their warp 0
                named_sync(kFbStage + s, 32 * kP + 32);
            settle();
            if (p == 0 && lane == 0) tick(0, tp0);
            uint32_t mine = 0;
#pragma unroll
            for (int j = 0; j < kDesc; ++j) mine += (d[j].y + 3) >> 2;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
#pragma unroll
            for (int j = 0; j < kDesc; ++j) { // stage directory for the resolvers (d.z: the row's place, from prepare_kernel)
                const uint32_t r = (uint32_t)kP * (lane + 32 * j) + p;
                if (r < kSketchMaxRows) {
                    meta[s].row_off4[r] = (uint16_t)d[j].z;
                    meta[s].row_len[r] = (uint16_t)d[j].y;
                }
            }
            if (p == 0 && lane == 0) meta[s].item = w;
            if (a.debug & 8u) mine = 0;
            __syncwarp();
            if (lane == 0) mbar_expect_tx(&full[s], mine * 16u); // expect_tx precedes my copies (release)
            __syncwarp();
            long long ti0 = 0;
            if constexpr (TIMED) ti0 = clock64();
            if (!(a.debug & 8u)) {
#pragma unroll
                for (int j = 0; j < kDesc; ++j)
                    if (d[j].y) bulk_g2s(dst + d[j].z, docids4 + d[j].x, ((d[j].y + 3) >> 2) * 16u, &full[s]);
            }
            if constexpr (TIMED) {
                if (p == 0 && lane == 0) {
                    tick(14, ti0); // issuing my share of the query's copies
                    t_issued[s] = clock64();
                }
            }
            have = have1;
            have1 = have2;
            have2 = have3;
            w = w1;
            w1 = w2;
            w2 = w3;
#pragma unroll
            for (int j = 0; j < kDesc; ++j) {
                d[j] = d1[j];
                d1[j] = d2[j];
            }
            if (p == 0 && lane == 0) {
                tick(1, tp0);
                if (timed) atomicAdd(&a.stats->dbg[2], 1ull);
            }
        }
        return;
    }

    if (warp >= kFirstResolver) {
        // ===== resolvers: group gidx takes every RG-th query; query `it` used sketch it & 1
        const uint32_t gidx = (warp - kFirstResolver) / kSkResolverWarps;
        const uint32_t rwarp = (warp - kFirstResolver) % kSkResolverWarps;
        const uint32_t rtid = rwarp * 32 + lane;
        const Group R{rtid, (uint32_t)kSkResolvers, kFbGroup + gidx};
        FindState &st = fs[gidx];
        for (uint32_t it = gidx;; it += RG) {
            const unsigned long long idx = blockIdx.x + (unsigned long long)it * gridDim.x;
            if (idx >= count) break;
            const uint32_t s = it % STAGES;
            FindMeta &m = meta[s];
            const long long tr0 = clock64();
            named_sync(kFbCounted + s, kGroup + kSkResolvers); // a counter group has counted query it and read its sketch back
            settle();
            if (gidx == 0 && rtid == 0) tick(3, tr0);
            const WorkItem w = m.item;
            const uint32_t n_hot = (a.debug & 2u) ? 0u : m.n_hot;
            // no byte carried <=> the bytes add up to the bias of every counter plus one per staged posting
            bool redo = m.sum != kSketchBytes * (0x80u - w.min_score) + 4u * w.total4 && !(a.debug & 3u);
            redo = redo || n_hot > kHotCap;
            if (n_hot != 0u && !redo) {
                // Find the postings of the hot counters: thread rtid owns row rtid of the stage (<= 128 rows); the
                // postings of counter c are the range of row keys whose top bits are c.  Every dependent shared-memory
                // round trip costs hundreds of cycles while the counters' atomics fill the pipe, and the stage is held
                // all the while, so the lower bound is a 9-ary search (eight independent probes per round, two rounds
                // for a row of <= 80, the last round's values stay in registers) and a warp reports what it found with
                // one append to a list, not a probe chain into a table.
                const bool has_row = rtid < w.n_rows;
                const uint32_t *row = reinterpret_cast<const uint32_t *>(stage + (size_t)s * STAGE_U4) +
                                      (has_row ? 4u * m.row_off4[rtid] : 0u);
                const uint32_t len = has_row ? m.row_len[rtid] : 0u;
#pragma unroll 1
                for (uint32_t c = 0; c < n_hot; ++c) {
                    const uint32_t kp = m.hot[c] << kKeyShift; // first row key of the counter
                    uint32_t lo = 0, hi = len; // every key before lo is below kp, every key from hi on is not
                    while (hi - lo > 8u) {
                        const uint32_t step = (hi - lo) / 9u + 1u;
                        uint32_t below = 0;
#pragma unroll
                        for (uint32_t j = 1; j <= 8; ++j) {
                            const uint32_t pj = lo + j * step - 1u;
                            if (pj < hi && row[pj] * kRowMult < kp) ++below; // the pivots below kp form a prefix
                        }
                        const uint32_t nxt = lo + (below + 1u) * step - 1u; // first pivot not below kp, if there is one
                        if (below < 8u && nxt < hi) hi = nxt;
                        lo += below * step;
                    }
                    uint32_t e[9], below = 0; // the window [lo, hi] (hi - lo <= 8): the lower bound lies in it
#pragma unroll
                    for (uint32_t j = 0; j < 9; ++j) {
                        e[j] = lo + j < len ? row[lo + j] : 0u;
                        if (lo + j < hi && e[j] * kRowMult < kp) ++below;
                    }
                    uint32_t d = e[0];
#pragma unroll
                    for (uint32_t j = 1; j < 9; ++j) d = below == j ? e[j] : d;
                    uint32_t pos = lo + below;
                    const bool has = pos < len && ((d * kRowMult ^ kp) >> kKeyShift) == 0u;
                    // a true match is found in most rows: one list entry per warp and docid, not one per row
                    uint32_t act = __ballot_sync(0xFFFFFFFFu, has);
                    while (act) {
                        const uint32_t leader = __ffs(act) - 1u;
                        const uint32_t d0 = __shfl_sync(0xFFFFFFFFu, d, leader);
                        const uint32_t same = __ballot_sync(0xFFFFFFFFu, has && d == d0);
                        if (lane == leader) {
                            const uint32_t at = atomicAdd(&st.n_found, 1u);
                            if (at < kFoundCap) {
                                st.found_id[at] = d0;
                                st.found_cnt[at] = __popc(same);
                            }
                        }
                        act &= ~same;
                    }
                    // further postings of this counter in my row: repeated (hash, id) pairs, other docids
                    if (has) {
                        for (++pos; pos < len; ++pos) {
                            d = row[pos];
                            if (((d * kRowMult ^ kp) >> kKeyShift) != 0u) break;
                            const uint32_t at = atomicAdd(&st.n_found, 1u);
                            if (at < kFoundCap) {
                                st.found_id[at] = d;
                                st.found_cnt[at] = 1u;
                            }
                        }
                    }
                }
            }
            R.sync(); // everyone is done with the stage; the findings are complete
            settle();
            if (gidx == 0 && rtid == 0) tick(11, tr0);
            if (rwarp == 0) {
                if (lane == 0) m.n_hot = m.sum = 0u;      // for the stage's next tenant
                named_arrive(kFbStage + s, 32 * PW + 32); // the stage goes back to the producers
                const uint32_t n_found = n_hot != 0u && !redo ? st.n_found : 0u;
                redo = redo || n_found > kFoundCap;
                // lane i takes finding i: sum the findings of its docid, keep the first of each docid with
                // score >= min_score (common.zig:140-145), rank and cut in registers (common.zig:147-166)
                const uint32_t d = lane < n_found ? st.found_id[lane] : pad, sc = lane < n_found ? st.found_cnt[lane] : 0u;
                __syncwarp();
                if (lane == 0) st.n_found = 0u;
                uint32_t n_out = 0;
                if (!redo && n_found != 0u) {
                    uint32_t total = 0;
                    bool first = true;
                    for (uint32_t j = 0; j < n_found; ++j) {
                        const uint32_t dj = __shfl_sync(0xFFFFFFFFu, d, j), cj = __shfl_sync(0xFFFFFFFFu, sc, j);
                        if (dj == d) {
                            total += cj;
                            if (j < lane) first = false;
                        }
                    }
                    const bool keep = lane < n_found && first && total >= w.min_score;
                    const unsigned long long key = keep ? rank_key(total, d) : ~0ull;
                    uint32_t rank = 0;
                    for (uint32_t j = 0; j < n_found; ++j) rank += (__shfl_sync(0xFFFFFFFFu, key, j) < key) ? 1u : 0u;
                    const uint32_t best = __reduce_max_sync(0xFFFFFFFFu, keep ? total : 0u);
                    // u32 wrapping product, truncating division; the best is emitted before the cutoff is raised
                    const uint32_t ms = max(w.min_score, (uint32_t)(best * w.min_score_pct) / 100u);
                    const bool emit = keep && rank < w.k_eff && (rank == 0u || total >= ms);
                    if (emit) {
                        a.out_ids[(size_t)w.q * a.k_stride + rank] = d;
                        a.out_scores[(size_t)w.q * a.k_stride + rank] = total;
                    }
                    n_out = __popc(__ballot_sync(0xFFFFFFFFu, emit));
                }
                if (lane == 0) {
                    if (redo) { // not decidable here: the exact count-table kernels take the query
                        enqueue(a, exact_class_for(w.postings, w.k_eff), w);
                        if (a.stats) atomicAdd(&a.stats->overflow_requeues, 1ull);
                    } else {
                        a.out_counts[w.q] = n_out;
                        if (a.stats) {
                            atomicAdd(&a.stats->results, (unsigned long long)n_out);
                            atomicAdd(&a.stats->sketch_queries, 1ull);
                        }
                    }
                }
            }
            R.sync(); // the group's scratch is reused by its next query
            if (gidx == 0 && rtid == 0) {
                tick(5, tr0);
                if (timed) atomicAdd(&a.stats->dbg[6], 1ull);
            }
        }
        return;
    }

    // ===== counters: group g = warps [g * GW, (g + 1) * GW) takes queries it = g, g + 2, ... with sketch g: count, then
    // (the whole group together) read the sketch back: byte sum, hot counters, clear to the bias of the sketch's next
    // tenant, this CTA's query it + 2
    const uint32_t g = warp / GW, gwarp = warp - g * GW, gtid = gwarp * 32 + lane;
    const uint32_t boff = g * kSketchBytes; // folded into the address by the same LOP3 that masks the word offset
    uint4 *sk4 = reinterpret_cast<uint4 *>(sketch_base + boff);
    constexpr int kScan = (int)((kSketchWords / 4 + kGroup - 1) / kGroup); // 16-byte pieces of the sketch per thread
    for (uint32_t it = g;; it += 2) {
        const unsigned long long idx = blockIdx.x + (unsigned long long)it * gridDim.x;
        if (idx >= count) break;
        const uint32_t s = it % STAGES;
        const unsigned long long idx2 = idx + 2ull * gridDim.x;
        const uint32_t ms2 = idx2 < count ? items[idx2].min_score : 2u; // from DRAM; needed at the read-back, issued here
        const long long tc0 = clock64(); // (also a compiler barrier: the load above stays above the wait below)
        if (gwarp == 0) { // one warp polls the TMA completion, the others park on the named barrier
            if (lane == 0) {
                mbar_wait(&full[s], (it / STAGES) & 1, 0);
                if (g == 0) tick(7, tc0);
                if constexpr (TIMED)
                    if (g == 0) tick(15, t_issued[s]); // from producer warp 0's last copy to the stage being complete (0 if it was)
            }
            __syncwarp();
        }
        named_sync(kFbCounters + g, kGroup);
        settle();
        if constexpr (TIMED)
            if (g == 0 && gtid == 32) tick(12, tc0); // a warp that parks on the barrier while warp 0 polls
        FindMeta &m = meta[s];
        const uint32_t total4 = m.item.total4;
        const uint4 *sg = stage + (size_t)s * STAGE_U4;
        // Row padding is made of unused docids spread over many values: counted like anything else (the byte sum
        // expects it), never found by the resolvers (they search the true row lengths).
        auto add4 = [&](const uint4 v) {
            const uint32_t dd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t t = (dd[e] * kMult) >> kKeyShift; // counter: word = t >> 2, byte = t & 3
                atomicAdd(reinterpret_cast<uint32_t *>(sketch_base + ((t & kWordMask) | boff)), __funnelshift_l(0u, 1u, t << 3));
            }
        };
        if (!(a.debug & 1u)) {
            uint32_t i = gtid;
            for (; i + 3 * kGroup < total4; i += 4 * kGroup) { // four loads in flight, then sixteen adds
                const uint4 v0 = sg[i], v1 = sg[i + kGroup], v2 = sg[i + 2 * kGroup], v3 = sg[i + 3 * kGroup];
                add4(v0);
                add4(v1);
                add4(v2);
                add4(v3);
            }
            for (; i < total4; i += kGroup) add4(sg[i]);
        }
        if (g == 0 && gtid == 0) tick(8, tc0);
        named_sync(kFbCounters + g, kGroup); // every warp of the group is done counting query it
        settle();
        if constexpr (TIMED)
            if (g == 0 && gtid == 0) tick(13, tc0);
        {
            const uint32_t b2 = (0x80u - ms2) * 0x01010101u;
            const uint4 clear4 = make_uint4(b2, b2, b2, b2);
            uint32_t acc = 0;
#pragma unroll 1
            for (int r0 = 0; r0 < kScan; r0 += 8) { // rounds of up to eight loads, all of them before anything else
                uint4 v[8];
#pragma unroll
                for (int k8 = 0; k8 < 8; ++k8) {
                    const uint32_t i = gtid + (uint32_t)(r0 + k8) * kGroup;
                    v[k8] = (r0 + k8 < kScan && i < kSketchWords / 4) ? sk4[i] : make_uint4(0u, 0u, 0u, 0u);
                }
#pragma unroll
                for (int k8 = 0; k8 < 8; ++k8) {
                    const uint32_t i = gtid + (uint32_t)(r0 + k8) * kGroup;
                    if (r0 + k8 < kScan && i < kSketchWords / 4 && !(a.debug & 16u)) sk4[i] = clear4;
                    acc = __dp4a(v[k8].x, 0x01010101u, acc);
                    acc = __dp4a(v[k8].y, 0x01010101u, acc);
                    acc = __dp4a(v[k8].z, 0x01010101u, acc);
                    acc = __dp4a(v[k8].w, 0x01010101u, acc);
                    if ((v[k8].x | v[k8].y | v[k8].z | v[k8].w) & 0x80808080u) { // rare: a counter at bias + min_score or more
#pragma unroll 1
                        for (uint32_t e = 0; e < 4; ++e) {
                            uint32_t hm = (e == 0 ? v[k8].x : e == 1 ? v[k8].y : e == 2 ? v[k8].z : v[k8].w) & 0x80808080u;
                            while (hm) {
                                const uint32_t bit = __ffs(hm) - 1u;
                                hm &= hm - 1u;
                                const uint32_t pos = atomicAdd(&m.n_hot, 1u);
                                if (pos < kHotCap) m.hot[pos] = ((i * 4u + e) << 2) | (bit >> 3); // word * 4 + byte
                            }
                        }
                    }
                }
            }
            acc = __reduce_add_sync(0xFFFFFFFFu, acc);
            if (lane == 0) atomicAdd(&m.sum, acc);
        }
        __syncwarp();
        named_arrive(kFbCounted + s, kGroup + kSkResolvers); // query it is counted and its sketch read back
        if (g == 0 && gtid == 0) {
            tick(9, tc0);
            if (timed) atomicAdd(&a.stats->dbg[10], 1ull);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// exact shared-memory path (classes 1..3)
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
// `rot`: multi-pass queries split the key space by the top `rot` bits of hv (one key range per pass, see
// search_smem_kernel), so those bits are the same for every docid of a pass: the key is rotated left by `rot` first.
// PEEK: read the slot first and only CAS when it is empty.  A slot's tag never changes once set, so a stale read is
// harmless; for queries whose docids repeat a lot (hot, capped rows share their low docids) most inserts find their
// docid already there and become one load + one add instead of a CAS (half the atomic rate) + an add.
template <int LOG, bool PEEK = false>
__device__ __forceinline__ void table_insert(uint32_t *tab, uint32_t d, uint32_t *ovf, uint32_t rot = 0) {
    using P = Packed<LOG>;
    const uint32_t hv = __funnelshift_l(d * kMult, d * kMult, rot);
    const uint32_t rem = hv & P::kRemMask;
    const uint32_t step = ((rem << 1) | 1u) & P::kSlotMask;
    const uint32_t tagbase = rem << (P::kProbeBits + P::kCntBits);
    uint32_t s = hv >> P::kRemBits;
#pragma unroll 1
    for (uint32_t i = 0; i < 32; ++i) {
        const uint32_t tag = tagbase | (i << P::kCntBits);
        uint32_t old = PEEK ? *const_cast<volatile uint32_t *>(tab + s) : 0u;
        if (old == 0u) old = atomicCAS(tab + s, 0u, tag | 1u);
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

template <int LOG> __device__ __forceinline__ uint32_t table_docid(uint32_t word, uint32_t slot, uint32_t rot = 0) {
    using P = Packed<LOG>;
    const uint32_t rem = word >> (P::kProbeBits + P::kCntBits);
    const uint32_t probe = (word >> P::kCntBits) & 31u;
    const uint32_t step = ((rem << 1) | 1u) & P::kSlotMask;
    const uint32_t home = (slot - probe * step) & P::kSlotMask;
    const uint32_t hv = (home << P::kRemBits) | rem;
    return __funnelshift_r(hv, hv, rot) * kMultInv;
}

constexpr uint32_t kSmemRowsChunk = 128; // row descriptors staged per round by the shared-memory kernels
template <int THREADS> struct SmemKbuf { static constexpr uint32_t kCap = THREADS <= 512 ? 1024u : 2048u; }; // >= kFastKbuf + THREADS
template <int LOG, int THREADS> constexpr size_t smem_bytes_for() {
    return (size_t)Packed<LOG>::kSlots * 4 + (size_t)SmemKbuf<THREADS>::kCap * 8 + kSmemRowsChunk * 16;
}

// THREADS: 256 for the two small tables (4 and 3 CTAs per SM); the 128 KB table leaves room for one CTA per SM
// only, which then runs 1024 threads to keep enough loads in flight.
template <int LOG, int THREADS>
__global__ void __launch_bounds__(THREADS, (LOG == 13 ? 4 : (LOG == 14 ? 3 : 1))) search_smem_kernel(BatchArgs a) {
    constexpr int kThreads = THREADS, kWarps = THREADS / 32; // shadow the file-level constants
    using P = Packed<LOG>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem_raw);
    // candidate buffer: the up to kFastKbuf best so far + one scan round (one slot per thread); a power of two
    constexpr uint32_t kKbufCap = SmemKbuf<THREADS>::kCap;
    constexpr uint32_t kRows = kSmemRowsChunk;
    unsigned long long *kbuf = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)P::kSlots * 4);
    uint4 *rows_s = reinterpret_cast<uint4 *>(smem_raw + (size_t)P::kSlots * 4 + kKbufCap * 8);
    __shared__ uint32_t s_idx, s_ncand, s_ovf, s_count, s_qual, s_bar, s_next;
    __shared__ uint32_t s_cur[kSmemRowsChunk]; // multi-pass: per row, the first granule of the next pass
    // Histogram of the scores of a pass's qualifying docs, 256 bins (the last one: 255 or more).  It lives in the row
    // descriptors' place, which is free between the insert phase and the next one; bin b sits at b + b / 32, so lane l of
    // warp 0 walks its bins 8l .. 8l + 7 without bank conflicts.
    constexpr uint32_t kHist = 256u, kHistPer = kHist / 32u;
    static_assert((kHist + kHist / 32u) * 4u <= kSmemRowsChunk * 16u, "the histogram fits the row descriptors' place");
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(rows_s);
    auto hbin = [](uint32_t b) { return b + (b >> 5); };
    __shared__ unsigned long long s_kth; // running threshold: only keys below it can still make the top k_eff

    constexpr int cls = LOG - 12;
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const Group g{tid, (uint32_t)kThreads, 0u};
    const uint4 *docids4 = reinterpret_cast<const uint4 *>(a.snap.docids);
    uint4 *tab4 = reinterpret_cast<uint4 *>(tab);

    const uint32_t qcount = a.counters->qcount[cls];
    if (qcount == 0u) return; // nothing queued for this class (the usual case): do not even clear the table
    for (uint32_t i = tid; i < P::kSlots / 4; i += kThreads) tab4[i] = make_uint4(0, 0, 0, 0);

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_idx = atomicAdd(&a.counters->qhead[cls], 1u);
            s_ncand = 0;
            s_ovf = 0;
            s_qual = 0;
            s_kth = ~0ull;
        }
        __syncthreads();
        if (s_idx >= qcount) break;
        const WorkItem w = a.items[(size_t)cls * a.n_queries + s_idx];
        const uint32_t thr = max(w.min_score, 1u); // a doc in the table has score >= 1
        const uint32_t k_eff = min(w.k_eff, kFastKbuf);
        const uint4 *rows = a.rows + w.rows_off;
        // More postings than the table holds: several passes, pass p counting the docids whose key h = docid * kMult
        // has the top log2(passes) bits p.  Rows are sorted by h (fpx_kernels.cuh: row_key), so a pass's postings are one
        // contiguous piece of every row and each pass reads only its piece: the warp that owns a row keeps a cursor.
        uint32_t passes = 1, rot = 0;
        if (LOG == 15)
            while ((unsigned long long)passes * 20480ull < w.postings && w.postings > 16384u) { // load <= 0.63 if all differ
                passes <<= 1;
                ++rot;
            }
        const bool keep_cursors = w.n_rows <= kRows; // one chunk of rows: the cursors survive from pass to pass

        for (uint32_t pass = 0; pass < passes; ++pass) {
            for (uint32_t r0 = 0; r0 < w.n_rows; r0 += kRows) {
                const uint32_t nr = min(kRows, w.n_rows - r0);
                __syncthreads();
                for (uint32_t i = tid; i < nr; i += kThreads) rows_s[i] = rows[r0 + i];
                if (tid == 0) s_next = 0u;
                __syncthreads();
                if (passes == 1) {
                    // a warp per posting row, 128-bit loads, two rows in flight; the tail of a row's last
                    // 16-byte granule is padding and is masked by position
                    for (uint32_t r = warp * 2; r < nr; r += kWarps * 2) {
                        const uint4 ra = rows_s[r];
                        const uint4 rb = (r + 1 < nr) ? rows_s[r + 1] : make_uint4(0u, 0u, 0u, 0u);
                        const uint32_t na = (ra.y + 3) >> 2, nb = (rb.y + 3) >> 2;
                        const uint32_t nmax = max(na, nb);
                        for (uint32_t i = lane; i < nmax; i += 32) {
                            uint4 va = make_uint4(0, 0, 0, 0), vb = va;
                            if (i < na) va = __ldg(docids4 + ra.x + i);
                            if (i < nb) vb = __ldg(docids4 + rb.x + i);
                            const uint32_t d[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                if (4 * i + (e & 3) >= (e < 4 ? ra.y : rb.y)) continue;
                                table_insert<LOG>(tab, d[e], &s_ovf);
                            }
                        }
                    }
                } else {
                    // Rows are handed out one at a time (a pass's pieces differ in length, and 100 rows over 32 warps
                    // leave a quarter of the warps idle in the last round of a fixed split).
                    const uint32_t shift = 32 - rot;
                    for (;;) {
                        uint32_t r = 0;
                        if (lane == 0) r = atomicAdd(&s_next, 1u);
                        r = __shfl_sync(0xFFFFFFFFu, r, 0);
                        if (r >= nr) break;
                        const uint4 ra = rows_s[r];
                        const uint32_t na = (ra.y + 3) >> 2;
                        uint32_t g0 = 0; // first granule that can hold a posting of this pass
                        if (pass != 0) {
                            if (keep_cursors) {
                                g0 = s_cur[r];
                            } else { // many rows: no room for cursors, search the granule (first key >= this pass's range)
                                uint32_t lo = 0, hi = na;
                                while (lo < hi) {
                                    const uint32_t mid = (lo + hi) >> 1;
                                    const uint32_t last = min(4 * mid + 3, ra.y - 1); // last real posting of the granule
                                    if (((__ldg(a.snap.docids + (size_t)ra.x * 4 + last) * kMult) >> shift) < pass) lo = mid + 1; else hi = mid;
                                }
                                g0 = lo;
                            }
                        }
                        uint32_t next = na;
                        for (uint32_t i0 = g0; i0 < na; i0 += 32) {
                            const uint32_t i = i0 + lane;
                            uint4 va = make_uint4(0, 0, 0, 0);
                            if (i < na) va = __ldg(docids4 + ra.x + i);
                            const uint32_t d[4] = {va.x, va.y, va.z, va.w};
                            bool past = false; // a real posting of my granule belongs to a later pass
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if (i >= na || 4 * i + e >= ra.y) continue;
                                const uint32_t part = (d[e] * kMult) >> shift;
                                if (part == pass) table_insert<LOG, true>(tab, d[e], &s_ovf, rot);
                                past = past || part > pass;
                            }
                            const uint32_t pm = __ballot_sync(0xFFFFFFFFu, past);
                            if (pm) { // the next pass starts at the first granule that reaches beyond this one
                                next = i0 + (__ffs(pm) - 1);
                                break;
                            }
                        }
                        if (keep_cursors && lane == 0) s_cur[r] = next;
                    }
                }
            }
            __syncthreads();
            // scan + clear; candidates are docs with score >= max(min_score,1)  (common.zig:140-145) whose key
            // can still make the top k_eff.  First a histogram of their scores ...
            // (skipped when the candidates of this pass fit for sure: at most postings / thr docs can reach thr)
            unsigned long long kth = s_kth;
            const bool sure = passes == 1 && w.postings / thr <= kKbufCap;
            if (!sure) {
                for (uint32_t i = tid; i < kHist + kHist / 32u; i += kThreads) s_hist[i] = 0u;
                __syncthreads();
            }
            for (uint32_t i = tid; i < P::kSlots / 4 && !sure; i += kThreads) {
                const uint4 wd = tab4[i];
                if ((wd.x | wd.y | wd.z | wd.w) == 0u) continue;
                const uint32_t ws[4] = {wd.x, wd.y, wd.z, wd.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t cnt = ws[e] & P::kCntMask;
                    if (ws[e] != 0u && cnt >= thr && rank_key(cnt, table_docid<LOG>(ws[e], i * 4 + e, rot)) < kth)
                        atomicAdd(&s_hist[hbin(min(cnt, kHist - 1u))], 1u);
                }
            }
            __syncthreads();
            // ... from which one warp derives the bar of this pass: only the k_eff best of a pass can reach the final top
            // k_eff, so with more qualifying docs than that the pass keeps the scores >= bar, bar = the largest score
            // that still leaves k_eff of them (hot, capped rows share their low docids: a Zipf query has thousands of
            // docs above the floor, and ranking them all was most of such a query's time)
            if (warp == 0 && !sure) {
                uint32_t loc = 0;
                for (uint32_t j = 0; j < kHistPer; ++j) loc += s_hist[hbin(lane * kHistPer + j)];
                uint32_t suf = loc; // docs with a score in my bins or above
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_down_sync(0xFFFFFFFFu, suf, o);
                    if (lane + o < 32) suf += y;
                }
                uint32_t bar = thr, qual = __shfl_sync(0xFFFFFFFFu, suf, 0);
                if (qual > k_eff) {
                    const uint32_t at = 31u - __clz(__ballot_sync(0xFFFFFFFFu, suf >= k_eff)); // lane 0 always votes
                    uint32_t acc = suf - loc, j = kHistPer;
                    if (lane == at)
                        while (j-- > 0u) {
                            acc += s_hist[hbin(lane * kHistPer + j)];
                            if (acc >= k_eff) break;
                        }
                    bar = max(thr, __shfl_sync(0xFFFFFFFFu, lane * kHistPer + j, at));
                    qual = __shfl_sync(0xFFFFFFFFu, acc, at);
                }
                if (lane == 0) {
                    s_qual = qual;
                    s_bar = bar;
                }
            }
            __syncthreads();
            const uint32_t qual = sure ? 0u : s_qual, bar = sure ? thr : s_bar;
            uint32_t have = s_ncand;
            __syncthreads(); // everyone has read `have` before the first thread appends (or the branches below could differ)
            auto shrink = [&](uint32_t n) { // keep the k_eff best of kbuf[0..n) and raise the running threshold
                group_sort_keys(g, kbuf, n, kKbufCap);
                __syncthreads();
                if (tid == 0) {
                    s_ncand = min(n, k_eff);
                    if (n >= k_eff && k_eff > 0) s_kth = kbuf[k_eff - 1];
                }
                __syncthreads();
            };
            if (have + qual > kKbufCap && have > k_eff) { // earlier passes' candidates are in the way
                shrink(have);
                have = s_ncand;
                kth = s_kth;
            }
            if (have + qual <= kKbufCap) {
                // ... the usual case: they all fit
                for (uint32_t i = tid; i < P::kSlots / 4; i += kThreads) {
                    const uint4 wd = tab4[i];
                    if ((wd.x | wd.y | wd.z | wd.w) == 0u) continue;
                    tab4[i] = make_uint4(0, 0, 0, 0);
                    const uint32_t ws[4] = {wd.x, wd.y, wd.z, wd.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t cnt = ws[e] & P::kCntMask;
                        if (ws[e] != 0u && cnt >= bar) {
                            const unsigned long long key = rank_key(cnt, table_docid<LOG>(ws[e], i * 4 + e, rot));
                            if (key < kth) kbuf[atomicAdd(&s_ncand, 1u)] = key;
                        }
                    }
                }
            } else {
                // ... more docs tie at the bar than the buffer holds: one slot per thread and round; whenever another
                // round might not fit, keep the k_eff best and raise the running threshold
                if (have + kThreads > kKbufCap) shrink(have);
                for (uint32_t base = 0; base < P::kSlots; base += kThreads) {
                    const uint32_t slot = base + tid;
                    const uint32_t wv = tab[slot];
                    tab[slot] = 0u;
                    const uint32_t cnt = wv & P::kCntMask;
                    if (wv != 0u && cnt >= bar) {
                        const unsigned long long key = rank_key(cnt, table_docid<LOG>(wv, slot, rot));
                        if (key < s_kth) kbuf[atomicAdd(&s_ncand, 1u)] = key;
                    }
                    __syncthreads();
                    const uint32_t n = s_ncand;
                    __syncthreads(); // everyone has read n before the next round appends (or the branch could differ)
                    if (n + kThreads > kKbufCap) shrink(n);
                }
            }
        }
        __syncthreads();
        const uint32_t n = s_ncand;
        if (s_ovf) {
            // not representable here (count or probe field overflow): the global-memory path takes the query
            if (tid == 0) {
                enqueue(a, kWideClass, w);
                if (a.stats) atomicAdd(&a.stats->overflow_requeues, 1ull);
            }
            continue;
        }
        group_sort_keys(g, kbuf, n, kKbufCap);
        group_emit_results(g, a, w, kbuf, n, &s_count);
    }
}

// ------------------------------------------------------------------------------------------------
// global-memory path: 64-bit slots (docid << 32 | count), linear probing, any size via hash partitions
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) search_wide_kernel(BatchArgs a) {
    __shared__ unsigned long long kbuf[kWideKbuf];
    __shared__ uint4 rows_s[kRowsChunk];
    __shared__ uint32_t s_idx, s_kn, s_new, s_fail, s_count;
    __shared__ unsigned long long s_kth;

    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const Group g{tid, (uint32_t)kThreads, 0u};
    unsigned long long *table = a.wide_tables + ((size_t)blockIdx.x << a.wide_cap_log2);
    const uint4 *docids4 = reinterpret_cast<const uint4 *>(a.snap.docids);
    const uint32_t qcount = a.counters->qcount[kWideClass];
    if (qcount == 0u) return;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_idx = atomicAdd(&a.counters->qhead[kWideClass], 1u);
        __syncthreads();
        if (s_idx >= qcount) break;
        const WorkItem w = a.items[(size_t)kWideClass * a.n_queries + s_idx];
        const uint32_t thr = max(w.min_score, 1u);
        const uint32_t k_eff = min(w.k_eff, kMaxResults);
        const uint4 *rows = a.rows + w.rows_off;
        if (tid == 0 && a.stats) atomicAdd(&a.stats->wide_queries, 1ull);

        // table size: aim for load <= 0.25; more hash partitions when one table cannot hold that
        const unsigned long long need = 4ull * w.postings;
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
                for (uint32_t r0 = 0; r0 < w.n_rows; r0 += kRowsChunk) {
                    const uint32_t nr = min(kRowsChunk, w.n_rows - r0);
                    __syncthreads();
                    for (uint32_t i = tid; i < nr; i += kThreads) rows_s[i] = rows[r0 + i];
                    __syncthreads();
                    for (uint32_t r = warp; r < nr; r += kWarps) {
                        const uint4 ra = rows_s[r];
                        const uint32_t na = (ra.y + 3) >> 2;
                        for (uint32_t i = lane; i < na; i += 32) {
                            const uint4 va = __ldg(docids4 + ra.x + i);
                            const uint32_t d[4] = {va.x, va.y, va.z, va.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const uint32_t id = d[e];
                                if (4 * i + e >= ra.y) continue; // padding of the row's last granule
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
                    const uint32_t n = s_kn;
                    __syncthreads(); // everyone has read n before the next round appends (or the branch could differ)
                    if (n > kWideKbuf - 4 * kThreads) {
                        group_sort_keys(g, kbuf, n, kWideKbuf);
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
                        a.out_counts[w.q] = 0;
                    }
                    done = true;
                    break;
                }
                passes <<= 1;
                continue;
            }
            __syncthreads();
            const uint32_t n = s_kn;
            group_sort_keys(g, kbuf, n, kWideKbuf);
            __syncthreads();
            group_emit_results(g, a, w, kbuf, n, &s_count);
            done = true;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// result packing for host batches: most queries return one or two results, so the k_stride-wide device
// arrays are packed to {count per query, (id, score) pairs back to back} and written straight into mapped
// pinned host memory — the device-to-host traffic is the results, not the padding.
// ------------------------------------------------------------------------------------------------
// exclusive scan of the per-query result counts in three small steps (any batch size, all SMs):
//   1. every block of 1024 queries leaves its total in offsets[first query of the block]
//   2. one block scans those totals in place (they become the blocks' bases) and writes offsets[n] = number of pairs
//   3. the pack kernel scans inside its block on top of the base
constexpr uint32_t kPackBlock = 1024;

__device__ __forceinline__ uint32_t block_excl_scan_1024(uint32_t v, uint32_t *warp_sum /* 32 */, uint32_t &total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += y;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t ws = warp_sum[lane];
        uint32_t wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, wi, o);
            if (lane >= (uint32_t)o) wi += y;
        }
        warp_sum[lane] = wi - ws; // exclusive prefix of the warp totals
        if (lane == 31) warp_sum[32] = wi;
    }
    __syncthreads();
    total = warp_sum[32];
    const uint32_t r = warp_sum[warp] + incl - v;
    __syncthreads(); // warp_sum may be reused by the caller's next round
    return r;
}

__global__ void __launch_bounds__(kPackBlock) pack_block_sums_kernel(const uint32_t *counts, uint32_t n, uint32_t k_stride,
                                                                      uint32_t *offsets) {
    __shared__ uint32_t warp_sum[33];
    const uint32_t q = blockIdx.x * kPackBlock + threadIdx.x;
    uint32_t total;
    block_excl_scan_1024(q < n ? min(counts[q], k_stride) : 0u, warp_sum, total);
    if (threadIdx.x == 0) offsets[(size_t)blockIdx.x * kPackBlock] = total;
}

__global__ void __launch_bounds__(kPackBlock) pack_scan_sums_kernel(uint32_t *offsets, uint32_t n) {
    __shared__ uint32_t warp_sum[33];
    const uint32_t n_blocks = (n + kPackBlock - 1) / kPackBlock;
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += kPackBlock) {
        const uint32_t b = b0 + threadIdx.x;
        const uint32_t v = b < n_blocks ? offsets[(size_t)b * kPackBlock] : 0u;
        uint32_t total;
        const uint32_t ex = block_excl_scan_1024(v, warp_sum, total);
        if (b < n_blocks) offsets[(size_t)b * kPackBlock] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) offsets[n] = carry;
}

__global__ void __launch_bounds__(kPackBlock) result_pack_kernel(const uint32_t *ids, const uint32_t *scores, const uint32_t *counts,
                                                                  uint32_t *offsets, uint32_t n, uint32_t k_stride,
                                                                  uint32_t *out_counts, uint2 *out_pairs, uint32_t capacity) {
    __shared__ uint32_t warp_sum[33];
    const uint32_t q = blockIdx.x * kPackBlock + threadIdx.x;
    const uint32_t c = q < n ? min(counts[q], k_stride) : 0u;
    const uint32_t base = offsets[(size_t)blockIdx.x * kPackBlock]; // read by every thread before thread 0 rewrites it
    uint32_t total;
    const uint32_t o = base + block_excl_scan_1024(c, warp_sum, total);
    if (q >= n) return;
    offsets[q] = o;
    out_counts[q] = c;
    for (uint32_t j = 0; j < c && o + j < capacity; ++j)
        out_pairs[o + j] = make_uint2(ids[(size_t)q * k_stride + j], scores[(size_t)q * k_stride + j]);
}

// ------------------------------------------------------------------------------------------------
// docid-range shards: merge the shards' packed top-k lists (each already in (score desc, id asc) order, searched with
// the absolute floor only) into the global list, then the cutoffs of common.zig:153-166 anchored on the global best.
// One warp per query; lane g walks shard g's list.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) merge_packed_shards_kernel(const uint32_t *packed, unsigned long long stride_words,
                                                                   uint32_t n_shards, uint32_t n, const SearchOpts *opts,
                                                                   uint32_t k_stride, uint32_t *out_ids, uint32_t *out_scores,
                                                                   uint32_t *out_counts) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= n) return;
    const uint2 *pairs = nullptr;
    uint32_t cnt = 0, pos = 0;
    if (lane < n_shards) {
        const uint32_t *base = packed + (size_t)lane * stride_words;
        cnt = min(base[q], k_stride);
        pairs = reinterpret_cast<const uint2 *>(base + 2 * (size_t)n + 2) + base[(size_t)n + q];
    }
    unsigned long long key = ~0ull;
    if (cnt) {
        const uint2 p = pairs[0];
        key = rank_key(p.y, p.x);
    }
    const SearchOpts o = opts[q];
    const uint32_t k_eff = min(o.max_results, k_stride);
    uint32_t ms = o.min_score, n_out = 0;
    while (n_out < k_eff) {
        unsigned long long best = key;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, d));
        if (best == ~0ull) break;
        const uint32_t score = 0xFFFFFFFFu - (uint32_t)(best >> 32);
        if (score < ms) break;
        if (n_out == 0) ms = max(ms, (uint32_t)(score * o.min_score_pct) / 100u); // after the test: the best is always kept
        if (lane == 0) {
            out_ids[(size_t)q * k_stride + n_out] = (uint32_t)best;
            out_scores[(size_t)q * k_stride + n_out] = score;
        }
        ++n_out;
        if (key == best) { // docid ranges are disjoint: exactly one lane
            ++pos;
            key = ~0ull;
            if (pos < cnt) {
                const uint2 p = pairs[pos];
                key = rank_key(p.y, p.x);
            }
        }
    }
    if (lane == 0) out_counts[q] = n_out;
}

} // namespace

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
// Warp split of the hot kernel: counter / resolver-group / producer warps (FPX_DEBUG_ABLATE bits 24..27 pick another
// one for A/B runs).
// Measured on C3 (tools/sweep.py, hot kernel per 100 K queries, profiles/r02/): 8+8 counter / 2x4 resolver / 8 producer
// warps 1.120 ms; 9+9 / 2x4 / 6 1.161; 7+7 / 2x4 / 10 1.205; the round-1 kernel 1.215.
#define FPX_FIND_CONFIGS(X) X(0, 8, 2, 8, 4, 15, kStageU4) X(1, 9, 2, 6, 4, 15, kStageU4) X(2, 7, 2, 10, 4, 15, kStageU4) X(3, 6, 2, 12, 4, 15, kStageU4)

cudaError_t configure_kernels() {
    cudaError_t e;
    e = cudaFuncSetAttribute(search_smem_kernel<13, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_for<13, 256>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(search_smem_kernel<14, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_for<14, 256>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(search_smem_kernel<15, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_for<15, 1024>());
    if (e != cudaSuccess) return e;
#define X(I, GW, RG, PW, ST, SK, SU4)                                                                                 \
    e = cudaFuncSetAttribute(search_find_kernel<GW, RG, PW, ST, SU4, SK>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             (int)find_smem_bytes<ST, SU4, SK>());                                                    \
    if (e != cudaSuccess) return e;
    FPX_FIND_CONFIGS(X)
#undef X
    e = cudaFuncSetAttribute(search_find_kernel<8, 2, 8, 4, kStageU4, 15, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)find_smem_bytes<4, kStageU4, 15>());
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
    // one wave of resident CTAs (five per SM, see the launch bounds); measured on C3: 0.321 ms, two waves 0.328, four
    // 0.334, a warp per query 0.589
    if (blocks > 148 * 5) blocks = 148 * 5;
    prepare_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(a);
}

void launch_prepare_long(const BatchArgs &a, cudaStream_t st, int n_sms) {
    prepare_long_kernel<<<n_sms, kThreads, 0, st>>>(a);
}

void launch_search_sketch(const BatchArgs &a, cudaStream_t st, int n_sms) {
    if (a.debug & 512u) { // the default configuration with phase timers that see the barrier waits
        search_find_kernel<8, 2, 8, 4, kStageU4, 15, true><<<n_sms, 1024, find_smem_bytes<4, kStageU4, 15>(), st>>>(a, kSketchClass);
        return;
    }
    switch ((a.debug >> 24) & 15u) {
#define X(I, GW, RG, PW, ST, SK, SU4)                                                                                 \
    case I:                                                                                                           \
        search_find_kernel<GW, RG, PW, ST, SU4, SK>                                                                   \
            <<<n_sms, (2 * GW + 4 * RG + PW) * 32, find_smem_bytes<ST, SU4, SK>(), st>>>(a, kSketchClass);            \
        break;
        FPX_FIND_CONFIGS(X)
#undef X
    default: break;
    }
}

void launch_search_class(const BatchArgs &a, int cls, cudaStream_t st, int n_sms) {
    switch (cls) {
    case 1: search_smem_kernel<13, 256><<<n_sms * 4, 256, smem_bytes_for<13, 256>(), st>>>(a); break;
    case 2: search_smem_kernel<14, 256><<<n_sms * 3, 256, smem_bytes_for<14, 256>(), st>>>(a); break;
    case 3: search_smem_kernel<15, 1024><<<n_sms * 1, 1024, smem_bytes_for<15, 1024>(), st>>>(a); break;
    default: break;
    }
}

void launch_result_pack(const uint32_t *ids, const uint32_t *scores, const uint32_t *counts, uint32_t *offsets, uint32_t n,
                        uint32_t k_stride, uint32_t *out_counts, uint2 *out_pairs, cudaStream_t st, uint32_t capacity) {
    if (n == 0) return;
    const uint32_t nb = (n + kPackBlock - 1) / kPackBlock;
    pack_block_sums_kernel<<<nb, kPackBlock, 0, st>>>(counts, n, k_stride, offsets);
    pack_scan_sums_kernel<<<1, kPackBlock, 0, st>>>(offsets, n);
    result_pack_kernel<<<nb, kPackBlock, 0, st>>>(ids, scores, counts, offsets, n, k_stride, out_counts, out_pairs, capacity);
}

void launch_merge_packed_shards(const uint32_t *packed, uint64_t stride_words, uint32_t n_shards, uint32_t n,
                                const SearchOpts *opts, uint32_t k_stride, uint32_t *out_ids, uint32_t *out_scores,
                                uint32_t *out_counts, cudaStream_t st) {
    if (n == 0) return;
    const unsigned blocks = (unsigned)(((unsigned long long)n * 32 + 255) / 256);
    merge_packed_shards_kernel<<<blocks, 256, 0, st>>>(packed, stride_words, n_shards, n, opts, k_stride, out_ids, out_scores,
                                                       out_counts);
}

int wide_ctas(int n_sms) { return n_sms > 64 ? 64 : n_sms; }

void launch_search_wide(const BatchArgs &a, cudaStream_t st, int n_ctas) {
    search_wide_kernel<<<n_ctas, kThreads, 0, st>>>(a);
}

} // namespace fpx
