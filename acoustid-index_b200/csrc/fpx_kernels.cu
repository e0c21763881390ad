// fpx_kernels.cu — sm_100a kernels of the batched `_search` path.
//
// Per query the reference does (src/Index.zig:170-177, src/common.zig:121-167):
//   sort+dedup terms -> per term scan its postings -> hash-map count per docid -> filter, sort, top-k.
// Here, over a whole batch, against the CSR built by fpx_snapshot_host.h:
//   prepare_kernel       one warp per query: dedup terms, probe the term directory, emit row
//                        descriptors, bin the query by posting volume
//   search_sketch_kernel the hot path (min_score >= 2, i.e. every HTTP-default query of >= 21 terms):
//                        warp-specialised persistent CTAs.  Producer warps gather the query's posting rows
//                        into a shared-memory stage with TMA bulk copies (cp.async.bulk + mbarrier
//                        complete_tx); consumer warps (1) scatter-add every docid into a per-query u16
//                        count sketch with fire-and-forget shared atomics, (2) re-read the staged postings
//                        and send only docids whose sketch counter reaches min_score to a small exact
//                        table, (3) rank the survivors.  The sketch never under-counts, so this is exact.
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

// Bin a prepared query (called by one thread).
__device__ void classify_and_enqueue(const BatchArgs &a, uint32_t q, uint32_t rows_off, uint32_t n_rows,
                                     unsigned long long postings, unsigned long long total4, uint32_t n_unique) {
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
    WorkItem w;
    w.q = q;
    w.rows_off = rows_off;
    w.n_rows = n_rows;
    w.total4 = total4 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)total4;
    w.postings = postings > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)postings;
    w.k_eff = k_eff;
    w.min_score = o.min_score;
    w.min_score_pct = o.min_score_pct;
    uint32_t cls;
    // The sketch path needs min_score >= 2 (it only recounts docids whose sketch counter reaches
    // min_score) and the query must fit one 32 KB stage.  With a low floor and many postings nearly every
    // counter is "hot" and the recount list would overflow, so those go to the exact count-table path.
    bool sketch_ok = a.use_sketch && o.min_score >= 2 && total4 <= kStageU4 && k_eff <= kFastKbuf;
    if (o.min_score == 2 && postings > 2048) sketch_ok = false;
    if (o.min_score == 3 && postings > 5000) sketch_ok = false;
    if (sketch_ok)
        cls = kSketchClass;
    else
        cls = exact_class_for(postings, k_eff);
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
        unsigned long long postings = 0, total4 = 0;
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
                total4 += (ln[j] + 3) >> 2;
            }
            n_rows += __popc(fm);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            postings += __shfl_xor_sync(0xFFFFFFFFu, postings, off);
            total4 += __shfl_xor_sync(0xFFFFFFFFu, total4, off);
        }
        if (lane == 0) classify_and_enqueue(a, q, (uint32_t)o0, n_rows, postings, total4, n_unique);
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
                a.rows[o0 + atomicAdd(&s_rows, 1u)] = make_uint2(st, ln);
                atomicAdd(&s_post, (unsigned long long)ln);
                atomicAdd(&s_tot4, (unsigned long long)((ln + 3) >> 2));
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

struct Group {
    uint32_t tid, size, bar; // thread index in the group, group size (multiple of 32), named barrier id
    __device__ __forceinline__ void sync() const {
        asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(size) : "memory");
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
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// for waits that are expected to be long: do not burn issue slots the other warps need
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(128);
}
// global -> shared bulk copy (16-byte aligned, size multiple of 16), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------------
// sketch path (class 0)
// ------------------------------------------------------------------------------------------------
constexpr int kSkConsumerWarps = 8;
constexpr int kSkProducerWarps = 4; // all producers fill the same stage: row r is issued by warp r % 4
constexpr int kSkStages = 2;
constexpr int kSkThreads = (kSkConsumerWarps + kSkProducerWarps) * 32;
constexpr int kSkConsumers = kSkConsumerWarps * 32;
constexpr uint32_t kSketchLog = 14;                          // 16384 u16 counters = 32 KB
constexpr uint32_t kSketchWords = (1u << kSketchLog) / 2;
constexpr uint32_t kExSlots = 1024;                          // exact table: at most 512 distinct docids
constexpr uint32_t kHotCap = kFastKbuf * 2;                  // sketch-hot postings per query (aliases kbuf)
constexpr size_t kSkSmemBytes = (size_t)kSketchWords * 4 + (size_t)kSkStages * kStageU4 * 16 +
                                kExSlots * 8 + kFastKbuf * 8;

__device__ __forceinline__ void sketch_add(uint32_t *sketch, uint32_t d, uint32_t pad) {
    const uint32_t h = (d * kMult) >> (32 - kSketchLog);
    atomicAdd(sketch + (h >> 1), (d != pad ? 1u : 0u) << ((h & 1u) * 16u));
}
__device__ __forceinline__ uint32_t sketch_get(const uint32_t *sketch, uint32_t d) {
    const uint32_t h = (d * kMult) >> (32 - kSketchLog);
    return (sketch[h >> 1] >> ((h & 1u) * 16u)) & 0xFFFFu;
}

__global__ void __launch_bounds__(kSkThreads, 2) search_sketch_kernel(BatchArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *sketch = reinterpret_cast<uint32_t *>(smem_raw);
    uint4 *stage = reinterpret_cast<uint4 *>(smem_raw + (size_t)kSketchWords * 4);
    uint32_t *ex_keys = reinterpret_cast<uint32_t *>(smem_raw + (size_t)kSketchWords * 4 +
                                                     (size_t)kSkStages * kStageU4 * 16);
    uint32_t *ex_cnts = ex_keys + kExSlots;
    unsigned long long *kbuf = reinterpret_cast<unsigned long long *>(ex_cnts + kExSlots);
    __shared__ uint64_t full[kSkStages], empty[kSkStages];
    __shared__ WorkItem meta[kSkStages];
    __shared__ uint32_t s_ncand, s_nkeys, s_ovf, s_count, s_nhot;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t count = a.counters->qcount[kSketchClass];
    const WorkItem *items = a.items + (size_t)kSketchClass * a.n_queries;
    const uint32_t pad = a.snap.pad_id;

    if (tid == 0) {
        for (int s = 0; s < kSkStages; ++s) {
            mbar_init(&full[s], kSkProducerWarps);
            mbar_init(&empty[s], 1);
        }
        s_ncand = s_nkeys = s_ovf = s_nhot = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < kSkConsumers) {
        for (uint32_t i = tid; i < kSketchWords / 4; i += kSkConsumers)
            reinterpret_cast<uint4 *>(sketch)[i] = make_uint4(0, 0, 0, 0);
        for (uint32_t i = tid; i < kExSlots; i += kSkConsumers) {
            ex_keys[i] = pad;
            ex_cnts[i] = 0;
        }
    }
    __syncthreads();

    if (warp >= kSkConsumerWarps) {
        // ===== producers: TMA bulk copies (UBLKCP) of the query's posting rows into stage it % 2.
        // Issue is the scarce resource (~70 clk per copy per warp), so the four producer warps split
        // every query row-wise; each computes all offsets itself, no inter-warp traffic.
        const uint32_t p = warp - kSkConsumerWarps;
        const uint4 *docids4 = reinterpret_cast<const uint4 *>(a.snap.docids);
        for (uint32_t it = 0;; ++it) {
            const unsigned long long idx = blockIdx.x + (unsigned long long)it * gridDim.x;
            if (idx >= count) break;
            const uint32_t s = it % kSkStages;
            const WorkItem w = items[idx];
            if (it >= kSkStages) { // wait until the consumers released the previous tenant of this stage
                if (lane == 0) mbar_wait_relaxed(&empty[s], ((it / kSkStages) - 1) & 1);
                __syncwarp();
            }
            uint4 *dst = stage + (size_t)s * kStageU4;
            uint32_t base = 0, mine = 0;
            // first pass over the row descriptors: my byte count (expect_tx must precede my copies)
            for (uint32_t r0 = 0; r0 < w.n_rows; r0 += 32) {
                const uint32_t r = r0 + lane;
                const uint32_t n4 = r < w.n_rows ? (a.rows[w.rows_off + r].y + 3) >> 2 : 0u;
                mine += (r % kSkProducerWarps == p) ? n4 : 0u;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
            if (lane == 0) {
                if (p == 0) meta[s] = w;
                mbar_expect_tx(&full[s], mine * 16u);
            }
            __syncwarp();
            for (uint32_t r0 = 0; r0 < w.n_rows; r0 += 32) {
                const uint32_t r = r0 + lane;
                const uint2 d = r < w.n_rows ? a.rows[w.rows_off + r] : make_uint2(0u, 0u);
                const uint32_t n4 = (d.y + 3) >> 2;
                uint32_t x = n4;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
                    if (lane >= (uint32_t)o) x += y;
                }
                if (n4 && (r % kSkProducerWarps == p)) bulk_g2s(dst + base + x - n4, docids4 + d.x, n4 * 16u, &full[s]);
                base += __shfl_sync(0xFFFFFFFFu, x, 31);
            }
        }
        return;
    }

    // ===== consumers
    const Group g{tid, (uint32_t)kSkConsumers, 1u};
    for (uint32_t it = 0;; ++it) {
        const unsigned long long idx = blockIdx.x + (unsigned long long)it * gridDim.x;
        if (idx >= count) break;
        const uint32_t s = it % kSkStages;
        if (warp == 0) { // one poller; everybody else sleeps in the hardware barrier
            if (lane == 0) mbar_wait(&full[s], (it / kSkStages) & 1);
            __syncwarp();
        }
        g.sync();
        const WorkItem w = meta[s];
        const uint4 *st = stage + (size_t)s * kStageU4;
        const uint32_t thr = w.min_score; // >= 2 in this class

        // pass 1: count sketch, branch-free (padding adds 0).  counter(h) >= true count of every docid
        // hashing to h, so no qualifying doc can be missed in pass 2.
        for (uint32_t i = tid; i < w.total4; i += kSkConsumers) {
            const uint4 v = st[i];
            sketch_add(sketch, v.x, pad);
            sketch_add(sketch, v.y, pad);
            sketch_add(sketch, v.z, pad);
            sketch_add(sketch, v.w, pad);
        }
        g.sync();
        // pass 2: postings whose counter reaches min_score are compacted into a short list ...
        uint32_t *hot = reinterpret_cast<uint32_t *>(kbuf); // kFastKbuf*2 u32 entries, free until the harvest
        for (uint32_t i = tid; i < w.total4; i += kSkConsumers) {
            const uint4 v = st[i];
            const uint32_t c0 = sketch_get(sketch, v.x), c1 = sketch_get(sketch, v.y);
            const uint32_t c2 = sketch_get(sketch, v.z), c3 = sketch_get(sketch, v.w);
            const uint32_t m = (c0 >= thr ? 1u : 0u) | (c1 >= thr ? 2u : 0u) | (c2 >= thr ? 4u : 0u) | (c3 >= thr ? 8u : 0u);
            if (m == 0u) continue;
            uint32_t pos = atomicAdd(&s_nhot, (uint32_t)__popc(m));
            if ((m & 1u) && pos < kHotCap) hot[pos++] = v.x; else pos += (m & 1u);
            if ((m & 2u) && pos < kHotCap) hot[pos++] = v.y; else pos += (m >> 1) & 1u;
            if ((m & 4u) && pos < kHotCap) hot[pos++] = v.z; else pos += (m >> 2) & 1u;
            if ((m & 8u) && pos < kHotCap) hot[pos] = v.w;
        }
        g.sync();
        if (tid == 0) mbar_arrive(&empty[s]); // stage may be refilled
        const uint32_t nhot = s_nhot;
        // ... and counted exactly, all lanes busy (pads never get here: their slot only holds real counts,
        // and a pad posting itself is filtered below)
        if (nhot <= kHotCap) {
            for (uint32_t i = tid; i < nhot; i += kSkConsumers) {
                const uint32_t d = hot[i];
                if (d == pad) continue;
                uint32_t x = (d * kMult2) >> 22;
                for (uint32_t tries = 0;; ++tries) {
                    const uint32_t old = atomicCAS(ex_keys + x, pad, d);
                    if (old == pad) {
                        if (atomicAdd(&s_nkeys, 1u) >= kExSlots / 2) s_ovf = 1u;
                    }
                    if (old == pad || old == d) {
                        atomicAdd(ex_cnts + x, 1u);
                        break;
                    }
                    x = (x + 1) & (kExSlots - 1);
                    if (tries >= kExSlots || s_ovf) {
                        s_ovf = 1u;
                        break;
                    }
                }
            }
        } else if (tid == 0) {
            s_ovf = 1u;
        }
        g.sync();
        // clear the sketch, harvest + clear the exact table
        for (uint32_t i = tid; i < kSketchWords / 4; i += kSkConsumers)
            reinterpret_cast<uint4 *>(sketch)[i] = make_uint4(0, 0, 0, 0);
        if (s_nkeys != 0) {
            for (uint32_t i = tid; i < kExSlots; i += kSkConsumers) {
                const uint32_t k = ex_keys[i];
                if (k == pad) continue;
                const uint32_t c = ex_cnts[i];
                ex_keys[i] = pad;
                ex_cnts[i] = 0;
                if (c >= thr) {
                    const uint32_t pos = atomicAdd(&s_ncand, 1u);
                    if (pos < kFastKbuf) kbuf[pos] = rank_key(c, k);
                }
            }
        }
        g.sync();
        const uint32_t n = s_ncand;
        const bool redo = s_ovf || n > kFastKbuf;
        if (redo) {
            // more sketch-hot docids than the exact table holds: the exact count-table path takes the query
            if (tid == 0) {
                enqueue(a, exact_class_for(w.postings, w.k_eff), w);
                if (a.stats) atomicAdd(&a.stats->overflow_requeues, 1ull);
            }
        } else if (n == 0) {
            if (tid == 0) a.out_counts[w.q] = 0;
        } else {
            group_sort_keys(g, kbuf, n, kFastKbuf);
            group_emit_results(g, a, w, kbuf, n, &s_count);
        }
        if (tid == 0) {
            if (a.stats && !redo) atomicAdd(&a.stats->sketch_queries, 1ull);
        }
        g.sync();
        if (tid == 0) s_ncand = s_nkeys = s_ovf = s_nhot = 0;
        // the next query's first updates of these counters happen after two more group barriers
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
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem_raw);
    unsigned long long *kbuf = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)P::kSlots * 4);
    uint2 *rows_s = reinterpret_cast<uint2 *>(smem_raw + (size_t)P::kSlots * 4 + kFastKbuf * 8);
    __shared__ uint32_t s_idx, s_ncand, s_ovf, s_count;

    constexpr int cls = LOG - 12;
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const Group g{tid, (uint32_t)kThreads, 0u};
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
        const WorkItem w = a.items[(size_t)cls * a.n_queries + s_idx];
        const uint32_t thr = max(w.min_score, 1u); // a doc in the table has score >= 1
        const uint2 *rows = a.rows + w.rows_off;
        uint32_t passes = 1;
        if (LOG == 15)
            while ((unsigned long long)passes * 12288ull < w.postings && w.postings > 16384u) passes <<= 1;
        const uint32_t pmask = passes - 1u;

        for (uint32_t pass = 0; pass < passes; ++pass) {
            for (uint32_t r0 = 0; r0 < w.n_rows; r0 += kRowsChunk) {
                const uint32_t nr = min(kRowsChunk, w.n_rows - r0);
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
                const uint4 wd = tab4[i];
                if ((wd.x | wd.y | wd.z | wd.w) == 0u) continue;
                tab4[i] = make_uint4(0, 0, 0, 0);
                const uint32_t ws[4] = {wd.x, wd.y, wd.z, wd.w};
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
                enqueue(a, kWideClass, w);
                if (a.stats) atomicAdd(&a.stats->overflow_requeues, 1ull);
            }
            continue;
        }
        group_sort_keys(g, kbuf, n, kFastKbuf);
        group_emit_results(g, a, w, kbuf, n, &s_count);
    }
}

// ------------------------------------------------------------------------------------------------
// global-memory path: 64-bit slots (docid << 32 | count), linear probing, any size via hash partitions
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) search_wide_kernel(BatchArgs a) {
    __shared__ unsigned long long kbuf[kWideKbuf];
    __shared__ uint2 rows_s[kRowsChunk];
    __shared__ uint32_t s_idx, s_kn, s_new, s_fail, s_count;
    __shared__ unsigned long long s_kth;

    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const Group g{tid, (uint32_t)kThreads, 0u};
    unsigned long long *table = a.wide_tables + ((size_t)blockIdx.x << a.wide_cap_log2);
    const uint4 *docids4 = reinterpret_cast<const uint4 *>(a.snap.docids);
    const uint32_t pad = a.snap.pad_id;
    const uint32_t qcount = a.counters->qcount[kWideClass];

    for (;;) {
        __syncthreads();
        if (tid == 0) s_idx = atomicAdd(&a.counters->qhead[kWideClass], 1u);
        __syncthreads();
        if (s_idx >= qcount) break;
        const WorkItem w = a.items[(size_t)kWideClass * a.n_queries + s_idx];
        const uint32_t thr = max(w.min_score, 1u);
        const uint32_t k_eff = min(w.k_eff, kMaxResults);
        const uint2 *rows = a.rows + w.rows_off;
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
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(search_sketch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSkSmemBytes);
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

void launch_search_sketch(const BatchArgs &a, cudaStream_t st, int n_sms) {
    search_sketch_kernel<<<n_sms * 2, kSkThreads, kSkSmemBytes, st>>>(a);
}

void launch_search_class(const BatchArgs &a, int cls, cudaStream_t st, int n_sms) {
    switch (cls) {
    case 1: search_smem_kernel<13><<<n_sms * 4, kThreads, smem_bytes_for<13>(), st>>>(a); break;
    case 2: search_smem_kernel<14><<<n_sms * 3, kThreads, smem_bytes_for<14>(), st>>>(a); break;
    case 3: search_smem_kernel<15><<<n_sms * 1, kThreads, smem_bytes_for<15>(), st>>>(a); break;
    default: break;
    }
}

int wide_ctas(int n_sms) { return n_sms > 64 ? 64 : n_sms; }

void launch_search_wide(const BatchArgs &a, cudaStream_t st, int n_ctas) {
    search_wide_kernel<<<n_ctas, kThreads, 0, st>>>(a);
}

} // namespace fpx
