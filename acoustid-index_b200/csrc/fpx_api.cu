// fpx_api.cu — the extern "C" boundary declared in include/fpx.h.
//
// Host-side responsibilities: context / workspace pool, snapshot lifetime (immutable, atomically
// refcounted — mirrors SharedPtr(Segments), src/shared_ptr.zig + src/Index.zig:430-485), upload of the
// compiled CSR, and the batched search driver (chunked H2D -> kernels -> D2H on rotating streams).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/fpx.h"
#include "fpx_gpu_build.h"
#include "fpx_kernels.cuh"
#include "fpx_snapshot_host.h"

using namespace fpx;

static_assert(sizeof(fpx_search_opts) == sizeof(SearchOpts), "opts layout");
static_assert(sizeof(TermEntry) == 16, "directory entry is one 128-bit load");
static_assert(FPX_MAX_QUERY_TERMS == kMaxQueryTerms && FPX_MAX_RESULTS == kMaxResults, "limits");

namespace {

thread_local std::string g_error;

fpx_status set_error(fpx_status st, const std::string &msg) {
    g_error = msg;
    return st;
}
fpx_status cuda_fail(cudaError_t e, const char *what) {
    g_error = std::string(what) + ": " + cudaGetErrorString(e);
    if (e == cudaErrorMemoryAllocation) return FPX_OUT_OF_MEMORY;
    return FPX_CUDA_ERROR;
}
#define FPX_CUDA(call)                                           \
    do {                                                         \
        cudaError_t e__ = (call);                                \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);    \
    } while (0)

enum KernelKind { KK_PREPARE = 0, KK_SEARCH = 1, KK_WIDE = 2, KK_H2D = 3, KK_D2H = 4, KK_SKETCH = 5 };

struct EventPair {
    int kind;
    cudaEvent_t a, b;
};

template <class T> struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

constexpr int kSlots = 3;

struct Workspace {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr; // last use (async device API)
    bool done_pending = false;
    // Everything one batch (or one chunk of a host batch) writes on the device.  Host batches rotate through
    // kSlots slots so that chunk c's small trailing kernels can run beside the big ones of the chunks after it.
    struct Slot {
        DevBuf<uint4> rows;
        DevBuf<WorkItem> items;
        DevBuf<uint32_t> long_queue;
        BatchCounters *counters = nullptr;
        DevBuf<uint32_t> d_ids, d_scores, d_counts, d_pack_offsets; // host batches: results before packing
        cudaEvent_t front_done = nullptr;
    } slot[kSlots];
    cudaStream_t tail_stream = nullptr;
    cudaStream_t d2h_stream = nullptr; // result DMA of chunk c must not hold up the trailing kernels of chunk c+1
    cudaEvent_t tail_done = nullptr;
    // staging for host batches
    DevBuf<uint32_t> d_terms;
    DevBuf<uint64_t> d_offsets;
    DevBuf<SearchOpts> d_opts;
    uint32_t *h_error = nullptr; // pinned
    // packed results of one chunk, written by the GPU into mapped pinned host memory
    uint32_t *h_counts = nullptr;
    uint2 *h_pairs = nullptr;
    size_t h_counts_cap = 0, h_pairs_cap = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_fence = nullptr; // the last host batch's kernels are done with the staging buffers
    std::vector<cudaEvent_t> h2d_done, chunk_done;

    cudaError_t reserve_events(size_t n) {
        while (h2d_done.size() < n) {
            cudaEvent_t a = nullptr, b = nullptr;
            cudaError_t e = cudaEventCreateWithFlags(&a, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b, cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
            h2d_done.push_back(a);
            chunk_done.push_back(b);
        }
        return cudaSuccess;
    }

    cudaError_t reserve_host(size_t nq, size_t pairs) {
        if (nq > h_counts_cap) {
            if (h_counts) cudaFreeHost(h_counts);
            h_counts = nullptr;
            h_counts_cap = 0;
            cudaError_t e = cudaHostAlloc(&h_counts, (nq + nq / 4 + 64) * sizeof(uint32_t), cudaHostAllocMapped);
            if (e != cudaSuccess) return e;
            h_counts_cap = nq + nq / 4 + 64;
        }
        if (pairs > h_pairs_cap) {
            if (h_pairs) cudaFreeHost(h_pairs);
            h_pairs = nullptr;
            h_pairs_cap = 0;
            cudaError_t e = cudaHostAlloc(&h_pairs, (pairs + pairs / 4 + 64) * sizeof(uint2), cudaHostAllocMapped);
            if (e != cudaSuccess) return e;
            h_pairs_cap = pairs + pairs / 4 + 64;
        }
        return cudaSuccess;
    }

    ~Workspace() {
        for (Slot &sl : slot) {
            sl.rows.release();
            sl.items.release();
            sl.long_queue.release();
            sl.d_ids.release();
            sl.d_scores.release();
            sl.d_counts.release();
            sl.d_pack_offsets.release();
            if (sl.counters) cudaFree(sl.counters);
            if (sl.front_done) cudaEventDestroy(sl.front_done);
        }
        d_terms.release();
        d_offsets.release();
        d_opts.release();
        if (tail_stream) cudaStreamDestroy(tail_stream);
        if (d2h_stream) cudaStreamDestroy(d2h_stream);
        if (tail_done) cudaEventDestroy(tail_done);
        if (h_error) cudaFreeHost(h_error);
        if (h_counts) cudaFreeHost(h_counts);
        if (h_pairs) cudaFreeHost(h_pairs);
        for (cudaEvent_t ev : h2d_done) cudaEventDestroy(ev);
        for (cudaEvent_t ev : chunk_done) cudaEventDestroy(ev);
        if (copy_fence) cudaEventDestroy(copy_fence);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (done) cudaEventDestroy(done);
        if (stream) cudaStreamDestroy(stream);
    }
};

constexpr uint32_t kWideCapLog2 = 20;
constexpr uint32_t kErrSlots = 4096; // per-chunk device error words of a host batch

} // namespace

namespace fpx {
void set_last_error(const std::string &msg) { g_error = msg; } // for the other translation units
} // namespace fpx

struct fpx_ctx {
    int device = 0;
    int n_sms = 148;
    unsigned host_threads = 1;
    uint32_t chunk_queries = 131072;
    uint32_t flags = 0;
    bool host_only = false;
    bool use_sketch = true;
    uint32_t debug = 0; // FPX_DEBUG_ABLATE environment variable, profiling only
    std::mutex mu;
    std::vector<Workspace *> free_ws;
    // one pool of global-memory count tables for the rare wide path, shared by all workspaces: concurrent batches'
    // wide kernels are chained on the device by `wide_done`
    unsigned long long *d_wide = nullptr;
    cudaEvent_t wide_done = nullptr;
    std::mutex wide_mu;
    DeviceStats *d_stats = nullptr;
    std::vector<EventPair> pending;
    fpx_profile prof{};
    std::atomic<uint32_t> sticky_error{0};
};

struct fpx_snapshot_builder {
    fpx_ctx *ctx;
    SnapshotCompiler compiler;                 // host build (and the debug CSR view)
    std::unique_ptr<GpuSnapshotBuilder> gpu;   // device build: the default on a device context
    explicit fpx_snapshot_builder(fpx_ctx *c) : ctx(c), compiler(c->host_threads) {}
};

struct fpx_snapshot {
    fpx_ctx *ctx = nullptr;
    std::atomic<int64_t> refs{1};
    SnapshotDev dev{};
    TermEntry *d_table = nullptr;
    uint32_t *d_docids = nullptr;
    fpx_snapshot_info info{};
    std::vector<uint32_t> h_terms, h_row_len, h_row_start4; // host copy of the term directory (ascending terms)
};

namespace {

fpx_status acquire_workspace(fpx_ctx *ctx, Workspace **out) {
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (!ctx->free_ws.empty()) {
            *out = ctx->free_ws.back();
            ctx->free_ws.pop_back();
            return FPX_OK;
        }
    }
    Workspace *w = new (std::nothrow) Workspace();
    if (!w) return set_error(FPX_OUT_OF_MEMORY, "workspace");
    cudaError_t e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->copy_fence, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w->tail_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w->d2h_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->tail_done, cudaEventDisableTiming);
    for (Workspace::Slot &sl : w->slot) {
        if (e == cudaSuccess) e = cudaMalloc(&sl.counters, sizeof(BatchCounters));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.front_done, cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaMallocHost(&w->h_error, kErrSlots * sizeof(uint32_t));
    if (e != cudaSuccess) {
        delete w;
        return cuda_fail(e, "workspace allocation");
    }
    std::memset(w->h_error, 0, kErrSlots * sizeof(uint32_t));
    *out = w;
    return FPX_OK;
}

constexpr size_t kMaxIdleWorkspaces = 8; // a burst of concurrent callers does not pin its staging memory for good

void release_workspace(fpx_ctx *ctx, Workspace *w) {
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (ctx->free_ws.size() < kMaxIdleWorkspaces) {
            ctx->free_ws.push_back(w);
            return;
        }
    }
    if (w->done_pending) cudaEventSynchronize(w->done); // its last batch may still be running
    delete w;
}

struct Timed {
    fpx_ctx *ctx;
    cudaStream_t st;
    int kind;
    cudaEvent_t a = nullptr, b = nullptr;
    Timed(fpx_ctx *c, cudaStream_t s, int k) : ctx(c), st(s), kind(k) {
        if (ctx->flags & FPX_FLAG_PROFILE) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, st);
        }
    }
    ~Timed() {
        if (a) {
            cudaEventRecord(b, st);
            std::lock_guard<std::mutex> lk(ctx->mu);
            ctx->pending.push_back(EventPair{kind, a, b});
        }
    }
};

// Enqueue the kernels of one batch.  All pointers are device pointers.  The "front" (prepare + the persistent
// sketch kernel) goes on `st`, the "tail" (exact count-table kernels, which also take the sketch kernel's
// re-queued queries, and the error word) on `ts`; ts == st keeps everything in one stream.
fpx_status enqueue_batch(fpx_snapshot *s, Workspace *w, Workspace::Slot &sl, cudaStream_t st, cudaStream_t ts,
                         uint64_t n_queries, uint64_t n_terms_total, const uint32_t *d_terms, const uint64_t *d_offsets,
                         uint64_t term_base, const SearchOpts *d_opts, uint32_t k_stride, uint32_t *d_ids,
                         uint32_t *d_scores, uint32_t *d_counts,
                         cudaEvent_t *trace = nullptr /* 3 events: after prepare, sketch, rest */, uint32_t err_slot = 0,
                         bool no_long_queries = false) {
    fpx_ctx *ctx = s->ctx;
    if (n_queries == 0) return FPX_OK;
    if (n_queries > 0x7FFFFFFFull || n_terms_total > 0xFFFFFFFFull)
        return set_error(FPX_INVALID_ARGUMENT, "batch too large (split it)");
    cudaError_t e = sl.rows.reserve(n_terms_total + 1);
    if (e == cudaSuccess) e = sl.items.reserve(n_queries * kNumClasses);
    if (e == cudaSuccess) e = sl.long_queue.reserve(n_queries);
    if (e != cudaSuccess) return cuda_fail(e, "workspace growth");

    BatchArgs a{};
    a.snap = s->dev;
    a.n_queries = (uint32_t)n_queries;
    a.k_stride = k_stride;
    a.terms = d_terms;
    a.term_offsets = d_offsets;
    a.term_base = term_base;
    a.opts = d_opts;
    a.out_ids = d_ids;
    a.out_scores = d_scores;
    a.out_counts = d_counts;
    a.rows = sl.rows.p;
    a.items = sl.items.p;
    a.long_queue = sl.long_queue.p;
    a.counters = sl.counters;
    a.stats = (ctx->flags & FPX_FLAG_PROFILE) ? ctx->d_stats : nullptr;
    a.wide_tables = ctx->d_wide;
    a.n_terms_total = n_terms_total;
    a.wide_cap_log2 = kWideCapLog2;
    a.use_sketch = ctx->use_sketch ? 1u : 0u;
    a.debug = ctx->debug;

    FPX_CUDA(cudaMemsetAsync(sl.counters, 0, sizeof(BatchCounters), st));
    {
        Timed t(ctx, st, KK_PREPARE);
        launch_prepare(a, st);
        if (!no_long_queries) launch_prepare_long(a, st, ctx->n_sms);
    }
    if (trace) cudaEventRecord(trace[0], st);
    if (a.use_sketch) {
        Timed t(ctx, st, KK_SKETCH);
        launch_search_sketch(a, st, ctx->n_sms);
    }
    if (trace) cudaEventRecord(trace[1], st);
    if (ts != st) {
        FPX_CUDA(cudaEventRecord(sl.front_done, st));
        FPX_CUDA(cudaStreamWaitEvent(ts, sl.front_done, 0));
    }
    {
        Timed t(ctx, ts, KK_SEARCH);
        for (int c = 1; c <= 3; ++c) launch_search_class(a, c, ts, ctx->n_sms);
    }
    {
        std::lock_guard<std::mutex> lk(ctx->wide_mu); // wait, launch and record as one step of the chain
        FPX_CUDA(cudaStreamWaitEvent(ts, ctx->wide_done, 0));
        {
            Timed t(ctx, ts, KK_WIDE);
            launch_search_wide(a, ts, wide_ctas(ctx->n_sms));
        }
        FPX_CUDA(cudaEventRecord(ctx->wide_done, ts));
    }
    if (trace) cudaEventRecord(trace[2], ts);
    FPX_CUDA(cudaGetLastError());
    FPX_CUDA(cudaMemcpyAsync(w->h_error + err_slot, &sl.counters->error, sizeof(uint32_t), cudaMemcpyDeviceToHost, ts));
    return FPX_OK;
}

// fpx_snapshot_commit for a device-built snapshot (fpx_gpu_build.cu)
fpx_status commit_gpu_built(fpx_snapshot_builder *b, fpx_snapshot **out) {
    fpx_ctx *ctx = b->ctx;
    FPX_CUDA(cudaSetDevice(ctx->device));
    GpuCsr csr;
    if (!b->gpu->build(csr)) {
        if (csr.d_docids) cudaFree(csr.d_docids);
        if (csr.d_terms) cudaFree(csr.d_terms);
        if (csr.d_row_len) cudaFree(csr.d_row_len);
        if (csr.d_row_start4) cudaFree(csr.d_row_start4);
        if (b->gpu->oom) return set_error(FPX_OUT_OF_MEMORY, b->gpu->error);
        if (b->gpu->unsupported) return set_error(FPX_UNSUPPORTED, b->gpu->error + " (use FPX_FLAG_HOST_BUILD)");
        return set_error(FPX_INVALID_SEGMENT, b->gpu->error);
    }
    fpx_snapshot *s = new (std::nothrow) fpx_snapshot();
    const uint64_t nt = csr.n_terms;
    uint32_t log2cap = 4;
    while ((1ull << log2cap) < 2 * nt) ++log2cap;
    cudaError_t e = cudaSuccess;
    if (!s || log2cap > 31) e = cudaErrorMemoryAllocation;
    const size_t cap = (size_t)1 << log2cap;
    if (e == cudaSuccess) e = cudaMalloc(&s->d_table, cap * sizeof(TermEntry));
    if (e == cudaSuccess) e = cudaMemset(s->d_table, 0, cap * sizeof(TermEntry));
    if (e == cudaSuccess && nt) {
        launch_build_table(s->d_table, log2cap, csr.d_terms, csr.d_row_len, csr.d_row_start4, nt, nullptr);
        e = cudaDeviceSynchronize();
    }
    if (e == cudaSuccess && nt)
        e = reorder_rows_by_key(&csr.d_docids, csr.total4 * 4, csr.d_row_len, csr.d_row_start4, csr.h_row_start4.data(), nt);
    if (csr.d_terms) cudaFree(csr.d_terms);
    if (csr.d_row_len) cudaFree(csr.d_row_len);
    if (csr.d_row_start4) cudaFree(csr.d_row_start4);
    if (e != cudaSuccess) {
        if (s && s->d_table) cudaFree(s->d_table);
        cudaFree(csr.d_docids);
        delete s;
        return cuda_fail(e, "snapshot table build");
    }
    s->ctx = ctx;
    s->d_docids = csr.d_docids;
    s->dev.table = s->d_table;
    s->dev.table_mask = (uint32_t)(cap - 1);
    s->dev.table_shift = 32 - log2cap;
    s->dev.docids = s->d_docids;
    s->dev.pad_id = csr.pad_id;
    s->dev.pad_spread = csr.pad_spread ? 1u : 0u;
    s->info.n_segments = b->gpu->n_segments();
    s->info.n_terms = nt;
    s->info.n_postings = csr.n_postings;
    s->info.n_postings_total = csr.n_postings_total;
    s->info.n_dropped_unreachable = csr.n_unreachable;
    s->info.n_dropped_superseded = csr.n_superseded;
    s->info.n_dropped_out_of_range = csr.n_out_of_range;
    s->info.device_bytes = cap * sizeof(TermEntry) + std::max<uint64_t>(csr.total4 * 4, 4) * sizeof(uint32_t);
    s->info.max_row_len = csr.max_row_len;
    s->info.pad_id = csr.pad_id;
    s->info.table_log2 = log2cap;
    s->info.doc_lo = b->gpu->doc_lo();
    s->info.doc_hi = b->gpu->doc_hi();
    s->h_terms = std::move(csr.h_terms);
    s->h_row_len = std::move(csr.h_row_len);
    s->h_row_start4 = std::move(csr.h_row_start4);
    delete b;
    *out = s;
    return FPX_OK;
}

} // namespace

extern "C" {

uint32_t fpx_abi_version(void) { return FPX_ABI_VERSION; }

const char *fpx_last_error_message(void) { return g_error.c_str(); }

fpx_status fpx_init(const fpx_config *config, fpx_ctx **out) {
    if (!out) return set_error(FPX_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    fpx_config cfg{};
    cfg.device = -1;
    if (config) cfg = *config;
    fpx_ctx *ctx = new (std::nothrow) fpx_ctx();
    if (!ctx) return set_error(FPX_OUT_OF_MEMORY, "ctx");
    ctx->flags = cfg.flags;
    ctx->host_threads = cfg.host_threads ? cfg.host_threads : std::max(1u, std::thread::hardware_concurrency());
    if (cfg.chunk_queries) ctx->chunk_queries = cfg.chunk_queries;
    ctx->host_only = (cfg.flags & FPX_FLAG_HOST_ONLY) != 0;
    ctx->use_sketch = (cfg.flags & FPX_FLAG_NO_SKETCH) == 0;
    if (const char *dbg = std::getenv("FPX_DEBUG_ABLATE")) ctx->debug = (uint32_t)std::strtoul(dbg, nullptr, 0);
    if (!ctx->host_only) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) {
            delete ctx;
            cudaGetLastError();
            return set_error(FPX_BACKEND_UNAVAILABLE, "no CUDA device available; the fpx search path requires a GPU");
        }
        if (cfg.device >= 0) {
            if (cfg.device >= n) {
                delete ctx;
                return set_error(FPX_INVALID_ARGUMENT, "device ordinal out of range");
            }
            e = cudaSetDevice(cfg.device);
            ctx->device = cfg.device;
        } else {
            e = cudaGetDevice(&ctx->device);
        }
        cudaDeviceProp prop{};
        if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, ctx->device);
        if (e == cudaSuccess && prop.major < 10) {
            delete ctx;
            return set_error(FPX_BACKEND_UNAVAILABLE, "fpx kernels are built for sm_100a (Blackwell) only");
        }
        if (e == cudaSuccess) {
            ctx->n_sms = prop.multiProcessorCount;
            e = configure_kernels();
        }
        if (e == cudaSuccess) e = cudaMalloc(&ctx->d_stats, sizeof(DeviceStats));
        if (e == cudaSuccess) e = cudaMemset(ctx->d_stats, 0, sizeof(DeviceStats));
        if (e == cudaSuccess) e = cudaMalloc(&ctx->d_wide, ((size_t)wide_ctas(ctx->n_sms) << kWideCapLog2) * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->wide_done, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            fpx_status st = cuda_fail(e, "fpx_init");
            if (ctx->d_stats) cudaFree(ctx->d_stats);
            if (ctx->d_wide) cudaFree(ctx->d_wide);
            if (ctx->wide_done) cudaEventDestroy(ctx->wide_done);
            delete ctx;
            return st;
        }
    }
    *out = ctx;
    return FPX_OK;
}

void fpx_shutdown(fpx_ctx *ctx) {
    if (!ctx) return;
    if (!ctx->host_only) {
        cudaSetDevice(ctx->device);
        cudaDeviceSynchronize();
        for (Workspace *w : ctx->free_ws) delete w;
        for (auto &p : ctx->pending) {
            cudaEventDestroy(p.a);
            cudaEventDestroy(p.b);
        }
        if (ctx->d_stats) cudaFree(ctx->d_stats);
        if (ctx->d_wide) cudaFree(ctx->d_wide);
        if (ctx->wide_done) cudaEventDestroy(ctx->wide_done);
    }
    delete ctx;
}

fpx_status fpx_snapshot_begin(fpx_ctx *ctx, fpx_snapshot_builder **out) {
    if (!ctx || !out) return set_error(FPX_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    std::unique_ptr<fpx_snapshot_builder> b(new (std::nothrow) fpx_snapshot_builder(ctx));
    if (!b) return set_error(FPX_OUT_OF_MEMORY, "builder");
    if (!ctx->host_only && !(ctx->flags & FPX_FLAG_HOST_BUILD)) {
        FPX_CUDA(cudaSetDevice(ctx->device)); // b is freed on this early return
        b->gpu.reset(new (std::nothrow) GpuSnapshotBuilder());
        if (!b->gpu) return set_error(FPX_OUT_OF_MEMORY, "device builder");
    }
    *out = b.release();
    return FPX_OK;
}

fpx_status fpx_snapshot_add_file_segment(fpx_snapshot_builder *b, const fpx_file_segment *seg) {
    if (!b || !seg) return set_error(FPX_INVALID_ARGUMENT, "null argument");
    if (seg->n_docs && !seg->doc_ids) return set_error(FPX_INVALID_ARGUMENT, "null doc_ids");
    if (b->gpu) {
        cudaSetDevice(b->ctx->device);
        if (!b->gpu->add_file_segment(seg->commit_id, seg->merges, seg->min_doc_id, seg->block_size, seg->blocks, seg->num_blocks,
                                      seg->block_index, seg->doc_ids, seg->n_docs))
            return set_error(b->gpu->oom ? FPX_OUT_OF_MEMORY : b->gpu->unsupported ? FPX_UNSUPPORTED : FPX_INVALID_SEGMENT,
                             b->gpu->error + (b->gpu->unsupported ? " (use FPX_FLAG_HOST_BUILD)" : ""));
        return FPX_OK;
    }
    try {
        if (!b->compiler.add_file_segment(seg->commit_id, seg->merges, seg->min_doc_id, seg->block_size, seg->blocks,
                                          seg->num_blocks, seg->block_index, seg->doc_ids, seg->n_docs))
            return set_error(FPX_INVALID_SEGMENT, b->compiler.error);
    } catch (const std::bad_alloc &) {
        return set_error(FPX_OUT_OF_MEMORY, "decoding file segment");
    }
    return FPX_OK;
}

fpx_status fpx_snapshot_add_memory_segment(fpx_snapshot_builder *b, const fpx_memory_segment *seg) {
    if (!b || !seg) return set_error(FPX_INVALID_ARGUMENT, "null argument");
    if (seg->n_docs && !seg->doc_ids) return set_error(FPX_INVALID_ARGUMENT, "null doc_ids");
    if (b->gpu) {
        cudaSetDevice(b->ctx->device);
        if (!b->gpu->add_memory_segment(seg->commit_id, seg->merges, seg->items, seg->n_items, seg->doc_ids, seg->n_docs))
            return set_error(b->gpu->oom ? FPX_OUT_OF_MEMORY : FPX_INVALID_SEGMENT, b->gpu->error);
        return FPX_OK;
    }
    try {
        if (!b->compiler.add_memory_segment(seg->commit_id, seg->merges, seg->items, seg->n_items, seg->doc_ids,
                                            seg->n_docs))
            return set_error(FPX_INVALID_SEGMENT, b->compiler.error);
    } catch (const std::bad_alloc &) {
        return set_error(FPX_OUT_OF_MEMORY, "reading memory segment");
    }
    return FPX_OK;
}

fpx_status fpx_snapshot_set_doc_range(fpx_snapshot_builder *b, uint32_t lo, uint32_t hi) {
    if (!b) return set_error(FPX_INVALID_ARGUMENT, "null argument");
    if (hi != 0 && hi < lo) return set_error(FPX_INVALID_ARGUMENT, "hi < lo");
    b->compiler.set_doc_range(lo, hi);
    if (b->gpu) b->gpu->set_doc_range(lo, hi);
    return FPX_OK;
}

fpx_status fpx_snapshot_compile(fpx_snapshot_builder *b) {
    if (!b) return set_error(FPX_INVALID_ARGUMENT, "null argument");
    if (b->gpu)
        return set_error(FPX_UNSUPPORTED, "the host-side CSR view needs a host-built snapshot (FPX_FLAG_HOST_BUILD or FPX_FLAG_HOST_ONLY)");
    try {
        if (!b->compiler.compile()) return set_error(FPX_INVALID_SEGMENT, b->compiler.error);
    } catch (const std::bad_alloc &) {
        return set_error(FPX_OUT_OF_MEMORY, "compiling snapshot");
    }
    return FPX_OK;
}

fpx_status fpx_snapshot_csr(fpx_snapshot_builder *b, fpx_csr_view *out) {
    if (!b || !out) return set_error(FPX_INVALID_ARGUMENT, "null argument");
    fpx_status st = fpx_snapshot_compile(b);
    if (st != FPX_OK) return st;
    CompiledCsr *c = b->compiler.compiled();
    b->compiler.make_dense(*c);
    out->n_terms = c->terms.size();
    out->terms = c->terms.data();
    out->row_offsets = c->dense_offsets.data();
    out->docids = c->dense_docids.data();
    return FPX_OK;
}

void fpx_snapshot_abort(fpx_snapshot_builder *b) { delete b; }

fpx_status fpx_snapshot_commit(fpx_snapshot_builder *b, fpx_snapshot **out) {
    if (!b || !out) return set_error(FPX_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    fpx_ctx *ctx = b->ctx;
    if (ctx->host_only) return set_error(FPX_BACKEND_UNAVAILABLE, "context was created with FPX_FLAG_HOST_ONLY");
    if (b->gpu) return commit_gpu_built(b, out);
    fpx_status st = fpx_snapshot_compile(b);
    if (st != FPX_OK) return st;
    CompiledCsr *c = b->compiler.compiled();
    FPX_CUDA(cudaSetDevice(ctx->device));

    fpx_snapshot *s = new (std::nothrow) fpx_snapshot();
    if (!s) return set_error(FPX_OUT_OF_MEMORY, "snapshot");
    s->ctx = ctx;
    const size_t nt = c->terms.size();
    uint32_t log2cap = 4;
    while ((1ull << log2cap) < 2 * (uint64_t)nt) ++log2cap;
    if (log2cap > 31) {
        delete s;
        return set_error(FPX_UNSUPPORTED, "too many distinct terms");
    }
    const size_t cap = (size_t)1 << log2cap;
    const size_t n_doc_words = std::max<size_t>(c->docids.size(), 4);
    uint32_t *d_t = nullptr, *d_l = nullptr, *d_s4 = nullptr;
    cudaError_t e = cudaMalloc(&s->d_table, cap * sizeof(TermEntry));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_docids, n_doc_words * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(s->d_table, 0, cap * sizeof(TermEntry));
    if (e == cudaSuccess && !c->docids.empty())
        e = cudaMemcpy(s->d_docids, c->docids.data(), c->docids.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && nt) {
        e = cudaMalloc(&d_t, nt * 4);
        if (e == cudaSuccess) e = cudaMalloc(&d_l, nt * 4);
        if (e == cudaSuccess) e = cudaMalloc(&d_s4, nt * 4);
        if (e == cudaSuccess) e = cudaMemcpy(d_t, c->terms.data(), nt * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_l, c->row_len.data(), nt * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_s4, c->row_start4.data(), nt * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            launch_build_table(s->d_table, log2cap, d_t, d_l, d_s4, nt, nullptr);
            e = cudaDeviceSynchronize();
        }
        if (e == cudaSuccess)
            e = reorder_rows_by_key(&s->d_docids, c->docids.size(), d_l, d_s4, c->row_start4.data(), nt);
    }
    if (d_t) cudaFree(d_t);
    if (d_l) cudaFree(d_l);
    if (d_s4) cudaFree(d_s4);
    if (e != cudaSuccess) {
        if (s->d_table) cudaFree(s->d_table);
        if (s->d_docids) cudaFree(s->d_docids);
        delete s;
        return cuda_fail(e, "snapshot upload");
    }
    s->dev.table = s->d_table;
    s->dev.table_mask = (uint32_t)(cap - 1);
    s->dev.table_shift = 32 - log2cap;
    s->dev.docids = s->d_docids;
    s->dev.pad_id = c->pad_id;
    s->dev.pad_spread = c->pad_spread ? 1u : 0u;
    s->info.n_segments = b->compiler.n_segments();
    s->info.n_terms = nt;
    s->info.n_postings = c->n_postings;
    s->info.n_postings_total = c->n_postings_total;
    s->info.n_dropped_unreachable = c->n_unreachable;
    s->info.n_dropped_superseded = c->n_superseded;
    s->info.n_dropped_out_of_range = c->n_out_of_range;
    s->info.device_bytes = cap * sizeof(TermEntry) + n_doc_words * sizeof(uint32_t);
    s->info.max_row_len = c->max_row_len;
    s->info.pad_id = c->pad_id;
    s->info.table_log2 = log2cap;
    s->info.doc_lo = b->compiler.doc_lo();
    s->info.doc_hi = b->compiler.doc_hi();
    s->h_terms = std::move(c->terms);
    s->h_row_len = std::move(c->row_len);
    s->h_row_start4 = std::move(c->row_start4);
    delete b;
    *out = s;
    return FPX_OK;
}

fpx_status fpx_snapshot_acquire(fpx_snapshot *s) {
    if (!s) return set_error(FPX_INVALID_ARGUMENT, "null snapshot");
    s->refs.fetch_add(1, std::memory_order_relaxed);
    return FPX_OK;
}

fpx_status fpx_snapshot_release(fpx_snapshot *s) {
    if (!s) return set_error(FPX_INVALID_ARGUMENT, "null snapshot");
    if (s->refs.fetch_sub(1, std::memory_order_acq_rel) == 1) { // shared_ptr.zig:23-34
        cudaSetDevice(s->ctx->device);
        cudaDeviceSynchronize(); // no search may still be reading the rows
        if (s->d_table) cudaFree(s->d_table);
        if (s->d_docids) cudaFree(s->d_docids);
        delete s;
    }
    return FPX_OK;
}

fpx_status fpx_snapshot_get_info(const fpx_snapshot *s, fpx_snapshot_info *out) {
    if (!s || !out) return set_error(FPX_INVALID_ARGUMENT, "null argument");
    *out = s->info;
    return FPX_OK;
}

fpx_status fpx_snapshot_row_lengths(const fpx_snapshot *s, const uint32_t *terms, uint64_t n, uint32_t *out_lengths) {
    if (!s || (n && (!terms || !out_lengths))) return set_error(FPX_INVALID_ARGUMENT, "null argument");
    const auto &t = s->h_terms;
    parallel_for(n, s->ctx->host_threads, [&](size_t a, size_t b, unsigned) {
        for (size_t i = a; i < b; ++i) {
            auto it = std::lower_bound(t.begin(), t.end(), terms[i]);
            out_lengths[i] = (it != t.end() && *it == terms[i]) ? s->h_row_len[(size_t)(it - t.begin())] : 0u;
        }
    });
    return FPX_OK;
}

fpx_status fpx_snapshot_read_row(const fpx_snapshot *s, uint32_t term, uint32_t *out_docids, uint64_t capacity, uint64_t *out_len) {
    if (!s || !out_len) return set_error(FPX_INVALID_ARGUMENT, "null argument");
    *out_len = 0;
    auto it = std::lower_bound(s->h_terms.begin(), s->h_terms.end(), term);
    if (it == s->h_terms.end() || *it != term) return FPX_OK;
    const size_t g = (size_t)(it - s->h_terms.begin());
    const uint64_t len = s->h_row_len[g];
    *out_len = len;
    if (len > capacity || (len && !out_docids)) return set_error(FPX_INVALID_ARGUMENT, "row does not fit the buffer");
    FPX_CUDA(cudaSetDevice(s->ctx->device));
    FPX_CUDA(cudaMemcpy(out_docids, s->d_docids + (size_t)s->h_row_start4[g] * 4, len * 4, cudaMemcpyDeviceToHost));
    // in HBM the row is ordered by row_key (fpx_kernels.cuh); callers get it ascending, as the reference keeps it
    std::sort(out_docids, out_docids + len);
    return FPX_OK;
}

uint32_t fpx_default_min_score(uint64_t raw_query_len) { return (uint32_t)((raw_query_len + 19) / 20); }

} // extern "C"

namespace {

// total number of terms is needed to size the row workspace; the device API reads it back once
fpx_status device_total_terms(const uint64_t *d_offsets, uint64_t n_queries, cudaStream_t st, uint64_t *first,
                              uint64_t *last) {
    FPX_CUDA(cudaMemcpyAsync(first, d_offsets, 8, cudaMemcpyDeviceToHost, st));
    FPX_CUDA(cudaMemcpyAsync(last, d_offsets + n_queries, 8, cudaMemcpyDeviceToHost, st));
    FPX_CUDA(cudaStreamSynchronize(st));
    return FPX_OK;
}

} // namespace

extern "C" {

fpx_status fpx_search_batch_device_async(fpx_snapshot *s, uint64_t n_queries, uint64_t term_base, uint64_t n_terms_total,
                                         const uint32_t *d_terms, const uint64_t *d_term_offsets, const fpx_search_opts *d_opts,
                                         uint32_t k_stride, uint32_t *d_out_ids, uint32_t *d_out_scores, uint32_t *d_out_counts,
                                         uint32_t *d_status, void *cuda_stream) {
    if (!s) return set_error(FPX_INVALID_ARGUMENT, "null snapshot");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if (n_queries == 0) {
        if (d_status) FPX_CUDA(cudaMemsetAsync(d_status, 0, sizeof(uint32_t), st));
        return FPX_OK;
    }
    if (!d_term_offsets || !d_opts || !d_out_counts || (k_stride && (!d_out_ids || !d_out_scores)) || (n_terms_total && !d_terms))
        return set_error(FPX_INVALID_ARGUMENT, "null buffer");
    if (k_stride > FPX_MAX_RESULTS) return set_error(FPX_UNSUPPORTED, "k_stride exceeds FPX_MAX_RESULTS");
    fpx_ctx *ctx = s->ctx;
    FPX_CUDA(cudaSetDevice(ctx->device));
    Workspace *w = nullptr;
    fpx_status rc = acquire_workspace(ctx, &w);
    if (rc != FPX_OK) return rc;
    if (w->done_pending) { // its previous batch may still be running on another stream
        cudaStreamWaitEvent(st, w->done, 0);
        w->done_pending = false;
    }
    // term_offsets index d_terms absolutely; the row workspace is indexed relative to term_base
    rc = enqueue_batch(s, w, w->slot[0], st, st, n_queries, n_terms_total, d_terms ? d_terms + term_base : nullptr, d_term_offsets,
                       term_base, reinterpret_cast<const SearchOpts *>(d_opts), k_stride, d_out_ids, d_out_scores, d_out_counts);
    if (rc == FPX_OK) {
        // what the kernels raised (FPX_UNSUPPORTED / FPX_INVALID_ARGUMENT for a query outside the limits, else 0)
        if (d_status) cudaMemcpyAsync(d_status, &w->slot[0].counters->error, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st);
        cudaEventRecord(w->done, st);
        w->done_pending = true;
    }
    release_workspace(ctx, w);
    return rc;
}

fpx_status fpx_search_batch_device(fpx_snapshot *s, uint64_t n_queries, const uint32_t *d_terms,
                                   const uint64_t *d_term_offsets, const fpx_search_opts *d_opts, uint32_t k_stride,
                                   uint32_t *d_out_ids, uint32_t *d_out_scores, uint32_t *d_out_counts,
                                   void *cuda_stream) {
    if (!s) return set_error(FPX_INVALID_ARGUMENT, "null snapshot");
    if (n_queries == 0) return FPX_OK;
    if (!d_term_offsets) return set_error(FPX_INVALID_ARGUMENT, "null buffer");
    FPX_CUDA(cudaSetDevice(s->ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    uint64_t first = 0, last = 0; // the one host round trip of this variant: the size of the batch's terms
    fpx_status rc = device_total_terms(d_term_offsets, n_queries, st, &first, &last);
    if (rc != FPX_OK) return rc;
    if (last < first) return set_error(FPX_INVALID_ARGUMENT, "term_offsets not ascending");
    return fpx_search_batch_device_async(s, n_queries, first, last - first, d_terms, d_term_offsets, d_opts, k_stride, d_out_ids,
                                         d_out_scores, d_out_counts, nullptr, cuda_stream);
}

} // extern "C"

namespace {

struct PackedOut { // results as {count per query, (id, score) pairs back to back} instead of k_stride-wide arrays
    uint32_t *pairs;
    uint64_t capacity; // in pairs
    uint64_t n_pairs;  // needed
};

fpx_status search_batch_host_impl(fpx_snapshot *s, uint64_t n_queries, const uint32_t *terms, const uint64_t *term_offsets,
                                  const fpx_search_opts *opts, uint32_t k_stride, uint32_t *out_ids, uint32_t *out_scores,
                                  uint32_t *out_counts, PackedOut *packed, uint32_t timeout_ms);

// No exception crosses the C ABI: a failed host allocation inside the batch driver (chunk tables, helper threads) becomes
// FPX_OUT_OF_MEMORY, after the device has drained whatever the call had already enqueued (it may write the caller's
// arrays).  The call's workspace is not returned to the pool in that case.
fpx_status search_batch_host(fpx_snapshot *s, uint64_t n_queries, const uint32_t *terms, const uint64_t *term_offsets,
                             const fpx_search_opts *opts, uint32_t k_stride, uint32_t *out_ids, uint32_t *out_scores,
                             uint32_t *out_counts, PackedOut *packed, uint32_t timeout_ms = 0) {
    try {
        return search_batch_host_impl(s, n_queries, terms, term_offsets, opts, k_stride, out_ids, out_scores, out_counts, packed,
                                      timeout_ms);
    } catch (const std::bad_alloc &) {
        cudaDeviceSynchronize();
        return set_error(FPX_OUT_OF_MEMORY, "host allocation in the batch driver");
    } catch (const std::exception &ex) {
        cudaDeviceSynchronize();
        return set_error(FPX_OUT_OF_MEMORY, std::string("batch driver: ") + ex.what());
    }
}

fpx_status search_batch_host_impl(fpx_snapshot *s, uint64_t n_queries, const uint32_t *terms, const uint64_t *term_offsets,
                                  const fpx_search_opts *opts, uint32_t k_stride, uint32_t *out_ids, uint32_t *out_scores,
                                  uint32_t *out_counts, PackedOut *packed, uint32_t timeout_ms) {
    if (!s) return set_error(FPX_INVALID_ARGUMENT, "null snapshot");
    if (n_queries == 0) return FPX_OK;
    if (!term_offsets || !opts || !out_counts || (!packed && k_stride && (!out_ids || !out_scores)))
        return set_error(FPX_INVALID_ARGUMENT, "null buffer");
    if (k_stride > FPX_MAX_RESULTS) return set_error(FPX_UNSUPPORTED, "k_stride exceeds FPX_MAX_RESULTS");
    if (n_queries > 0x7FFFFFFFull) return set_error(FPX_INVALID_ARGUMENT, "batch too large (split it)");
    uint64_t longest = 0;
    for (uint64_t q = 0; q < n_queries; ++q) {
        if (term_offsets[q + 1] < term_offsets[q]) return set_error(FPX_INVALID_ARGUMENT, "term_offsets not ascending");
        longest = std::max(longest, term_offsets[q + 1] - term_offsets[q]);
    }
    if (longest > FPX_MAX_QUERY_TERMS) return set_error(FPX_UNSUPPORTED, "query has more than FPX_MAX_QUERY_TERMS terms");
    const bool no_long_queries = longest <= kWarpQueryTerms; // prepare_long_kernel has nothing to do
    const uint64_t t_first = term_offsets[0], nt_total = term_offsets[n_queries] - t_first;
    if (nt_total && !terms) return set_error(FPX_INVALID_ARGUMENT, "null terms");
    if (nt_total > 0xFFFFFFFFull) return set_error(FPX_INVALID_ARGUMENT, "batch too large (split it)");
    fpx_ctx *ctx = s->ctx;
    FPX_CUDA(cudaSetDevice(ctx->device));

    // In-order streams.  The copy stream moves the queries to the device chunk by chunk; the compute stream
    // runs the big kernels back to back (prepare -> persistent sketch kernel, chunk after chunk); the tail stream
    // runs each chunk's small trailing kernels (exact count-table kernels, result packing) in the gaps.  Chunks
    // alternate between two workspace slots.  The packed results ({count per query, (id, score) pairs back to
    // back}) are written by the GPU straight into mapped pinned host memory, and the calling thread unpacks
    // chunk c into the caller's arrays while the GPU works on the following chunks.
    // The first chunks are small so that the kernels start early, the last ones so that little unpacking is left
    // when the GPU is done.
    std::vector<uint64_t> bounds; // chunk c = queries [bounds[c], bounds[c+1])
    uint64_t max_nq = 0, max_nt = 0;
    {
        const uint64_t chunk = std::max<uint64_t>(ctx->chunk_queries, (n_queries + kErrSlots - 17) / (kErrSlots - 16));
        // every chunk costs a dozen kernel launches and a fill / drain of the persistent kernels (~0.05 ms), and the
        // copy engine feeds queries faster than the kernels answer them: a small first chunk so that the kernels start
        // early, then quickly growing ones
        // (FPX_DEBUG_ABLATE bits 16..18: other first-chunk sizes / growth factors for A/B runs)
        static const uint32_t kSchedDiv[8] = {16, 32, 32, 16, 64, 32, 64, 16}, kSchedGrow[8] = {4, 3, 2, 3, 3, 4, 4, 2};
        const uint32_t sched = (ctx->debug >> 16) & 7u;
        const uint64_t small = std::max<uint64_t>(1024, chunk / kSchedDiv[sched]);
        uint64_t q = 0, step = small;
        bounds.push_back(0);
        while (q < n_queries) {
            const uint64_t left = n_queries - q;
            uint64_t take = std::min(step, left);
            if (left - take < small) take = left;                        // no tiny remainder ...
            if (take == left && left >= 4 * small) take = left - small;  // ... but a small last chunk: little is left to
            const uint64_t q1 = q + take;                                // copy back when the GPU is done
            max_nq = std::max(max_nq, q1 - q);
            max_nt = std::max(max_nt, term_offsets[q1] - term_offsets[q]);
            q = q1;
            bounds.push_back(q);
            step = std::min(chunk, step * kSchedGrow[sched]);
        }
    }
    const uint64_t n_chunks = bounds.size() - 1;
    // Results: when the caller's arrays are pinned host memory the k_stride-wide device arrays go there by DMA,
    // chunk by chunk (the copy engine is otherwise idle and the calling thread has nothing to do).  Otherwise
    // (pageable memory) the GPU packs them to {count, (id, score) pairs} in the library's own pinned memory and
    // the calling thread scatters them.
    bool dma_out = packed == nullptr;
    if (dma_out) {
        const void *outs[3] = {out_counts, k_stride ? out_ids : out_counts, k_stride ? out_scores : out_counts};
        for (const void *p : outs) {
            cudaPointerAttributes at{};
            if (cudaPointerGetAttributes(&at, p) != cudaSuccess || at.type != cudaMemoryTypeHost) dma_out = false;
        }
        cudaGetLastError();
        if (ctx->debug & 4096u) dma_out = false; // FPX_DEBUG_ABLATE bit 12: force the packed path
    }
    Workspace *w = nullptr;
    fpx_status rc = acquire_workspace(ctx, &w);
    if (rc != FPX_OK) return rc;
    cudaStream_t st = w->stream, cs = w->copy_stream, ts = w->tail_stream;
    cudaError_t e = cudaSuccess;
    if (w->done_pending) { // the workspace's previous (device-API) batch may still be running on another stream
        cudaStreamWaitEvent(st, w->done, 0);
        w->done_pending = false;
    }
    // whole-call staging for what the copy stream writes and the GPU -> host packing writes; chunk-sized
    // buffers for everything that only the (in-order) compute stream touches
    e = w->d_terms.reserve(nt_total + 1);
    if (e == cudaSuccess) e = w->d_offsets.reserve(n_queries + n_chunks + 1);
    if (e == cudaSuccess) e = w->d_opts.reserve(n_queries);
    for (Workspace::Slot &sl : w->slot) {
        if (e == cudaSuccess) e = sl.d_ids.reserve(max_nq * (uint64_t)k_stride + 1);
        if (e == cudaSuccess) e = sl.d_scores.reserve(max_nq * (uint64_t)k_stride + 1);
        if (e == cudaSuccess) e = sl.d_counts.reserve(max_nq);
        if (e == cudaSuccess) e = sl.d_pack_offsets.reserve(max_nq + 1);
        if (e == cudaSuccess) e = sl.rows.reserve(max_nt + 1);
        if (e == cudaSuccess) e = sl.items.reserve(max_nq * kNumClasses);
        if (e == cudaSuccess) e = sl.long_queue.reserve(max_nq);
    }
    if (e == cudaSuccess && !dma_out) e = w->reserve_host(n_queries, n_queries * (uint64_t)k_stride + 1);
    if (e == cudaSuccess) e = w->reserve_events(n_chunks);
    if (e != cudaSuccess) {
        release_workspace(ctx, w);
        return cuda_fail(e, "staging buffers");
    }
    cudaStreamWaitEvent(cs, w->copy_fence, 0); // the previous call's kernels have read the staging buffers

    // FPX_DEBUG_ABLATE bit 11: print a timeline of this call (GPU events relative to the first, host clock)
    const bool tracing = (ctx->debug & 2048u) != 0;
    struct ChunkTrace {
        cudaEvent_t ev[6]; // start, after H2D, prepare, sketch, exact+wide, pack
        double host_col0, host_col1;
    };
    std::vector<ChunkTrace> tr(tracing ? n_chunks : 0);
    const auto host0 = std::chrono::steady_clock::now();
    auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host0).count(); };
    if (tracing)
        for (auto &t : tr)
            for (auto &ev : t.ev) cudaEventCreate(&ev);

    // ---- enqueue everything
    std::vector<uint64_t> pair_base(n_chunks + 1, 0); // worst-case start of chunk c's pairs in h_pairs
    for (uint64_t c = 0; c < n_chunks && rc == FPX_OK; ++c) {
        const uint64_t q0 = bounds[c], q1 = bounds[c + 1], nq = q1 - q0;
        const uint64_t t0 = term_offsets[q0], t1 = term_offsets[q1], nt = t1 - t0;
        pair_base[c + 1] = pair_base[c] + nq * (uint64_t)k_stride;
        uint64_t *d_off = w->d_offsets.p + q0 + c; // chunk c's nq+1 offsets (absolute term positions)
        if (tracing) cudaEventRecord(tr[c].ev[0], cs);
        {
            Timed t(ctx, cs, KK_H2D);
            if (nt) cudaMemcpyAsync(w->d_terms.p + (t0 - t_first), terms + t0, nt * 4, cudaMemcpyHostToDevice, cs);
            cudaMemcpyAsync(d_off, term_offsets + q0, (nq + 1) * 8, cudaMemcpyHostToDevice, cs);
            cudaMemcpyAsync(w->d_opts.p + q0, opts + q0, nq * sizeof(SearchOpts), cudaMemcpyHostToDevice, cs);
        }
        if (tracing) cudaEventRecord(tr[c].ev[1], cs);
        cudaEventRecord(w->h2d_done[c], cs);
        cudaStreamWaitEvent(st, w->h2d_done[c], 0);
        Workspace::Slot &sl = w->slot[c % kSlots];
        if (c >= (uint64_t)kSlots) cudaStreamWaitEvent(st, w->chunk_done[c - kSlots], 0); // the slot's previous tenant is done
        rc = enqueue_batch(s, w, sl, st, ts, nq, nt, w->d_terms.p + (t0 - t_first), d_off, t0, w->d_opts.p + q0, k_stride,
                           sl.d_ids.p, sl.d_scores.p, sl.d_counts.p, tracing ? &tr[c].ev[2] : nullptr, (uint32_t)c, no_long_queries);
        if (rc != FPX_OK) break;
        cudaStream_t rs = ts; // stream the results leave on
        if (dma_out) {        // DMA on its own stream: the next chunk's trailing kernels need not wait for it
            rs = w->d2h_stream;
            cudaEventRecord(w->tail_done, ts);
            cudaStreamWaitEvent(rs, w->tail_done, 0);
        }
        {
            Timed t(ctx, rs, KK_D2H);
            if (dma_out) {
                if (k_stride) {
                    cudaMemcpyAsync(out_ids + q0 * k_stride, sl.d_ids.p, nq * (uint64_t)k_stride * 4, cudaMemcpyDeviceToHost, rs);
                    cudaMemcpyAsync(out_scores + q0 * k_stride, sl.d_scores.p, nq * (uint64_t)k_stride * 4, cudaMemcpyDeviceToHost, rs);
                }
                cudaMemcpyAsync(out_counts + q0, sl.d_counts.p, nq * 4, cudaMemcpyDeviceToHost, rs);
            } else {
                launch_result_pack(sl.d_ids.p, sl.d_scores.p, sl.d_counts.p, sl.d_pack_offsets.p, (uint32_t)nq, k_stride,
                                   w->h_counts + q0, w->h_pairs + pair_base[c], rs);
            }
        }
        if (tracing) cudaEventRecord(tr[c].ev[5], rs);
        cudaEventRecord(w->chunk_done[c], rs);
        if (ctx->flags & FPX_FLAG_PROFILE) {
            std::lock_guard<std::mutex> lk(ctx->mu);
            ctx->prof.h2d_bytes += nt * 4 + (nq + 1) * 8 + nq * sizeof(SearchOpts);
        }
    }
    cudaEventRecord(w->copy_fence, st);
    const double host_enq = host_ms();

    // ---- collect: wait for chunk c, scatter its packed results into out_ids / out_scores / out_counts.
    // Scattering is memory-bound host work (two strided cache lines per query); big batches share it among a few
    // helper threads that live for the duration of the call.
    struct Job {
        uint64_t q0, q1;   // queries (absolute)
        const uint2 *pairs; // first pair of query q0
    };
    auto scatter = [&](const Job &j) {
        const uint2 *hp = j.pairs;
        for (uint64_t q = j.q0; q < j.q1; ++q) {
            const uint32_t n = w->h_counts[q];
            out_counts[q] = n;
            uint32_t *oi = out_ids + q * k_stride, *os = out_scores + q * k_stride;
            for (uint32_t i = 0; i < n; ++i) {
                oi[i] = hp[i].x;
                os[i] = hp[i].y;
            }
            hp += n;
        }
    };
    const unsigned n_helpers = (!dma_out && !packed && n_queries >= 16384) ? std::min(3u, ctx->host_threads > 1 ? ctx->host_threads - 1 : 0u) : 0u;
    std::mutex jm;
    std::condition_variable jcv;
    std::deque<Job> jobs;
    bool jobs_closed = false;
    std::vector<std::thread> helpers;
    for (unsigned i = 0; i < n_helpers; ++i)
        helpers.emplace_back([&] {
            for (;;) {
                Job j;
                {
                    std::unique_lock<std::mutex> lk(jm);
                    jcv.wait(lk, [&] { return !jobs.empty() || jobs_closed; });
                    if (jobs.empty()) return;
                    j = jobs.front();
                    jobs.pop_front();
                }
                scatter(j);
            }
        });
    uint32_t dev_err = 0;
    uint64_t d2h_bytes = 0;
    for (uint64_t c = 0; c < n_chunks && rc == FPX_OK; ++c) {
        if (tracing) tr[c].host_col0 = host_ms();
        if (timeout_ms) { // MultiIndex.zig:311-322: the request's deadline cancels the search (error.SearchTimeout)
            while ((e = cudaEventQuery(w->chunk_done[c])) == cudaErrorNotReady && host_ms() < (double)timeout_ms) std::this_thread::yield();
            if (e == cudaErrorNotReady) {
                cudaGetLastError();
                rc = set_error(FPX_TIMEOUT, "search timed out");
                break; // what is in flight still writes the library's own buffers: the streams are drained below
            }
        } else {
            e = cudaEventSynchronize(w->chunk_done[c]);
        }
        if (e != cudaSuccess) {
            rc = cuda_fail(e, "search batch");
            break;
        }
        dev_err |= w->h_error[c];
        const uint64_t q0 = bounds[c], q1 = bounds[c + 1], nq = q1 - q0;
        if (dma_out) {
            d2h_bytes += nq * (uint64_t)k_stride * 8 + nq * 4;
            if (tracing) tr[c].host_col1 = host_ms();
            continue;
        }
        const uint2 *hp = w->h_pairs + pair_base[c];
        if (packed) { // the chunk's counts and pairs as they are: two contiguous copies
            uint64_t o = 0;
            for (uint64_t q = q0; q < q1; ++q) o += w->h_counts[q];
            std::memcpy(out_counts + q0, w->h_counts + q0, nq * sizeof(uint32_t));
            const uint64_t room = packed->n_pairs < packed->capacity ? packed->capacity - packed->n_pairs : 0;
            if (std::min(o, room)) std::memcpy(packed->pairs + 2 * packed->n_pairs, hp, std::min(o, room) * sizeof(uint2));
            packed->n_pairs += o;
            d2h_bytes += nq * 4 + o * 8;
            if (tracing) tr[c].host_col1 = host_ms();
            continue;
        }
        const unsigned parts = (n_helpers && nq >= 4096) ? n_helpers + 1 : 1;
        Job mine{q0, q1, hp};
        if (parts > 1) {
            uint64_t o = 0;
            std::lock_guard<std::mutex> lk(jm);
            for (unsigned p = 0; p < parts; ++p) {
                const uint64_t a0 = q0 + nq * p / parts, a1 = q0 + nq * (p + 1) / parts;
                const Job j{a0, a1, hp + o};
                for (uint64_t q = a0; q < a1; ++q) o += w->h_counts[q];
                if (p + 1 < parts)
                    jobs.push_back(j);
                else
                    mine = j;
            }
            d2h_bytes += nq * 4 + o * 8;
        }
        if (parts > 1) jcv.notify_all();
        scatter(mine);
        if (parts == 1) {
            uint64_t o = 0;
            for (uint64_t q = q0; q < q1; ++q) o += w->h_counts[q];
            d2h_bytes += nq * 4 + o * 8;
        }
        if (tracing) tr[c].host_col1 = host_ms();
    }
    {
        std::lock_guard<std::mutex> lk(jm);
        jobs_closed = true;
    }
    jcv.notify_all();
    for (auto &t : helpers) t.join();
    if (ctx->flags & FPX_FLAG_PROFILE) {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->prof.d2h_bytes += d2h_bytes;
    }
    e = cudaStreamSynchronize(ts);
    if (e == cudaSuccess) e = cudaStreamSynchronize(w->d2h_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess && rc == FPX_OK) rc = cuda_fail(e, "search batch");
    if (tracing) {
        cudaDeviceSynchronize();
        std::fprintf(stderr, "[fpx trace] %llu chunks enqueued by %.3f ms (host)\n", (unsigned long long)n_chunks, host_enq);
        for (uint64_t c = 0; c < n_chunks; ++c) {
            float t[6] = {0, 0, 0, 0, 0, 0};
            for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&t[i], tr[0].ev[0], tr[c].ev[i]);
            std::fprintf(stderr, "[fpx trace] chunk %llu q=%llu | gpu: h2d %.3f-%.3f prepare %.3f sketch %.3f rest %.3f pack %.3f | host: collect %.3f-%.3f\n",
                         (unsigned long long)c, (unsigned long long)(bounds[c + 1] - bounds[c]), t[0], t[1], t[2], t[3], t[4], t[5],
                         tr[c].host_col0, tr[c].host_col1);
        }
        for (auto &t : tr)
            for (auto &ev : t.ev) cudaEventDestroy(ev);
    }
    release_workspace(ctx, w);
    if (rc == FPX_OK && dev_err) rc = set_error((fpx_status)dev_err, "a query in the batch is outside the device path's limits");
    return rc;
}

} // namespace

extern "C" {

fpx_status fpx_search_batch(fpx_snapshot *s, uint64_t n_queries, const uint32_t *terms, const uint64_t *term_offsets,
                            const fpx_search_opts *opts, uint32_t k_stride, uint32_t *out_ids, uint32_t *out_scores,
                            uint32_t *out_counts) {
    return search_batch_host(s, n_queries, terms, term_offsets, opts, k_stride, out_ids, out_scores, out_counts, nullptr);
}

fpx_status fpx_search_batch_timeout(fpx_snapshot *s, uint64_t n_queries, const uint32_t *terms, const uint64_t *term_offsets,
                                    const fpx_search_opts *opts, uint32_t k_stride, uint32_t *out_ids, uint32_t *out_scores,
                                    uint32_t *out_counts, uint32_t timeout_ms) {
    return search_batch_host(s, n_queries, terms, term_offsets, opts, k_stride, out_ids, out_scores, out_counts, nullptr, timeout_ms);
}

fpx_status fpx_search_batch_packed(fpx_snapshot *s, uint64_t n_queries, const uint32_t *terms, const uint64_t *term_offsets,
                                   const fpx_search_opts *opts, uint32_t k_stride, uint32_t *out_counts, uint32_t *out_pairs,
                                   uint64_t capacity_pairs, uint64_t *out_n_pairs) {
    if (!out_n_pairs || (capacity_pairs && !out_pairs)) return set_error(FPX_INVALID_ARGUMENT, "null buffer");
    PackedOut po{out_pairs, capacity_pairs, 0};
    *out_n_pairs = 0;
    const fpx_status rc = search_batch_host(s, n_queries, terms, term_offsets, opts, k_stride, nullptr, nullptr, out_counts, &po);
    *out_n_pairs = po.n_pairs;
    if (rc == FPX_OK && po.n_pairs > capacity_pairs)
        return set_error(FPX_INVALID_ARGUMENT, "out_pairs too small: *out_n_pairs pairs are needed (counts are complete)");
    return rc;
}

fpx_status fpx_search(fpx_snapshot *s, const uint32_t *terms, uint64_t n_terms, const fpx_search_opts *opts,
                      uint32_t *out_ids, uint32_t *out_scores, uint32_t capacity, uint32_t *out_count) {
    if (!opts || !out_count) return set_error(FPX_INVALID_ARGUMENT, "null argument");
    uint64_t offs[2] = {0, n_terms};
    uint32_t cap = std::min<uint32_t>(capacity, FPX_MAX_RESULTS);
    if (opts->max_results > cap && capacity > FPX_MAX_RESULTS)
        return set_error(FPX_UNSUPPORTED, "max_results exceeds FPX_MAX_RESULTS");
    *out_count = 0;
    return fpx_search_batch(s, 1, terms, offs, opts, cap, out_ids, out_scores, out_count);
}

fpx_status fpx_pack_results_device(uint64_t n_queries, uint32_t k_stride, const uint32_t *d_ids, const uint32_t *d_scores,
                                   const uint32_t *d_counts, uint32_t *d_packed, uint32_t capacity_pairs, void *cuda_stream) {
    if (n_queries == 0) return FPX_OK;
    if (n_queries > 0x7FFFFFFFull) return set_error(FPX_INVALID_ARGUMENT, "batch too large (split it)");
    if (!d_counts || !d_packed || (k_stride && (!d_ids || !d_scores))) return set_error(FPX_INVALID_ARGUMENT, "null buffer");
    // layout of d_packed (u32): [0, n) counts | [n, 2n] row offsets into the pairs, [2n] = pairs needed |
    //                           then capacity_pairs x (id, score)
    const uint32_t n = (uint32_t)n_queries;
    launch_result_pack(d_ids, d_scores, d_counts, d_packed + n, n, k_stride, d_packed,
                       reinterpret_cast<uint2 *>(d_packed + 2 * (size_t)n + 2), static_cast<cudaStream_t>(cuda_stream),
                       capacity_pairs);
    FPX_CUDA(cudaGetLastError());
    return FPX_OK;
}

fpx_status fpx_merge_packed_shards_device(uint32_t n_shards, uint64_t n_queries, const uint32_t *d_packed,
                                          uint64_t shard_stride_words, const fpx_search_opts *d_opts, uint32_t k_stride,
                                          uint32_t *d_out_ids, uint32_t *d_out_scores, uint32_t *d_out_counts, void *cuda_stream) {
    if (n_queries == 0) return FPX_OK;
    if (n_shards == 0 || n_shards > 32) return set_error(FPX_INVALID_ARGUMENT, "1..32 shards");
    if (n_queries > 0x7FFFFFFFull) return set_error(FPX_INVALID_ARGUMENT, "batch too large (split it)");
    if (!d_packed || !d_opts || !d_out_counts || (k_stride && (!d_out_ids || !d_out_scores)))
        return set_error(FPX_INVALID_ARGUMENT, "null buffer");
    if (shard_stride_words < 2 * n_queries + 2) return set_error(FPX_INVALID_ARGUMENT, "shard stride shorter than a packed header");
    launch_merge_packed_shards(d_packed, shard_stride_words, n_shards, (uint32_t)n_queries,
                               reinterpret_cast<const SearchOpts *>(d_opts), k_stride, d_out_ids, d_out_scores, d_out_counts,
                               static_cast<cudaStream_t>(cuda_stream));
    FPX_CUDA(cudaGetLastError());
    return FPX_OK;
}

fpx_status fpx_merge_shard_results(uint32_t n_shards, uint64_t n_queries, uint32_t k_stride, const uint32_t *ids,
                                   const uint32_t *scores, const uint32_t *counts, const fpx_search_opts *opts,
                                   uint32_t *out_ids, uint32_t *out_scores, uint32_t *out_counts) {
    if (!n_shards || !counts || !opts || !out_counts) return set_error(FPX_INVALID_ARGUMENT, "null argument");
    std::vector<unsigned long long> keys;
    for (uint64_t q = 0; q < n_queries; ++q) {
        keys.clear();
        for (uint32_t g = 0; g < n_shards; ++g) {
            const size_t base = ((size_t)g * n_queries + q) * k_stride;
            const uint32_t n = std::min(counts[(size_t)g * n_queries + q], k_stride);
            for (uint32_t i = 0; i < n; ++i)
                keys.push_back(((unsigned long long)(0xFFFFFFFFu - scores[base + i]) << 32) | ids[base + i]);
        }
        std::sort(keys.begin(), keys.end());
        const uint32_t k_eff = std::min(opts[q].max_results, k_stride);
        uint32_t ms = opts[q].min_score, n_out = 0;
        for (size_t i = 0; i < keys.size() && n_out < k_eff; ++i) {
            const uint32_t score = 0xFFFFFFFFu - (uint32_t)(keys[i] >> 32);
            if (score < ms) break;
            if (n_out == 0) ms = std::max(ms, (uint32_t)(score * opts[q].min_score_pct) / 100u);
            out_ids[q * k_stride + n_out] = (uint32_t)keys[i];
            out_scores[q * k_stride + n_out] = score;
            ++n_out;
        }
        out_counts[q] = n_out;
    }
    return FPX_OK;
}

fpx_status fpx_profile_reset(fpx_ctx *ctx) {
    if (!ctx) return set_error(FPX_INVALID_ARGUMENT, "null ctx");
    if (!ctx->host_only) {
        FPX_CUDA(cudaSetDevice(ctx->device));
        FPX_CUDA(cudaDeviceSynchronize());
        FPX_CUDA(cudaMemset(ctx->d_stats, 0, sizeof(DeviceStats)));
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (auto &p : ctx->pending) {
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    ctx->pending.clear();
    ctx->prof = fpx_profile{};
    return FPX_OK;
}

fpx_status fpx_set_chunk_queries(fpx_ctx *ctx, uint32_t chunk_queries) {
    if (!ctx || chunk_queries == 0) return set_error(FPX_INVALID_ARGUMENT, "null ctx or zero chunk");
    ctx->chunk_queries = chunk_queries;
    return FPX_OK;
}

fpx_status fpx_set_profile(fpx_ctx *ctx, int enabled) {
    if (!ctx) return set_error(FPX_INVALID_ARGUMENT, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->flags = enabled ? (ctx->flags | FPX_FLAG_PROFILE) : (ctx->flags & ~FPX_FLAG_PROFILE);
    return FPX_OK;
}

fpx_status fpx_debug_set(fpx_ctx *ctx, uint32_t bits) {
    if (!ctx) return set_error(FPX_INVALID_ARGUMENT, "null ctx");
    ctx->debug = bits;
    if (const uint32_t g = (bits >> 16) & 0xFFu) { // experiment: L2 fetch granularity hint (32 / 64 / 128 bytes)
        FPX_CUDA(cudaSetDevice(ctx->device));
        FPX_CUDA(cudaDeviceSynchronize());
        FPX_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g));
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        std::fprintf(stderr, "[fpx dbg] L2 fetch granularity limit now %zu\n", got);
    }
    return FPX_OK;
}

fpx_status fpx_profile_read(fpx_ctx *ctx, fpx_profile *out) {
    if (!ctx || !out) return set_error(FPX_INVALID_ARGUMENT, "null argument");
    if (!ctx->host_only) {
        FPX_CUDA(cudaSetDevice(ctx->device));
        FPX_CUDA(cudaDeviceSynchronize());
        std::vector<EventPair> ev;
        {
            std::lock_guard<std::mutex> lk(ctx->mu);
            ev.swap(ctx->pending);
        }
        for (auto &p : ev) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, p.a, p.b);
            switch (p.kind) {
            case KK_PREPARE: ctx->prof.prepare_ms += ms; ctx->prof.prepare_launches += 2; break;
            case KK_SEARCH: ctx->prof.search_ms += ms; ctx->prof.search_launches += 3; break;
            case KK_SKETCH: ctx->prof.sketch_ms += ms; ctx->prof.sketch_launches += 1; break;
            case KK_WIDE: ctx->prof.wide_ms += ms; ctx->prof.wide_launches += 1; break;
            case KK_H2D: ctx->prof.h2d_ms += ms; break;
            case KK_D2H: ctx->prof.d2h_ms += ms; break;
            }
            cudaEventDestroy(p.a);
            cudaEventDestroy(p.b);
        }
        DeviceStats ds{};
        FPX_CUDA(cudaMemcpy(&ds, ctx->d_stats, sizeof ds, cudaMemcpyDeviceToHost));
        ctx->prof.queries = ds.queries;
        ctx->prof.unique_terms = ds.unique_terms;
        ctx->prof.postings = ds.postings;
        ctx->prof.results = ds.results;
        ctx->prof.wide_queries = ds.wide_queries;
        ctx->prof.sketch_queries = ds.sketch_queries;
        ctx->prof.overflow_requeues = ds.overflow_requeues;
        if (ctx->debug & 512u) { // phase timers of CTA 0 (clock cycles per own query, from the start of the role's iteration)
            auto per = [&](int slot, int n_slot) { return (double)ds.dbg[slot] / (double)(ds.dbg[n_slot] ? ds.dbg[n_slot] : 1); };
            std::fprintf(stderr, "[fpx dbg] producers: wait_stage %.0f iter %.0f (%llu) | counter group 0: wait_full %.0f (parked warp %.0f) "
                                 "+count %.0f +group %.0f +readback %.0f (%llu) | resolver group 0: wait_counted %.0f +find %.0f end %.0f (%llu)\n",
                         per(0, 2), per(1, 2), ds.dbg[2], per(7, 10), per(12, 10), per(8, 10), per(13, 10), per(9, 10), ds.dbg[10],
                         per(3, 6), per(11, 6), per(5, 6), ds.dbg[6]);
            std::fprintf(stderr, "[fpx dbg] producer warp 0: issuing its copies %.0f | last copy issued -> stage complete as seen by counter group 0 "
                                 "(includes the group's own lateness) %.0f\n", per(14, 2), per(15, 10));
        }
    }
    *out = ctx->prof;
    return FPX_OK;
}

} // extern "C"
