// fpx_batcher.cu — request micro-batcher behind the single-query seam (include/fpx.h, fpx_batcher_*).
//
// The reference answers one request per coroutine (MultiIndex.search, src/MultiIndex.zig:287-330): acquire a
// reader on the current snapshot, arm the request's timeout (AutoCancel, :314-316 -> error.SearchTimeout), run
// IndexReader.search, copy the results out.  A GPU wants batches, so this object sits where that call is made:
// many host threads call fpx_batcher_search concurrently; one worker thread keeps taking whatever has
// accumulated (up to max_batch, waiting at most max_wait_us for a batch to fill while the GPU is idle), runs it
// through fpx_search_batch on the snapshot that was current when the batch was formed, and wakes the callers.
// While a batch is on the GPU the next one accumulates, so the batch size adapts to the load by itself.
//
// Semantics kept from the reference:
//   * the snapshot is pinned for the duration of a batch (acquireReader / SharedPtr, Index.zig:430-434);
//     fpx_batcher_set_snapshot is the swapSnapshot hook (Index.zig:469-485) and never blocks on searches;
//   * options ride in the request, already resolved (fpx_default_min_score for MultiIndex.zig:304);
//   * timeout in milliseconds, 0 = no bound; an expired request returns FPX_TIMEOUT (error.SearchTimeout -> 503,
//     server.zig:111-126) and its slot is dropped when the batch it joined completes.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fpx.h"

namespace fpx {
void set_last_error(const std::string &msg); // fpx_api.cu
}

namespace {

struct Request {
    std::vector<uint32_t> terms;
    fpx_search_opts opts{};
    uint32_t capacity = 0;
    // filled by the worker
    std::vector<uint32_t> ids, scores;
    uint32_t count = 0;
    fpx_status status = FPX_OK;
    std::string error;
    bool done = false;
    std::condition_variable cv; // guarded by the batcher's mutex
};

} // namespace

struct fpx_batcher {
    fpx_ctx *ctx = nullptr;
    fpx_batcher_config cfg{};
    std::mutex mu;
    std::condition_variable work_cv;
    std::deque<std::shared_ptr<Request>> queue;
    fpx_snapshot *snapshot = nullptr; // current; the batcher holds one reference
    bool stopping = false;
    std::thread worker;
    fpx_batcher_stats stats{};

    void run();
};

void fpx_batcher::run() {
    std::vector<std::shared_ptr<Request>> batch;
    std::vector<uint32_t> terms, ids, scores, counts;
    std::vector<uint64_t> offsets;
    std::vector<fpx_search_opts> opts;
    for (;;) {
        fpx_snapshot *snap = nullptr;
        {
            std::unique_lock<std::mutex> lk(mu);
            work_cv.wait(lk, [&] { return stopping || !queue.empty(); });
            if (stopping && queue.empty()) return;
            // the GPU is idle: give the batch a moment to fill
            if (queue.size() < cfg.max_batch && cfg.max_wait_us) {
                const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(cfg.max_wait_us);
                work_cv.wait_until(lk, until, [&] { return stopping || queue.size() >= cfg.max_batch; });
            }
            batch.clear();
            while (!queue.empty() && batch.size() < cfg.max_batch) {
                batch.push_back(std::move(queue.front()));
                queue.pop_front();
            }
            snap = snapshot;
            if (snap) fpx_snapshot_acquire(snap); // pinned while the batch runs (Index.zig:430-434)
        }
        fpx_status rc = FPX_OK;
        std::string err;
        uint32_t k_stride = 1;
        if (!snap) {
            rc = FPX_INVALID_ARGUMENT;
            err = "no snapshot installed";
        } else {
            terms.clear();
            offsets.assign(1, 0);
            opts.clear();
            for (auto &r : batch) {
                terms.insert(terms.end(), r->terms.begin(), r->terms.end());
                offsets.push_back(terms.size());
                opts.push_back(r->opts);
                k_stride = std::max(k_stride, std::min(r->capacity, r->opts.max_results));
            }
            k_stride = std::min<uint32_t>(k_stride, FPX_MAX_RESULTS);
            const size_t n = batch.size();
            ids.resize(n * k_stride);
            scores.resize(n * k_stride);
            counts.assign(n, 0);
            rc = fpx_search_batch(snap, n, terms.empty() ? nullptr : terms.data(), offsets.data(), opts.data(), k_stride,
                                  ids.data(), scores.data(), counts.data());
            if (rc != FPX_OK) err = fpx_last_error_message();
            if (rc == FPX_UNSUPPORTED && n > 1) {
                // one query outside the device path's limits must not fail its neighbours: answer them one by one
                for (size_t i = 0; i < n; ++i) {
                    uint64_t o2[2] = {0, batch[i]->terms.size()};
                    batch[i]->status = fpx_search_batch(snap, 1, batch[i]->terms.data(), o2, &opts[i], k_stride,
                                                        ids.data() + i * k_stride, scores.data() + i * k_stride, &counts[i]);
                    if (batch[i]->status != FPX_OK) batch[i]->error = fpx_last_error_message();
                }
                rc = FPX_OK;
            }
            fpx_snapshot_release(snap);
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            for (size_t i = 0; i < batch.size(); ++i) {
                Request &r = *batch[i];
                if (rc != FPX_OK) {
                    r.status = rc;
                    r.error = err;
                } else if (r.status == FPX_OK) {
                    r.count = std::min(counts[i], std::min(r.capacity, k_stride));
                    r.ids.assign(ids.begin() + i * k_stride, ids.begin() + i * k_stride + r.count);
                    r.scores.assign(scores.begin() + i * k_stride, scores.begin() + i * k_stride + r.count);
                }
                r.done = true;
                r.cv.notify_one();
            }
            stats.batches += 1;
            stats.queries += batch.size();
            stats.max_batch_seen = std::max<uint64_t>(stats.max_batch_seen, batch.size());
        }
        batch.clear();
    }
}

extern "C" {

fpx_status fpx_batcher_create(fpx_ctx *ctx, const fpx_batcher_config *config, fpx_batcher **out) {
    if (!ctx || !out) {
        fpx::set_last_error("null argument");
        return FPX_INVALID_ARGUMENT;
    }
    *out = nullptr;
    fpx_batcher *b = new (std::nothrow) fpx_batcher();
    if (!b) return FPX_OUT_OF_MEMORY;
    b->ctx = ctx;
    b->cfg.max_batch = 4096;
    b->cfg.max_wait_us = 100;
    if (config) {
        if (config->max_batch) b->cfg.max_batch = config->max_batch;
        b->cfg.max_wait_us = config->max_wait_us;
    }
    try {
        b->worker = std::thread([b] { b->run(); });
    } catch (...) {
        delete b;
        fpx::set_last_error("cannot start the batcher thread");
        return FPX_OUT_OF_MEMORY;
    }
    *out = b;
    return FPX_OK;
}

fpx_status fpx_batcher_set_snapshot(fpx_batcher *b, fpx_snapshot *snapshot) {
    if (!b) return FPX_INVALID_ARGUMENT;
    if (snapshot) {
        const fpx_status rc = fpx_snapshot_acquire(snapshot);
        if (rc != FPX_OK) return rc;
    }
    fpx_snapshot *old;
    {
        std::lock_guard<std::mutex> lk(b->mu);
        old = b->snapshot;
        b->snapshot = snapshot;
    }
    if (old) fpx_snapshot_release(old); // batches in flight hold their own reference (Segments.deinit, Index.zig:57-63)
    return FPX_OK;
}

fpx_status fpx_batcher_search(fpx_batcher *b, const uint32_t *terms, uint64_t n_terms, const fpx_search_opts *opts,
                              uint32_t timeout_ms, uint32_t *out_ids, uint32_t *out_scores, uint32_t capacity,
                              uint32_t *out_count) {
    if (!b || !opts || !out_count || (n_terms && !terms) || (capacity && (!out_ids || !out_scores))) {
        fpx::set_last_error("null argument");
        return FPX_INVALID_ARGUMENT;
    }
    *out_count = 0;
    if (n_terms > FPX_MAX_QUERY_TERMS) {
        fpx::set_last_error("query has more than FPX_MAX_QUERY_TERMS terms");
        return FPX_UNSUPPORTED;
    }
    std::shared_ptr<Request> r;
    try {
        r = std::make_shared<Request>();
        r->terms.assign(terms, terms + n_terms);
    } catch (const std::bad_alloc &) {
        return FPX_OUT_OF_MEMORY;
    }
    r->opts = *opts;
    r->capacity = capacity;
    std::unique_lock<std::mutex> lk(b->mu);
    if (b->stopping) {
        fpx::set_last_error("batcher is shutting down");
        return FPX_INVALID_ARGUMENT;
    }
    b->queue.push_back(r);
    b->work_cv.notify_one();
    if (timeout_ms == 0) { // MultiIndex.zig:315: 0 = no bound
        r->cv.wait(lk, [&] { return r->done; });
    } else if (!r->cv.wait_for(lk, std::chrono::milliseconds(timeout_ms), [&] { return r->done; })) {
        // error.SearchTimeout (MultiIndex.zig:320).  If the request is still queued, take it out; if a batch
        // already has it, the worker's shared_ptr keeps the slot alive until that batch is done.
        for (auto it = b->queue.begin(); it != b->queue.end(); ++it)
            if (it->get() == r.get()) {
                b->queue.erase(it);
                break;
            }
        b->stats.timeouts += 1;
        fpx::set_last_error("search timed out");
        return FPX_TIMEOUT;
    }
    if (r->status != FPX_OK) {
        fpx::set_last_error(r->error);
        return r->status;
    }
    const uint32_t n = std::min(r->count, capacity);
    if (n) {
        std::memcpy(out_ids, r->ids.data(), n * sizeof(uint32_t));
        std::memcpy(out_scores, r->scores.data(), n * sizeof(uint32_t));
    }
    *out_count = n;
    return FPX_OK;
}

fpx_status fpx_batcher_get_stats(fpx_batcher *b, fpx_batcher_stats *out) {
    if (!b || !out) return FPX_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(b->mu);
    *out = b->stats;
    return FPX_OK;
}

void fpx_batcher_destroy(fpx_batcher *b) {
    if (!b) return;
    {
        std::lock_guard<std::mutex> lk(b->mu);
        b->stopping = true;
    }
    b->work_cv.notify_all();
    if (b->worker.joinable()) b->worker.join(); // answers what is still queued first
    if (b->snapshot) fpx_snapshot_release(b->snapshot);
    delete b;
}

} // extern "C"
