// fpx_gpu_build.h — device-side snapshot build (see fpx_gpu_build.cu).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

namespace fpx {

struct GpuCsr { // what fpx_snapshot_commit needs to finish a snapshot
    uint32_t *d_docids = nullptr;                                         // padded rows (kept by the snapshot)
    uint32_t *d_terms = nullptr, *d_row_len = nullptr, *d_row_start4 = nullptr; // row directory (freed after the table build)
    uint64_t n_terms = 0, total4 = 0;
    uint64_t n_postings = 0, n_postings_total = 0, n_unreachable = 0, n_superseded = 0, n_out_of_range = 0;
    uint32_t max_row_len = 0, pad_id = 0;
    bool pad_spread = true;
    std::vector<uint32_t> h_terms, h_row_len, h_row_start4; // host copy of the directory (ascending terms)
};

// Last step of every snapshot build (device-built or uploaded from the host compiler): re-order each row of the
// padded CSR ascending by row_key(docid) (fpx_kernels.cuh).  *d_docids is replaced by a new allocation of n_words
// uint32 (the old one is freed).  h_row_start4 is the host copy of d_row_start4 (used to cut the work into pieces
// of < 2^31 words).  Padding keeps its place at the end of each row.
cudaError_t reorder_rows_by_key(uint32_t **d_docids, uint64_t n_words, const uint32_t *d_row_len,
                                const uint32_t *d_row_start4, const uint32_t *h_row_start4, uint64_t n_rows);

class GpuSnapshotBuilder {
  public:
    GpuSnapshotBuilder();
    ~GpuSnapshotBuilder();
    GpuSnapshotBuilder(const GpuSnapshotBuilder &) = delete;
    GpuSnapshotBuilder &operator=(const GpuSnapshotBuilder &) = delete;

    std::string error;
    bool oom = false, unsupported = false;

    // same contracts as SnapshotCompiler (fpx_snapshot_host.h); the inputs are uploaded during the call
    bool add_file_segment(uint64_t commit_id, uint64_t merges, uint32_t min_doc_id, uint32_t block_size, const uint8_t *blocks,
                          uint64_t num_blocks, const uint32_t *block_index, const uint32_t *doc_ids, uint64_t n_docs);
    bool add_memory_segment(uint64_t commit_id, uint64_t merges, const uint64_t *items, uint64_t n_items,
                            const uint32_t *doc_ids, uint64_t n_docs);
    void set_doc_range(uint32_t lo, uint32_t hi) {
        lo_ = lo;
        hi_ = hi;
    }
    uint32_t doc_lo() const { return lo_; }
    uint32_t doc_hi() const { return hi_; }
    size_t n_segments() const { return segs_.size(); }
    bool build(GpuCsr &out);

  private:
    struct Segment;
    struct SegCsr;
    // decode, flag and select segment si; its raw bytes are freed on the way (newest: the liveness map, ctr: BuildCounters)
    bool build_segment(size_t si, bool multi, const void *newest, void *ctr, SegCsr &out);
    bool fail(const char *m) {
        error = m;
        return false;
    }
    bool fail_unsupported(const char *m) {
        unsupported = true;
        error = m;
        return false;
    }
    bool cuda_fail(cudaError_t e, const char *what);
    bool check_order(uint64_t commit_id, bool is_file);
    std::vector<Segment *> segs_;
    uint32_t lo_ = 0, hi_ = 0;
    bool poisoned_ = false; // an add_*_segment call failed half way: commit must fail
};

} // namespace fpx
