// fpx_gpu_build.cu — snapshot build on the device.
//
// The reference swaps snapshots on every update (Index.zig:469-485), so rebuilding the HBM mirror has to keep pace
// with checkpoints and merges.  The host compiler (fpx_snapshot_host.h) decodes ~10^9 postings on CPU threads;
// here the segments' raw bytes are uploaded as they are and everything else happens on the GPU:
//   decode_blocks_kernel   one warp per 512-byte block: StreamVByte 0124 (hash deltas) and 1234 (docid deltas
//                          that restart at min_doc_id on every hash change) — block.zig:66-312,
//                          streamvbyte.zig:76-412 — with warp scans instead of pshufb tables
//   run_state / reach      the per-hash scan caps of FileSegment.search (FileSegment.zig:25-26, 156-175): the run
//                          that is open at a block's first item is tracked by a scan over blocks, every other run
//                          starts inside its block and is reachable
//   liveness               "no newer segment mentions the id" (Index.zig:133-149 + common.zig:121-129) through a
//                          device hash map id -> newest mentioning segment
//   select / sort / rows   CUB select of the kept postings, radix sort of (hash << 32 | docid) when several
//                          segments feed the snapshot, run-length encoding into terms + row lengths, padded rows
// CUB is library code used for the non-hot build path only.  The result is identical to the host compiler's
// (tests/test_gpu_parity.py compares rows and runs the whole parity suite on GPU-built snapshots).
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include <algorithm>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "fpx_codec.h"
#include "fpx_gpu_build.h"
#include "fpx_kernels.cuh"

namespace fpx {

namespace {

constexpr uint32_t kMaxBlocksPerHashDev = 4;   // FileSegment.zig:25
constexpr uint32_t kMaxDocsPerHashDev = 1000;  // FileSegment.zig:26
constexpr int kDecodeWarps = 4;

struct BlockMeta { // per block, written by the decoder
    uint32_t first_hash, last_hash;
    uint32_t n_items;
    uint32_t last_run_start; // index (within the block) of the first item whose hash is last_hash
};

struct RunState { // the run that is open at some point of the block sequence
    unsigned long long start; // item index (within the segment) of the run's first item
    uint32_t block;           // block that holds it
    uint32_t is_const;        // scan: 1 = this element sets the state, 0 = it passes the previous one on
};
struct RunStateOp {
    __device__ __forceinline__ RunState operator()(const RunState &a, const RunState &b) const { return b.is_const ? b : a; }
};

struct BuildCounters {
    unsigned long long unreachable, superseded, out_of_range;
    uint32_t error; // 1 corrupt block, 2 block_index mismatch, 3 hashes not ascending, 4 items not sorted,
                    // 5 docs map holds 0xFFFFFFFF, 6 row longer than u32
    uint32_t max_row_len;
};

__global__ void block_counts_kernel(const uint8_t *blocks, uint32_t block_size, unsigned long long n_blocks, uint32_t *counts,
                                    BuildCounters *ctr) {
    const unsigned long long b = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint8_t *p = blocks + b * block_size;
    const uint32_t n = (uint32_t)p[4] | ((uint32_t)p[5] << 8);
    if (n == 0 || n > kWriterWindow) ctr->error = 1; // empty block inside the segment / impossible count
    counts[b] = n;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= (uint32_t)o) v += y;
    }
    return v;
}

// Decode one StreamVByte column of `n` values (ceil(n/4) control bytes at `ctrl`, data right after) into out[].
// kVariant1234: lengths {1,2,3,4}, else {0,1,2,4}.  Returns false if the data runs past `limit`.
template <bool kVariant1234>
__device__ bool decode_column(const uint8_t *blk, uint32_t ctrl, uint32_t n, uint32_t limit, uint32_t *out, uint32_t lane) {
    const uint32_t quads = (n + 3) >> 2;
    uint32_t data = ctrl + quads;
    bool ok = data <= limit;
    for (uint32_t q0 = 0; q0 < quads; q0 += 32) {
        const uint32_t q = q0 + lane;
        const uint32_t c = (q < quads && ok) ? blk[ctrl + q] : 0u;
        uint32_t len[4], tot = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t code = (c >> (2 * k)) & 3u;
            len[k] = kVariant1234 ? code + 1u : (code == 3u ? 4u : code);
            if (q >= quads) len[k] = 0;
            tot += len[k];
        }
        const uint32_t incl = warp_incl_scan(tot, lane);
        uint32_t off = data + incl - tot;
        if (q < quads) {
            if (off + tot > limit) ok = false;
            else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t v = 0;
                    for (uint32_t j = 0; j < len[k]; ++j) v |= (uint32_t)blk[off + j] << (8 * j);
                    off += len[k];
                    if (4 * q + k < n) out[4 * q + k] = v;
                }
            }
        }
        data += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    return __all_sync(0xFFFFFFFFu, ok);
}

// One warp per block.  smem per warp: the block's bytes + two u32 arrays of max_items.
__global__ void __launch_bounds__(kDecodeWarps * 32)
decode_blocks_kernel(const uint8_t *blocks, uint32_t block_size, unsigned long long n_blocks, const unsigned long long *blk_off,
                     const uint32_t *block_index, uint32_t min_doc_id, uint32_t max_items, unsigned long long *keys,
                     BlockMeta *meta, BuildCounters *ctr) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t per_warp = ((size_t)block_size + 15) / 16 * 16 + (size_t)max_items * 8;
    uint8_t *blk = smem + warp * per_warp;
    uint32_t *hs = reinterpret_cast<uint32_t *>(blk + ((size_t)block_size + 15) / 16 * 16);
    uint32_t *ds = hs + max_items;
    for (unsigned long long b = blockIdx.x * (unsigned long long)kDecodeWarps + warp; b < n_blocks;
         b += (unsigned long long)gridDim.x * kDecodeWarps) {
        const uint4 *src = reinterpret_cast<const uint4 *>(blocks + b * block_size); // block_size is a multiple of 16
        for (uint32_t i = lane; i < block_size / 16; i += 32) reinterpret_cast<uint4 *>(blk)[i] = src[i];
        __syncwarp();
        const uint32_t min_hash = (uint32_t)blk[0] | ((uint32_t)blk[1] << 8) | ((uint32_t)blk[2] << 16) | ((uint32_t)blk[3] << 24);
        const uint32_t n = (uint32_t)blk[4] | ((uint32_t)blk[5] << 8);
        const uint32_t doff = (uint32_t)blk[6] | ((uint32_t)blk[7] << 8);
        bool ok = n > 0 && n <= max_items && kBlockHeaderBytes + doff <= block_size;
        if (ok) ok = decode_column<false>(blk, kBlockHeaderBytes, n, block_size, hs, lane);
        if (ok) ok = decode_column<true>(blk, kBlockHeaderBytes + doff, n, block_size, ds, lane);
        if (!ok) { // reported by the host after the pass; the kernels that follow see an empty block
            if (lane == 0) {
                ctr->error = 1;
                meta[b] = BlockMeta{};
            }
            __syncwarp();
            continue;
        }
        __syncwarp();
        // hashes: prefix sum of the deltas from min_hash; docids: the delta chain restarts at min_doc_id at the
        // block's first item and whenever the hash changes (block.zig:235-265, 438-495)
        uint32_t h_carry = min_hash, d_carry = min_doc_id, last_start = 0;
        const unsigned long long base = blk_off[b];
        for (uint32_t i0 = 0; i0 < n; i0 += 32) {
            const uint32_t i = i0 + lane;
            const uint32_t hd = i < n ? hs[i] : 0u;
            const bool head = i < n && (i == 0 || hd != 0u);
            const uint32_t h = h_carry + warp_incl_scan(hd, lane);
            uint32_t v = i < n ? ds[i] + (head ? min_doc_id : 0u) : 0u;
            uint32_t f = head ? 1u : 0u; // "a head lies in (start of my summed range, me]"
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v2 = __shfl_up_sync(0xFFFFFFFFu, v, o), f2 = __shfl_up_sync(0xFFFFFFFFu, f, o);
                if (lane >= (uint32_t)o && !f) {
                    v += v2;
                    f |= f2;
                }
            }
            if (!f) v += d_carry; // my run started in an earlier chunk of this block
            if (i < n) keys[base + i] = ((unsigned long long)h << 32) | v;
            const uint32_t hm = __ballot_sync(0xFFFFFFFFu, head);
            if (hm) last_start = i0 + (31 - __clz(hm));
            const uint32_t last_lane = min(31u, n - 1 - i0);
            h_carry = __shfl_sync(0xFFFFFFFFu, h, last_lane);
            d_carry = __shfl_sync(0xFFFFFFFFu, v, last_lane);
        }
        if (lane == 0) {
            BlockMeta m;
            m.first_hash = min_hash + hs[0];
            m.last_hash = h_carry;
            m.n_items = n;
            m.last_run_start = last_start;
            meta[b] = m;
            if (block_index[b] != h_carry) ctr->error = 2; // filefmt.zig:117: block_index[b] = hash of the last item
        }
        __syncwarp();
    }
}

// element b of the scan over blocks: the run open at the END of block b
__global__ void run_state_kernel(const BlockMeta *meta, const unsigned long long *blk_off, unsigned long long n_blocks,
                                 RunState *st, BuildCounters *ctr) {
    const unsigned long long b = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const BlockMeta m = meta[b];
    const bool is_new = b == 0 || meta[b - 1].last_hash != m.first_hash;
    if (b > 0 && meta[b - 1].last_hash > m.first_hash) ctr->error = 3; // hashes must ascend across the segment
    const bool single = m.first_hash == m.last_hash;
    RunState s;
    if (!single) {
        s.start = blk_off[b] + m.last_run_start;
        s.block = (uint32_t)b;
        s.is_const = 1;
    } else if (is_new) {
        s.start = blk_off[b];
        s.block = (uint32_t)b;
        s.is_const = 1;
    } else {
        s.start = 0;
        s.block = 0;
        s.is_const = 0; // passes the previous block's state on
    }
    st[b] = s;
}

struct NewestMap { // open addressing: id -> 1-based index of the newest segment that mentions it
    uint32_t *keys;
    uint32_t *vals;
    uint32_t mask;
};
__device__ __forceinline__ uint32_t newest_slot(uint32_t id, uint32_t mask) { return (id * 0x9E3779B1u) & mask; }

__global__ void newest_insert_kernel(NewestMap m, const uint32_t *ids, unsigned long long n, uint32_t seg1, BuildCounters *ctr) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = ids[i];
    if (id == 0xFFFFFFFFu) { // the map's empty marker: the host build handles such a docs map
        ctr->error = 5;
        return;
    }
    uint32_t s = newest_slot(id, m.mask);
    for (;;) {
        const uint32_t old = atomicCAS(m.keys + s, 0xFFFFFFFFu, id);
        if (old == 0xFFFFFFFFu || old == id) {
            atomicMax(m.vals + s, seg1);
            return;
        }
        s = (s + 1) & m.mask;
    }
}
__device__ __forceinline__ uint32_t newest_lookup(const NewestMap &m, uint32_t id) {
    uint32_t s = newest_slot(id, m.mask);
    for (;;) {
        const uint32_t k = m.keys[s];
        if (k == id) return m.vals[s];
        if (k == 0xFFFFFFFFu) return 0;
        s = (s + 1) & m.mask;
    }
}

// keep flags of one segment's items.  File segments: one warp per block (the caps concern the block's first run
// only); memory segments: blocks == nullptr, plain grid-stride.
__global__ void keep_flags_kernel(const unsigned long long *keys, unsigned long long n_items, const BlockMeta *meta,
                                  const unsigned long long *blk_off, const RunState *st, unsigned long long n_blocks,
                                  NewestMap newest, uint32_t seg1, uint32_t use_newest, uint32_t lo, uint32_t hi,
                                  uint8_t *flags, BuildCounters *ctr) {
    unsigned long long unreach = 0, sup = 0, oor = 0;
    const bool ranged = !(lo == 0 && hi == 0);
    auto classify = [&](unsigned long long key, bool reachable) -> uint8_t {
        if (!reachable) {
            ++unreach;
            return 0;
        }
        const uint32_t id = (uint32_t)key;
        if (use_newest && newest_lookup(newest, id) > seg1) { // a newer segment mentions the id
            ++sup;
            return 0;
        }
        if (ranged && !(id >= lo && (hi == 0u || id < hi))) { // hi == 0 with lo > 0: open-ended
            ++oor;
            return 0;
        }
        return 1;
    };
    if (meta) {
        const uint32_t lane = threadIdx.x & 31;
        const unsigned long long warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
        for (unsigned long long b = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < n_blocks; b += warps) {
            const BlockMeta m = meta[b];
            const unsigned long long base = blk_off[b];
            // the run open at this block's first item
            RunState open;
            const bool is_new = b == 0 || meta[b - 1].last_hash != m.first_hash;
            if (is_new) {
                open.start = base;
                open.block = (uint32_t)b;
            } else {
                open = st[b - 1];
            }
            const bool first_run_ok = ((uint32_t)b - open.block) < kMaxBlocksPerHashDev && (base - open.start) <= kMaxDocsPerHashDev;
            for (uint32_t i = lane; i < m.n_items; i += 32) {
                const unsigned long long key = keys[base + i];
                const bool reachable = (uint32_t)(key >> 32) != m.first_hash || first_run_ok;
                flags[base + i] = classify(key, reachable);
            }
        }
    } else {
        for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n_items;
             i += (unsigned long long)gridDim.x * blockDim.x) {
            const unsigned long long key = keys[i];
            if (i > 0 && keys[i - 1] > key) ctr->error = 4; // MemorySegment items are sorted (segment.zig:87-106)
            flags[i] = classify(key, true);
        }
    }
    // warp-reduce the statistics
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unreach += __shfl_xor_sync(0xFFFFFFFFu, unreach, o);
        sup += __shfl_xor_sync(0xFFFFFFFFu, sup, o);
        oor += __shfl_xor_sync(0xFFFFFFFFu, oor, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (unreach) atomicAdd(&ctr->unreachable, unreach);
        if (sup) atomicAdd(&ctr->superseded, sup);
        if (oor) atomicAdd(&ctr->out_of_range, oor);
    }
}

struct Quads { // 64-bit so that the scan accumulates in 64 bits
    __host__ __device__ __forceinline__ unsigned long long operator()(const uint32_t &len) const { return (len + 3) >> 2; }
};

struct HeadOfRun { // posting i starts a row of its segment
    const unsigned long long *keys;
    __host__ __device__ __forceinline__ uint8_t operator()(const unsigned long long &i) const {
        return (i == 0 || (uint32_t)(keys[i] >> 32) != (uint32_t)(keys[i - 1] >> 32)) ? 1 : 0;
    }
};
struct HeadOfTerm { // sorted entry i starts a merged row
    const uint32_t *terms;
    __host__ __device__ __forceinline__ uint8_t operator()(const uint32_t &i) const {
        return (i == 0 || terms[i] != terms[i - 1]) ? 1 : 0;
    }
};
struct GatherLen {
    const uint32_t *lens;
    __host__ __device__ __forceinline__ unsigned long long operator()(const uint32_t &entry) const { return lens[entry]; }
};

__global__ void iota_kernel(uint32_t *v, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[i] = (uint32_t)i;
}

// row j of a segment: its term and length from the row starts
__global__ void segment_rows_kernel(const unsigned long long *kept, unsigned long long n_kept, const unsigned long long *first,
                                    unsigned long long nt, uint32_t *terms, uint32_t *lens, BuildCounters *ctr) {
    const unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (j >= nt) return;
    const unsigned long long f = first[j], e = j + 1 < nt ? first[j + 1] : n_kept;
    terms[j] = (uint32_t)(kept[f] >> 32);
    if (e - f > 0xFFFFFFFFull) ctr->error = 6;
    lens[j] = (uint32_t)(e - f);
}

// merged row g = sorted entries [head_pos[g], head_pos[g+1]); E = running sum of the entries' lengths
__global__ void merged_rows_kernel(const uint32_t *sorted_terms, const unsigned long long *E, const uint32_t *head_pos,
                                   unsigned long long n_rows, uint32_t *terms, uint32_t *lens, BuildCounters *ctr) {
    const unsigned long long g = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (g >= n_rows) return;
    const uint32_t h = head_pos[g];
    const unsigned long long len = E[head_pos[g + 1]] - E[h];
    terms[g] = sorted_terms[h];
    if (len > 0xFFFFFFFFull) ctr->error = 6;
    lens[g] = (uint32_t)len;
}

// every entry of merged row g: its place in the padded array (in words), and entry -> sorted position
__global__ void entry_places_kernel(const uint32_t *sorted_entry, const unsigned long long *E, const uint32_t *head_pos,
                                    const unsigned long long *start4, unsigned long long n_rows, unsigned long long *dst_word,
                                    uint32_t *inv) {
    const unsigned long long g = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (g >= n_rows) return;
    const uint32_t h0 = head_pos[g], h1 = head_pos[g + 1];
    const unsigned long long base = start4[g] * 4ull, e0 = E[h0];
    for (uint32_t i = h0; i < h1; ++i) {
        dst_word[i] = base + (E[i] - e0);
        inv[sorted_entry[i]] = i;
    }
}

// one warp per row of one segment: copy its docids to their place in the merged, padded rows
__global__ void copy_entries_kernel(const uint32_t *seg_docids, const unsigned long long *first, unsigned long long n_kept,
                                    unsigned long long nt, const uint32_t *inv /* of this segment's entries */,
                                    const unsigned long long *dst_word, uint32_t *docids) {
    const uint32_t lane = threadIdx.x & 31;
    const unsigned long long warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long j = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < nt; j += warps) {
        const unsigned long long f = first[j], e = j + 1 < nt ? first[j + 1] : n_kept;
        uint32_t *dst = docids + dst_word[inv[j]];
        for (unsigned long long k = lane; k < e - f; k += 32) dst[k] = seg_docids[f + k];
    }
}

// padding of every row's last 16-byte granule = unused docids above max_live, varying from row to row
// (fpx_snapshot_host.h, same formula)
__global__ void pad_rows_kernel(const uint32_t *row_len, const uint32_t *row_start4, unsigned long long n_rows, uint32_t pad_base,
                                uint32_t *docids, BuildCounters *ctr) {
    uint32_t longest = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_rows; g += stride) {
        const uint32_t len = row_len[g];
        uint32_t *dst = docids + (size_t)row_start4[g] * 4;
        for (uint32_t j = len; j < ((len + 3) & ~3u); ++j) dst[j] = pad_base + (uint32_t)((g * 3 + j) & 0xFFFFu);
        longest = max(longest, len);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) longest = max(longest, __shfl_xor_sync(0xFFFFFFFFu, longest, o));
    if ((threadIdx.x & 31) == 0 && longest) atomicMax(&ctr->max_row_len, longest);
}

__global__ void narrow_kernel(const unsigned long long *in, uint32_t *out, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (uint32_t)in[i];
}

template <class T> struct Dev {
    T *p = nullptr;
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
    void free() {
        if (p) cudaFree(p);
        p = nullptr;
    }
    ~Dev() { free(); }
};

#define GB_CUDA(call)                                                   \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);           \
    } while (0)

} // namespace

struct GpuSnapshotBuilder::Segment {
    bool is_file = false;
    uint64_t commit_id = 0, merges = 0;
    uint32_t min_doc_id = 0, block_size = 0;
    uint64_t num_blocks = 0, n_items = 0, n_docs = 0;
    uint8_t *d_blocks = nullptr;
    uint32_t *d_block_index = nullptr;
    unsigned long long *d_items = nullptr;
    uint32_t *d_doc_ids = nullptr;
    ~Segment() {
        if (d_blocks) cudaFree(d_blocks);
        if (d_block_index) cudaFree(d_block_index);
        if (d_items) cudaFree(d_items);
        if (d_doc_ids) cudaFree(d_doc_ids);
    }
};

GpuSnapshotBuilder::GpuSnapshotBuilder() = default;
GpuSnapshotBuilder::~GpuSnapshotBuilder() {
    for (Segment *s : segs_) delete s;
}

bool GpuSnapshotBuilder::cuda_fail(cudaError_t e, const char *what) {
    error = std::string(what) + ": " + cudaGetErrorString(e);
    oom = e == cudaErrorMemoryAllocation;
    cudaGetLastError();
    return false;
}

bool GpuSnapshotBuilder::check_order(uint64_t commit_id, bool is_file) {
    if (!segs_.empty()) {
        const Segment &last = *segs_.back();
        if (is_file && !last.is_file) return fail("file segments must precede memory segments (Index.zig:33-41)");
        if (commit_id <= last.commit_id) return fail("segments must be added oldest to newest (ascending commit_id)");
    }
    return true;
}

bool GpuSnapshotBuilder::add_file_segment(uint64_t commit_id, uint64_t merges, uint32_t min_doc_id, uint32_t block_size,
                                          const uint8_t *blocks, uint64_t num_blocks, const uint32_t *block_index,
                                          const uint32_t *doc_ids, uint64_t n_docs) {
    if (poisoned_) return fail("an earlier segment failed to upload; abort this builder");
    if (!check_order(commit_id, true)) return false;
    if (block_size < kMinBlockSize || block_size > kMaxBlockSize) return fail("block_size out of range");
    // filefmt.zig:237 accepts any size in 64..4096; the device decoder reads blocks as 16-byte pieces
    if (block_size % 16) return fail_unsupported("block_size is not a multiple of 16");
    if (num_blocks && (!blocks || !block_index)) return fail("null blocks / block_index");
    if (num_blocks > 0xFFFFFFF0ull) return fail("too many blocks");
    std::unique_ptr<Segment> s(new Segment());
    s->is_file = true;
    s->commit_id = commit_id;
    s->merges = merges;
    s->min_doc_id = min_doc_id;
    s->block_size = block_size;
    s->num_blocks = num_blocks;
    s->n_docs = n_docs;
    poisoned_ = true; // until the uploads below have succeeded
    if (num_blocks) {
        GB_CUDA(cudaMalloc(&s->d_blocks, num_blocks * block_size));
        GB_CUDA(cudaMalloc(&s->d_block_index, num_blocks * 4));
        GB_CUDA(cudaMemcpy(s->d_blocks, blocks, num_blocks * block_size, cudaMemcpyHostToDevice));
        GB_CUDA(cudaMemcpy(s->d_block_index, block_index, num_blocks * 4, cudaMemcpyHostToDevice));
    }
    if (n_docs) {
        GB_CUDA(cudaMalloc(&s->d_doc_ids, n_docs * 4));
        GB_CUDA(cudaMemcpy(s->d_doc_ids, doc_ids, n_docs * 4, cudaMemcpyHostToDevice));
    }
    segs_.push_back(s.release());
    poisoned_ = false;
    return true;
}

bool GpuSnapshotBuilder::add_memory_segment(uint64_t commit_id, uint64_t merges, const uint64_t *items, uint64_t n_items,
                                            const uint32_t *doc_ids, uint64_t n_docs) {
    if (poisoned_) return fail("an earlier segment failed to upload; abort this builder");
    if (!check_order(commit_id, false)) return false;
    if (n_items && !items) return fail("null items");
    std::unique_ptr<Segment> s(new Segment());
    s->commit_id = commit_id;
    s->merges = merges;
    s->n_items = n_items;
    s->n_docs = n_docs;
    poisoned_ = true;
    if (n_items) {
        GB_CUDA(cudaMalloc(&s->d_items, n_items * 8));
        GB_CUDA(cudaMemcpy(s->d_items, items, n_items * 8, cudaMemcpyHostToDevice));
    }
    if (n_docs) {
        GB_CUDA(cudaMalloc(&s->d_doc_ids, n_docs * 4));
        GB_CUDA(cudaMemcpy(s->d_doc_ids, doc_ids, n_docs * 4, cudaMemcpyHostToDevice));
    }
    segs_.push_back(s.release());
    poisoned_ = false;
    return true;
}

// One segment's kept postings as a CSR: ascending terms, row lengths, row starts, docids (ascending inside a row).
struct GpuSnapshotBuilder::SegCsr {
    Dev<uint32_t> terms, lens, docids;
    Dev<unsigned long long> first; // nt entries: start of row j in docids
    uint64_t nt = 0, kept = 0;
};

// Decode, flag and select one segment; its raw bytes are freed on the way.
bool GpuSnapshotBuilder::build_segment(size_t si, bool multi, const void *newest_p, void *ctr_p, SegCsr &out) {
    Segment &s = *segs_[si];
    const NewestMap newest = *static_cast<const NewestMap *>(newest_p);
    BuildCounters *ctr = static_cast<BuildCounters *>(ctr_p);
    Dev<BlockMeta> meta;
    Dev<unsigned long long> blk_off;
    Dev<RunState> st;
    Dev<unsigned long long> keys;
    if (s.is_file) {
        const uint64_t nb = s.num_blocks;
        Dev<uint32_t> counts;
        GB_CUDA(counts.alloc(nb));
        GB_CUDA(blk_off.alloc(nb + 1));
        GB_CUDA(meta.alloc(nb));
        GB_CUDA(cudaMemset(meta.p, 0, std::max<uint64_t>(nb, 1) * sizeof(BlockMeta)));
        GB_CUDA(st.alloc(nb));
        if (nb) {
            block_counts_kernel<<<(unsigned)((nb + 255) / 256), 256>>>(s.d_blocks, s.block_size, nb, counts.p, ctr);
            size_t tmp_bytes = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts.p, blk_off.p, (long long)nb);
            Dev<uint8_t> tmp;
            GB_CUDA(tmp.alloc(tmp_bytes));
            GB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, counts.p, blk_off.p, (long long)nb));
            unsigned long long last_off = 0;
            uint32_t last_cnt = 0;
            GB_CUDA(cudaMemcpy(&last_off, blk_off.p + (nb - 1), 8, cudaMemcpyDeviceToHost));
            GB_CUDA(cudaMemcpy(&last_cnt, counts.p + (nb - 1), 4, cudaMemcpyDeviceToHost));
            s.n_items = last_off + last_cnt;
            GB_CUDA(cudaMemcpy(blk_off.p + nb, &s.n_items, 8, cudaMemcpyHostToDevice));
        } else {
            s.n_items = 0;
        }
        GB_CUDA(keys.alloc(s.n_items));
        if (nb) {
            // items per block: every quad costs two control-byte shares and >= 4 docid bytes (block.zig:479-485)
            const uint32_t max_items = std::min<uint32_t>(kWriterWindow, 4 * ((s.block_size - kBlockHeaderBytes) / 6 + 1));
            const size_t per_warp = ((size_t)s.block_size + 15) / 16 * 16 + (size_t)max_items * 8;
            const size_t smem = per_warp * kDecodeWarps;
            GB_CUDA(cudaFuncSetAttribute(decode_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const unsigned grid = (unsigned)std::min<uint64_t>((nb + kDecodeWarps - 1) / kDecodeWarps, 148ull * 32);
            decode_blocks_kernel<<<grid, kDecodeWarps * 32, smem>>>(s.d_blocks, s.block_size, nb, blk_off.p, s.d_block_index,
                                                                   s.min_doc_id, max_items, keys.p, meta.p, ctr);
            run_state_kernel<<<(unsigned)((nb + 255) / 256), 256>>>(meta.p, blk_off.p, nb, st.p, ctr);
            size_t tmp_bytes = 0;
            cub::DeviceScan::InclusiveScan(nullptr, tmp_bytes, st.p, st.p, RunStateOp(), (long long)nb);
            Dev<uint8_t> tmp;
            GB_CUDA(tmp.alloc(tmp_bytes));
            GB_CUDA(cub::DeviceScan::InclusiveScan(tmp.p, tmp_bytes, st.p, st.p, RunStateOp(), (long long)nb));
        }
    } else {
        keys.p = s.d_items; // (hash << 32) | id already
        s.d_items = nullptr;
        if (!keys.p) GB_CUDA(keys.alloc(1));
    }
    Dev<uint8_t> flags;
    GB_CUDA(flags.alloc(s.n_items));
    if (s.n_items)
        keep_flags_kernel<<<148 * 8, 256>>>(keys.p, s.n_items, s.is_file ? meta.p : nullptr, blk_off.p, st.p, s.num_blocks, newest,
                                            (uint32_t)si + 1, multi ? 1u : 0u, lo_, hi_, flags.p, ctr);
    GB_CUDA(cudaDeviceSynchronize());
    if (s.d_blocks) cudaFree(s.d_blocks); // the raw blocks are not needed any more
    s.d_blocks = nullptr;
    if (s.d_block_index) cudaFree(s.d_block_index);
    s.d_block_index = nullptr;
    meta.free();
    blk_off.free();
    st.free();
    if (s.n_items == 0) return true;

    // kept postings, still (hash << 32 | docid) ascending
    Dev<unsigned long long> kept, d_num;
    GB_CUDA(kept.alloc(s.n_items));
    GB_CUDA(d_num.alloc(1));
    {
        size_t tmp_bytes = 0;
        cub::DeviceSelect::Flagged(nullptr, tmp_bytes, keys.p, flags.p, kept.p, d_num.p, (long long)s.n_items);
        Dev<uint8_t> tmp;
        GB_CUDA(tmp.alloc(tmp_bytes));
        GB_CUDA(cub::DeviceSelect::Flagged(tmp.p, tmp_bytes, keys.p, flags.p, kept.p, d_num.p, (long long)s.n_items));
        unsigned long long got = 0;
        GB_CUDA(cudaMemcpy(&got, d_num.p, 8, cudaMemcpyDeviceToHost));
        out.kept = got;
    }
    keys.free();
    flags.free();
    if (out.kept == 0) return true;
    // rows: positions where the hash changes
    GB_CUDA(out.first.alloc(out.kept)); // at most one row per posting; trimmed below
    {
        auto pos = thrust::counting_iterator<unsigned long long>(0);
        auto heads = thrust::make_transform_iterator(pos, HeadOfRun{kept.p});
        size_t tmp_bytes = 0;
        cub::DeviceSelect::Flagged(nullptr, tmp_bytes, pos, heads, out.first.p, d_num.p, (long long)out.kept);
        Dev<uint8_t> tmp;
        GB_CUDA(tmp.alloc(tmp_bytes));
        GB_CUDA(cub::DeviceSelect::Flagged(tmp.p, tmp_bytes, pos, heads, out.first.p, d_num.p, (long long)out.kept));
        unsigned long long got = 0;
        GB_CUDA(cudaMemcpy(&got, d_num.p, 8, cudaMemcpyDeviceToHost));
        out.nt = got;
    }
    if (out.nt >= 0xFFFFFFFFull) return fail_unsupported("too many distinct terms in one segment");
    {   // trim the row starts to nt entries (the scratch above was sized for the worst case)
        Dev<unsigned long long> first;
        GB_CUDA(first.alloc(out.nt));
        GB_CUDA(cudaMemcpy(first.p, out.first.p, out.nt * 8, cudaMemcpyDeviceToDevice));
        out.first.free();
        out.first.p = first.p;
        first.p = nullptr;
    }
    GB_CUDA(out.terms.alloc(out.nt));
    GB_CUDA(out.lens.alloc(out.nt));
    GB_CUDA(out.docids.alloc(out.kept));
    segment_rows_kernel<<<(unsigned)((out.nt + 255) / 256), 256>>>(kept.p, out.kept, out.first.p, out.nt, out.terms.p, out.lens.p, ctr);
    narrow_kernel<<<148 * 16, 256>>>(kept.p, out.docids.p, out.kept);
    GB_CUDA(cudaDeviceSynchronize());
    return true;
}

bool GpuSnapshotBuilder::build(GpuCsr &out) {
    if (poisoned_) return fail("a segment failed to upload; abort this builder");
    const size_t ns = segs_.size();
    const bool multi = ns > 1;
    Dev<BuildCounters> ctr;
    GB_CUDA(ctr.alloc(1));
    GB_CUDA(cudaMemset(ctr.p, 0, sizeof(BuildCounters)));

    // ---- id -> newest mentioning segment (only needed when several segments feed the snapshot)
    NewestMap newest{nullptr, nullptr, 0};
    Dev<uint32_t> nk, nv;
    if (multi) {
        uint64_t total_docs = 0;
        for (Segment *s : segs_) total_docs += s->n_docs;
        uint64_t cap = 16;
        while (cap < 2 * total_docs + 2) cap <<= 1;
        if (cap > 0x80000000ull) return fail_unsupported("too many docs for the device liveness map");
        GB_CUDA(nk.alloc(cap));
        GB_CUDA(nv.alloc(cap));
        GB_CUDA(cudaMemset(nk.p, 0xFF, cap * 4));
        GB_CUDA(cudaMemset(nv.p, 0, cap * 4));
        newest = NewestMap{nk.p, nv.p, (uint32_t)(cap - 1)};
        for (size_t si = 0; si < ns; ++si)
            if (segs_[si]->n_docs)
                newest_insert_kernel<<<(unsigned)((segs_[si]->n_docs + 255) / 256), 256>>>(newest, segs_[si]->d_doc_ids,
                                                                                          segs_[si]->n_docs, (uint32_t)si + 1, ctr.p);
    }

    // ---- per segment: decode, flag, select -> the segment's own CSR (one segment's intermediates at a time, so the
    // peak is the raw blocks + one decoded segment + the CSRs built so far; no 2^31 limit on the whole snapshot)
    std::vector<std::unique_ptr<SegCsr>> csr(ns);
    uint64_t n_total = 0, n_kept = 0, n_entries = 0;
    for (size_t si = 0; si < ns; ++si) {
        csr[si].reset(new SegCsr());
        if (!build_segment(si, multi, &newest, ctr.p, *csr[si])) return false;
        n_total += segs_[si]->n_items;
        n_kept += csr[si]->kept;
        n_entries += csr[si]->nt;
    }
    nk.free();
    nv.free();
    BuildCounters hc{};
    GB_CUDA(cudaMemcpy(&hc, ctr.p, sizeof hc, cudaMemcpyDeviceToHost));
    switch (hc.error) {
    case 0: break;
    case 1: return fail("corrupt block (stream overruns the block)");
    case 2: return fail("block_index does not match the blocks");
    case 3: return fail("hashes not ascending");
    case 5: return fail_unsupported("docs map holds the id 0xFFFFFFFF (the device liveness map's empty marker)");
    case 6: return fail_unsupported("a row of more than 2^32 - 1 postings");
    default: return fail("memory segment items not sorted");
    }
    if (n_kept != n_total - hc.unreachable - hc.superseded - hc.out_of_range) return fail("internal: kept-posting count mismatch");
    if (n_entries >= 0xFFFFFFFFull) return fail_unsupported("too many (segment, term) rows for the device build");

    // ---- merge the segments' rows: sort all (term, entry) pairs by term (stable: a term's entries stay in segment order,
    // i.e. ascending docids for non-overlapping segments; the row order is re-done by reorder_rows_by_key anyway)
    const uint64_t N = n_entries;
    Dev<uint32_t> terms, lens, start4; // the merged row directory
    uint64_t n_rows = 0, total4 = 0;
    uint32_t max_live = 0;
    Dev<unsigned long long> dst_word; // per sorted entry: where its postings go in the padded array
    Dev<uint32_t> inv;                // entry -> sorted position
    std::vector<uint64_t> seg_base(ns + 1, 0);
    for (size_t si = 0; si < ns; ++si) seg_base[si + 1] = seg_base[si] + csr[si]->nt;
    if (N) {
        Dev<uint32_t> t_a, t_b, i_a, i_b, all_lens;
        GB_CUDA(t_a.alloc(N));
        GB_CUDA(t_b.alloc(N));
        GB_CUDA(i_a.alloc(N));
        GB_CUDA(i_b.alloc(N));
        GB_CUDA(all_lens.alloc(N));
        for (size_t si = 0; si < ns; ++si) {
            const uint64_t nt = csr[si]->nt;
            if (!nt) continue;
            GB_CUDA(cudaMemcpy(t_a.p + seg_base[si], csr[si]->terms.p, nt * 4, cudaMemcpyDeviceToDevice));
            GB_CUDA(cudaMemcpy(all_lens.p + seg_base[si], csr[si]->lens.p, nt * 4, cudaMemcpyDeviceToDevice));
            csr[si]->terms.free();
            csr[si]->lens.free();
            // largest live docid
            Dev<uint32_t> d_max;
            GB_CUDA(d_max.alloc(1));
            size_t tmp_bytes = 0;
            cub::DeviceReduce::Max(nullptr, tmp_bytes, csr[si]->docids.p, d_max.p, (long long)csr[si]->kept);
            Dev<uint8_t> tmp;
            GB_CUDA(tmp.alloc(tmp_bytes));
            GB_CUDA(cub::DeviceReduce::Max(tmp.p, tmp_bytes, csr[si]->docids.p, d_max.p, (long long)csr[si]->kept));
            uint32_t m = 0;
            GB_CUDA(cudaMemcpy(&m, d_max.p, 4, cudaMemcpyDeviceToHost));
            max_live = std::max(max_live, m);
        }
        iota_kernel<<<148 * 8, 256>>>(i_a.p, N);
        cub::DoubleBuffer<uint32_t> dk(t_a.p, t_b.p), dv(i_a.p, i_b.p);
        {
            size_t tmp_bytes = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (long long)N);
            Dev<uint8_t> tmp;
            GB_CUDA(tmp.alloc(tmp_bytes));
            GB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, dk, dv, (long long)N));
        }
        const uint32_t *st = dk.Current(), *sidx = dv.Current();
        // lengths in sorted order and their running sum
        Dev<unsigned long long> E;
        GB_CUDA(E.alloc(N + 1));
        {
            auto ls = thrust::make_transform_iterator(sidx, GatherLen{all_lens.p});
            size_t tmp_bytes = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ls, E.p, (long long)N);
            Dev<uint8_t> tmp;
            GB_CUDA(tmp.alloc(tmp_bytes));
            GB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, ls, E.p, (long long)N));
            GB_CUDA(cudaMemcpy(E.p + N, &n_kept, 8, cudaMemcpyHostToDevice));
        }
        // merged rows: positions where the sorted term changes
        Dev<uint32_t> head_pos;
        Dev<unsigned long long> d_num;
        GB_CUDA(head_pos.alloc(N + 1));
        GB_CUDA(d_num.alloc(1));
        {
            auto pos = thrust::counting_iterator<uint32_t>(0);
            auto heads = thrust::make_transform_iterator(pos, HeadOfTerm{st});
            size_t tmp_bytes = 0;
            cub::DeviceSelect::Flagged(nullptr, tmp_bytes, pos, heads, head_pos.p, d_num.p, (long long)N);
            Dev<uint8_t> tmp;
            GB_CUDA(tmp.alloc(tmp_bytes));
            GB_CUDA(cub::DeviceSelect::Flagged(tmp.p, tmp_bytes, pos, heads, head_pos.p, d_num.p, (long long)N));
            unsigned long long got = 0;
            GB_CUDA(cudaMemcpy(&got, d_num.p, 8, cudaMemcpyDeviceToHost));
            n_rows = got;
            const uint32_t n32 = (uint32_t)N;
            GB_CUDA(cudaMemcpy(head_pos.p + n_rows, &n32, 4, cudaMemcpyHostToDevice));
        }
        GB_CUDA(terms.alloc(n_rows));
        GB_CUDA(lens.alloc(n_rows));
        GB_CUDA(start4.alloc(n_rows + 1));
        merged_rows_kernel<<<(unsigned)((n_rows + 255) / 256), 256>>>(st, E.p, head_pos.p, n_rows, terms.p, lens.p, ctr.p);
        Dev<unsigned long long> start4_64;
        GB_CUDA(start4_64.alloc(n_rows + 1));
        {
            auto q4 = thrust::make_transform_iterator(static_cast<const uint32_t *>(lens.p), Quads());
            size_t tmp_bytes = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, q4, start4_64.p, (long long)n_rows);
            Dev<uint8_t> tmp;
            GB_CUDA(tmp.alloc(tmp_bytes));
            GB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, q4, start4_64.p, (long long)n_rows));
        }
        unsigned long long last4 = 0;
        uint32_t last_len = 0;
        GB_CUDA(cudaMemcpy(&last4, start4_64.p + (n_rows - 1), 8, cudaMemcpyDeviceToHost));
        GB_CUDA(cudaMemcpy(&last_len, lens.p + (n_rows - 1), 4, cudaMemcpyDeviceToHost));
        GB_CUDA(cudaMemcpy(&hc, ctr.p, sizeof hc, cudaMemcpyDeviceToHost));
        if (hc.error == 6) return fail_unsupported("a row of more than 2^32 - 1 postings");
        total4 = last4 + ((last_len + 3) >> 2);
        if (total4 > 0xFFFFFFFFull) return fail_unsupported("snapshot exceeds 2^32 16-byte granules");
        narrow_kernel<<<(unsigned)((n_rows + 255) / 256), 256>>>(start4_64.p, start4.p, n_rows);
        // where every (segment, term) entry goes, and the inverse of the sort
        GB_CUDA(dst_word.alloc(N));
        GB_CUDA(inv.alloc(N));
        entry_places_kernel<<<(unsigned)((n_rows + 255) / 256), 256>>>(sidx, E.p, head_pos.p, start4_64.p, n_rows, dst_word.p, inv.p);
        GB_CUDA(cudaDeviceSynchronize());
    }
    if (max_live >= 0xFFFE0000u) return fail_unsupported("docids reach the top of the u32 range (host build picks the padding)");

    // ---- padded rows
    out.pad_id = max_live + 1;
    out.pad_spread = true;
    out.n_terms = n_rows;
    out.total4 = total4;
    out.n_postings = n_kept;
    out.n_postings_total = n_total;
    out.n_unreachable = hc.unreachable;
    out.n_superseded = hc.superseded;
    out.n_out_of_range = hc.out_of_range;
    GB_CUDA(cudaMalloc(&out.d_docids, std::max<uint64_t>(total4 * 4, 4) * 4));
    if (n_rows) {
        for (size_t si = 0; si < ns; ++si) {
            if (!csr[si]->nt) continue;
            copy_entries_kernel<<<148 * 8, 256>>>(csr[si]->docids.p, csr[si]->first.p, csr[si]->kept, csr[si]->nt, inv.p + seg_base[si],
                                                  dst_word.p, out.d_docids);
            GB_CUDA(cudaDeviceSynchronize());
            csr[si].reset();
        }
        pad_rows_kernel<<<148 * 8, 256>>>(lens.p, start4.p, n_rows, max_live + 2, out.d_docids, ctr.p);
        GB_CUDA(cudaDeviceSynchronize());
        GB_CUDA(cudaMemcpy(&hc, ctr.p, sizeof hc, cudaMemcpyDeviceToHost));
    }
    out.max_row_len = hc.max_row_len;
    // hand the row directory arrays over (the caller builds the term hash table from them and frees them)
    out.d_terms = terms.p;
    out.d_row_len = lens.p;
    out.d_row_start4 = start4.p;
    terms.p = lens.p = start4.p = nullptr;
    try {
        out.h_terms.resize(n_rows);
        out.h_row_len.resize(n_rows);
        out.h_row_start4.resize(n_rows);
    } catch (const std::bad_alloc &) {
        oom = true;
        return fail("host copy of the term directory");
    }
    if (n_rows) {
        GB_CUDA(cudaMemcpy(out.h_terms.data(), out.d_terms, n_rows * 4, cudaMemcpyDeviceToHost));
        GB_CUDA(cudaMemcpy(out.h_row_len.data(), out.d_row_len, n_rows * 4, cudaMemcpyDeviceToHost));
        GB_CUDA(cudaMemcpy(out.h_row_start4.data(), out.d_row_start4, n_rows * 4, cudaMemcpyDeviceToHost));
    }
    return true;
}


// ------------------------------------------------------------------------------------------------
// rows in row_key order (fpx_kernels.cuh): docid -> key in place, segmented sort, key -> docid
// ------------------------------------------------------------------------------------------------
namespace {

template <bool INVERSE> __global__ void row_key_kernel(uint32_t *v, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        v[i] = INVERSE ? row_key_inv(v[i]) : row_key(v[i]);
}

struct RowBegin {
    const uint32_t *start4;
    unsigned long long base_word;
    __host__ __device__ int operator()(unsigned long long r) const { return (int)((unsigned long long)start4[r] * 4ull - base_word); }
};
struct RowEnd {
    const uint32_t *start4, *len;
    unsigned long long base_word;
    __host__ __device__ int operator()(unsigned long long r) const {
        return (int)((unsigned long long)start4[r] * 4ull - base_word + len[r]);
    }
};

} // namespace

cudaError_t reorder_rows_by_key(uint32_t **d_docids, uint64_t n_words, const uint32_t *d_row_len,
                                const uint32_t *d_row_start4, const uint32_t *h_row_start4, uint64_t n_rows) {
    if (n_rows == 0 || n_words == 0) return cudaSuccess;
    uint32_t *in = *d_docids, *out = nullptr;
    cudaError_t e = cudaMalloc(&out, n_words * sizeof(uint32_t));
    if (e != cudaSuccess) return e;
    row_key_kernel<false><<<148 * 8, 256>>>(in, n_words);
    // the sort writes the segments only: padding (and anything between rows) is carried over by a plain copy
    e = cudaMemcpy(out, in, n_words * sizeof(uint32_t), cudaMemcpyDeviceToDevice);
    void *tmp = nullptr;
    size_t tmp_cap = 0;
    constexpr uint64_t kMaxPiece = 1ull << 30; // words per cub call (its item count is an int)
    uint64_t r0 = 0;
    while (e == cudaSuccess && r0 < n_rows) {
        const uint64_t base_word = (uint64_t)h_row_start4[r0] * 4;
        // last row r1-1 such that the piece [base_word, start of row r1) stays below kMaxPiece; a single longer row
        // (more than 2^30 postings) cannot be sorted in one call
        uint64_t lo = r0 + 1, hi = n_rows;
        while (lo < hi) {
            const uint64_t mid = (lo + hi + 1) / 2;
            if ((uint64_t)h_row_start4[mid - 1] * 4 - base_word < kMaxPiece) lo = mid; else hi = mid - 1;
        }
        const uint64_t r1 = lo;
        const uint64_t end_word = r1 < n_rows ? (uint64_t)h_row_start4[r1] * 4 : n_words;
        if (end_word - base_word >= (1ull << 31)) {
            e = cudaErrorInvalidValue;
            break;
        }
        auto begins = thrust::make_transform_iterator(thrust::counting_iterator<unsigned long long>(r0),
                                                      RowBegin{d_row_start4, base_word});
        auto ends = thrust::make_transform_iterator(thrust::counting_iterator<unsigned long long>(r0),
                                                    RowEnd{d_row_start4, d_row_len, base_word});
        size_t need = 0;
        e = cub::DeviceSegmentedSort::SortKeys(nullptr, need, in + base_word, out + base_word, (int)(end_word - base_word),
                                               (int)(r1 - r0), begins, ends);
        if (e != cudaSuccess) break;
        if (need > tmp_cap) {
            if (tmp) cudaFree(tmp);
            tmp = nullptr;
            e = cudaMalloc(&tmp, need);
            if (e != cudaSuccess) break;
            tmp_cap = need;
        }
        e = cub::DeviceSegmentedSort::SortKeys(tmp, need, in + base_word, out + base_word, (int)(end_word - base_word),
                                               (int)(r1 - r0), begins, ends);
        r0 = r1;
    }
    if (e == cudaSuccess) {
        row_key_kernel<true><<<148 * 8, 256>>>(out, n_words);
        e = cudaDeviceSynchronize();
    }
    if (tmp) cudaFree(tmp);
    if (e != cudaSuccess) {
        cudaFree(out);
        return e;
    }
    cudaFree(in);
    *d_docids = out;
    return cudaSuccess;
}

} // namespace fpx
