//! zig/fpx.zig — the binding a fpindex maintainer would add as src/fpx.zig: `extern fn` declarations of
//! include/fpx.h and include/fpx_segment.h, plus the two call-site helpers of INTEGRATION.md.
//! SOURCE ONLY: there is no Zig toolchain in this repository's build image, so this file is not compiled or
//! tested here; the same C ABI is exercised through Python ctypes (acoustid-index_b200/_ffi.py) by every
//! `-m gpu` test, and tests/test_abi.py checks that libfpx.so exports every symbol named below.
//! Link: exe.linkSystemLibrary("fpx") in build.zig; libfpx.so needs only the CUDA driver at run time.

const std = @import("std");

pub const Status = c_int; // fpx_status: 0 OK, 1 OUT_OF_MEMORY, 2 INVALID_ARGUMENT, 3 INVALID_SEGMENT, 4 TIMEOUT,
//                           5 CUDA_ERROR, 6 BACKEND_UNAVAILABLE, 7 UNSUPPORTED
pub const Ctx = opaque {};
pub const Builder = opaque {};
pub const Snapshot = opaque {};
pub const Batcher = opaque {};
pub const SegmentFile = opaque {};

pub const Config = extern struct { device: i32 = -1, host_threads: u32 = 0, chunk_queries: u32 = 0, flags: u32 = 0 };

/// FileSegment.zig:33-53 as borrowed pointers (valid for the call only)
pub const FileSegmentDesc = extern struct {
    commit_id: u64,
    merges: u64,
    min_doc_id: u32,
    block_size: u32,
    blocks: ?[*]const u8,
    num_blocks: u64,
    block_index: ?[*]const u32,
    doc_ids: ?[*]const u32, // the docs map as two parallel arrays (inserts AND tombstones)
    doc_alive: ?[*]const u8,
    n_docs: u64,
};

/// MemorySegment.zig:21-28; Item is packed struct(u64){id: low 32, hash: high 32} (segment.zig:87-106): bit-compatible
pub const MemorySegmentDesc = extern struct {
    commit_id: u64,
    merges: u64,
    items: ?[*]const u64,
    n_items: u64,
    doc_ids: ?[*]const u32,
    doc_alive: ?[*]const u8,
    n_docs: u64,
};

/// common.zig:50-54, already resolved (MultiIndex.zig:302-306)
pub const SearchOpts = extern struct { max_results: u32, min_score: u32, min_score_pct: u32 };

pub const BatcherConfig = extern struct { max_batch: u32 = 0, max_wait_us: u32 = 100 };
pub const SegmentInfo = extern struct { commit_id: u64, merges: u64, version: u64, has_version: u32, reserved: u32 };
pub const SegmentBuf = opaque {};
pub const BatcherStats = extern struct { batches: u64, queries: u64, max_batch_seen: u64, timeouts: u64 };
pub const SnapshotInfo = extern struct {
    n_segments: u64,
    n_terms: u64,
    n_postings: u64,
    n_postings_total: u64,
    n_dropped_unreachable: u64,
    n_dropped_superseded: u64,
    n_dropped_out_of_range: u64,
    device_bytes: u64,
    max_row_len: u64,
    pad_id: u32,
    table_log2: u32,
    doc_lo: u32,
    doc_hi: u32,
};
pub const CsrView = extern struct { n_terms: u64, terms: ?[*]const u32, row_offsets: ?[*]const u64, docids: ?[*]const u32 };
pub const WireSearchRequest = extern struct { query: ?[*]u32, n_terms: u64, timeout: u32, limit: u32, has_min_score: u32, min_score: u32, score_pct: u32 };
pub const Profile = extern struct {
    prepare_ms: f64,
    prepare_launches: u64,
    sketch_ms: f64,
    sketch_launches: u64,
    search_ms: f64,
    search_launches: u64,
    wide_ms: f64,
    wide_launches: u64,
    h2d_ms: f64,
    d2h_ms: f64,
    queries: u64,
    unique_terms: u64,
    postings: u64,
    results: u64,
    sketch_queries: u64,
    wide_queries: u64,
    overflow_requeues: u64,
    h2d_bytes: u64,
    d2h_bytes: u64,
};

pub extern fn fpx_abi_version() u32;
pub extern fn fpx_last_error_message() [*:0]const u8;
pub extern fn fpx_init(cfg: ?*const Config, out: *?*Ctx) Status;
pub extern fn fpx_shutdown(ctx: *Ctx) void;

// ---- snapshot build: Index.swapSnapshot (Index.zig:469-485)
pub extern fn fpx_snapshot_begin(ctx: *Ctx, out: *?*Builder) Status;
pub extern fn fpx_snapshot_add_file_segment(b: *Builder, seg: *const FileSegmentDesc) Status;
pub extern fn fpx_snapshot_add_memory_segment(b: *Builder, seg: *const MemorySegmentDesc) Status;
pub extern fn fpx_snapshot_set_doc_range(b: *Builder, lo: u32, hi: u32) Status;
pub extern fn fpx_snapshot_commit(b: *Builder, out: *?*Snapshot) Status;
pub extern fn fpx_snapshot_abort(b: *Builder) void;
pub extern fn fpx_snapshot_acquire(s: *Snapshot) Status;
pub extern fn fpx_snapshot_release(s: *Snapshot) Status;

// ---- search: IndexReader.search + SearchResults.finish (Index.zig:170-177, common.zig:131-167)
pub extern fn fpx_default_min_score(raw_query_len: u64) u32;
pub extern fn fpx_search(s: *Snapshot, terms: [*]const u32, n_terms: u64, opts: *const SearchOpts, out_ids: [*]u32, out_scores: [*]u32, capacity: u32, out_count: *u32) Status;
pub extern fn fpx_search_batch(s: *Snapshot, n_queries: u64, terms: [*]const u32, term_offsets: [*]const u64, opts: [*]const SearchOpts, k_stride: u32, out_ids: [*]u32, out_scores: [*]u32, out_counts: [*]u32) Status;
pub extern fn fpx_search_batch_timeout(s: *Snapshot, n_queries: u64, terms: [*]const u32, term_offsets: [*]const u64, opts: [*]const SearchOpts, k_stride: u32, out_ids: [*]u32, out_scores: [*]u32, out_counts: [*]u32, timeout_ms: u32) Status;
/// results as the reference returns them: a list per query (counts + (id, score) pairs back to back)
pub extern fn fpx_search_batch_packed(s: *Snapshot, n_queries: u64, terms: [*]const u32, term_offsets: [*]const u64, opts: [*]const SearchOpts, k_stride: u32, out_counts: [*]u32, out_pairs: [*]u32, capacity_pairs: u64, out_n_pairs: *u64) Status;
// device-resident variants: every pointer is a CUDA device pointer, `cuda_stream` a cudaStream_t (null = the legacy default
// stream); for hosts that own GPU buffers.  `_async` never waits for the device; what the kernels reject comes back
// through the status word in device memory.
pub extern fn fpx_search_batch_device_async(s: *Snapshot, n_queries: u64, term_base: u64, n_terms_total: u64, d_terms: ?*const anyopaque, d_term_offsets: ?*const anyopaque, d_opts: ?*const anyopaque, k_stride: u32, d_out_ids: ?*anyopaque, d_out_scores: ?*anyopaque, d_out_counts: ?*anyopaque, d_status: ?*anyopaque, cuda_stream: ?*anyopaque) Status;
pub extern fn fpx_search_batch_device(s: *Snapshot, n_queries: u64, d_terms: ?*const anyopaque, d_term_offsets: ?*const anyopaque, d_opts: ?*const anyopaque, k_stride: u32, d_out_ids: ?*anyopaque, d_out_scores: ?*anyopaque, d_out_counts: ?*anyopaque, cuda_stream: ?*anyopaque) Status;
// docid-range sharded corpus (one shard per GPU): pack a shard's lists, merge the gathered lists (device and host forms)
pub extern fn fpx_pack_results_device(n_queries: u64, k_stride: u32, d_ids: ?*const anyopaque, d_scores: ?*const anyopaque, d_counts: ?*const anyopaque, d_packed: ?*anyopaque, capacity_pairs: u32, cuda_stream: ?*anyopaque) Status;
pub extern fn fpx_merge_packed_shards_device(n_shards: u32, n_queries: u64, d_packed: ?*const anyopaque, shard_stride_words: u64, d_opts: ?*const anyopaque, k_stride: u32, d_out_ids: ?*anyopaque, d_out_scores: ?*anyopaque, d_out_counts: ?*anyopaque, cuda_stream: ?*anyopaque) Status;
pub extern fn fpx_merge_shard_results(n_shards: u32, n_queries: u64, k_stride: u32, ids: [*]const u32, scores: [*]const u32, counts: [*]const u32, opts: [*]const SearchOpts, out_ids: [*]u32, out_scores: [*]u32, out_counts: [*]u32) Status;

// ---- the single-query seam of MultiIndex.search (MultiIndex.zig:287-330)
pub extern fn fpx_batcher_create(ctx: *Ctx, cfg: ?*const BatcherConfig, out: *?*Batcher) Status;
pub extern fn fpx_batcher_set_snapshot(b: *Batcher, s: ?*Snapshot) Status;
pub extern fn fpx_batcher_search(b: *Batcher, terms: [*]const u32, n_terms: u64, opts: *const SearchOpts, timeout_ms: u32, out_ids: [*]u32, out_scores: [*]u32, capacity: u32, out_count: *u32) Status;
pub extern fn fpx_batcher_get_stats(b: *Batcher, out: *BatcherStats) Status;
pub extern fn fpx_batcher_destroy(b: *Batcher) void;

// ---- segment files for a search-only process (filefmt.zig:209-285, manifest.zig:17-39)
pub extern fn fpx_segment_file_read(path: [*:0]const u8, out: *?*SegmentFile) Status;
pub extern fn fpx_segment_file_view(f: *SegmentFile, seg: *FileSegmentDesc, info: *SegmentInfo) Status;
pub extern fn fpx_segment_file_close(f: *SegmentFile) void;
pub extern fn fpx_manifest_parse(data: [*]const u8, size: u64, out: [*]SegmentInfo, capacity: u64, out_n: *u64) Status;
pub extern fn fpx_segment_file_parse(data: [*]const u8, size: u64, out: *?*SegmentFile) Status;
pub extern fn fpx_segment_file_num_items(f: *const SegmentFile) u64;
pub extern fn fpx_segment_file_metadata_count(f: *const SegmentFile) u64;
pub extern fn fpx_segment_file_metadata_get(f: *const SegmentFile, i: u64, key: *[*]const u8, key_len: *u64, value: *[*]const u8, value_len: *u64) Status;
pub extern fn fpx_segment_file_serialize(seg: *const FileSegmentDesc, info: *const SegmentInfo, out: *?[*]u8, out_size: *u64) Status;
pub extern fn fpx_bytes_free(p: ?[*]u8) void;
pub extern fn fpx_segment_file_name(commit_id: u64, merges: u64, buf: [*]u8, cap: u64) i32;
pub extern fn fpx_crc64_xz(data: [*]const u8, size: u64) u64;

// ---- the block codec by itself (block.zig:438-567 / filefmt.zig:94-138): writer and single-block reader
pub extern fn fpx_segment_write(items: [*]const u64, n_items: u64, min_doc_id: u32, block_size: u32, threads: u32, out: *?*SegmentBuf) Status;
pub extern fn fpx_segment_buf_blocks(b: *const SegmentBuf) [*]const u8;
pub extern fn fpx_segment_buf_block_index(b: *const SegmentBuf) [*]const u32;
pub extern fn fpx_segment_buf_num_blocks(b: *const SegmentBuf) u64;
pub extern fn fpx_segment_buf_num_items(b: *const SegmentBuf) u64;
pub extern fn fpx_segment_buf_block_size(b: *const SegmentBuf) u32;
pub extern fn fpx_segment_buf_free(b: *SegmentBuf) void;
pub extern fn fpx_block_decode(block: [*]const u8, block_size: u32, min_doc_id: u32, out_hashes: [*]u32, out_docids: [*]u32) i32;

// ---- snapshot introspection (tests, accounting) and the host-side CSR view of a builder
pub extern fn fpx_snapshot_compile(b: *Builder) Status;
pub extern fn fpx_snapshot_csr(b: *Builder, out: *CsrView) Status;
pub extern fn fpx_snapshot_get_info(s: *const Snapshot, out: *SnapshotInfo) Status;
pub extern fn fpx_snapshot_read_row(s: *const Snapshot, term: u32, out_docids: [*]u32, capacity: u64, out_len: *u64) Status;
pub extern fn fpx_snapshot_row_lengths(s: *const Snapshot, terms: [*]const u32, n: u64, out_lengths: [*]u32) Status;

// ---- wire codecs of the search endpoint (api.zig:14-72, server.zig:84-142, legacy.zig:185-210, 286-296)
pub extern fn fpx_wire_decode_search_request(format: u32, data: [*]const u8, size: u64, out: *WireSearchRequest) Status;
pub extern fn fpx_wire_encode_search_response(format: u32, ids: [*]const u32, scores: [*]const u32, n: u32, out: *?[*]u8, out_size: *u64) Status;
pub extern fn fpx_legacy_parse_fingerprint(text: [*]const u8, len: u64, out_terms: *?[*]u32, out_n: *u64) Status;
pub extern fn fpx_legacy_format_results(ids: [*]const u32, scores: [*]const u32, n: u32, out: *?[*]u8, out_size: *u64) Status;
pub extern fn fpx_wire_free(p: ?*anyopaque) void;

// ---- tuning and profiling
pub extern fn fpx_set_chunk_queries(ctx: *Ctx, chunk_queries: u32) Status;
pub extern fn fpx_set_profile(ctx: *Ctx, enabled: c_int) Status;
pub extern fn fpx_profile_reset(ctx: *Ctx) Status;
pub extern fn fpx_profile_read(ctx: *Ctx, out: *Profile) Status;
pub extern fn fpx_debug_set(ctx: *Ctx, bits: u32) Status;

pub fn toError(st: Status) anyerror!void {
    return switch (st) {
        0 => {},
        1 => error.OutOfMemory,
        3 => error.InvalidSegment, // filefmt.zig:235-284
        4 => error.SearchTimeout, // MultiIndex.zig:320
        6, 7 => error.GpuPathUnavailable, // the caller keeps the CPU path
        else => error.GpuError,
    };
}

/// Call site 1, Index.swapSnapshot: mirror a new Segments snapshot in HBM.  `Segments`, `collectDocs` are the
/// reference's own types / a small helper that walks a docs map into two arrays (see INTEGRATION.md section 2).
pub fn buildGpuMirror(ctx: *Ctx, segs: anytype, allocator: std.mem.Allocator) ?*Snapshot {
    var b: ?*Builder = null;
    if (fpx_snapshot_begin(ctx, &b) != 0) return null;
    for (segs.file) |ref| { // oldest -> newest, file segments before memory segments (Index.zig:33-41)
        const s = ref.value;
        const docs = collectDocs(allocator, s.docs) catch {
            fpx_snapshot_abort(b.?);
            return null;
        };
        defer docs.deinit();
        const d = FileSegmentDesc{
            .commit_id = s.info.commit_id,
            .merges = s.info.merges,
            .min_doc_id = s.min_doc_id,
            .block_size = @intCast(s.block_size),
            .blocks = s.blocks.ptr,
            .num_blocks = s.num_blocks,
            .block_index = s.block_index.ptr,
            .doc_ids = docs.ids.ptr,
            .doc_alive = docs.alive.ptr,
            .n_docs = docs.ids.len,
        };
        if (fpx_snapshot_add_file_segment(b.?, &d) != 0) {
            fpx_snapshot_abort(b.?);
            return null;
        }
    }
    for (segs.memory) |ref| {
        const s = ref.value;
        const docs = collectDocs(allocator, s.docs) catch {
            fpx_snapshot_abort(b.?);
            return null;
        };
        defer docs.deinit();
        const d = MemorySegmentDesc{
            .commit_id = s.info.commit_id,
            .merges = s.info.merges,
            .items = @ptrCast(s.items.items.ptr),
            .n_items = s.items.items.len,
            .doc_ids = docs.ids.ptr,
            .doc_alive = docs.alive.ptr,
            .n_docs = docs.ids.len,
        };
        if (fpx_snapshot_add_memory_segment(b.?, &d) != 0) {
            fpx_snapshot_abort(b.?);
            return null;
        }
    }
    var snap: ?*Snapshot = null;
    if (fpx_snapshot_commit(b.?, &snap) != 0) return null; // the CPU path stays in charge
    return snap;
}

const Docs = struct {
    ids: []u32,
    alive: []u8,
    allocator: std.mem.Allocator,
    fn deinit(self: Docs) void {
        self.allocator.free(self.ids);
        self.allocator.free(self.alive);
    }
};

fn collectDocs(allocator: std.mem.Allocator, docs: anytype) !Docs {
    const n = docs.count();
    const ids = try allocator.alloc(u32, n);
    errdefer allocator.free(ids);
    const alive = try allocator.alloc(u8, n);
    var it = docs.iterator();
    var i: usize = 0;
    while (it.next()) |e| : (i += 1) {
        ids[i] = e.key_ptr.*;
        alive[i] = @intFromBool(e.value_ptr.*);
    }
    return .{ .ids = ids, .alive = alive, .allocator = allocator };
}

/// Call site 2, IndexReader.search: returns false when the query lies outside the device path's limits (the caller
/// then runs the CPU body unchanged).  `results` is the reference's SearchResults (common.zig:91-119).
pub fn searchOnGpu(batcher: *Batcher, hashes: []const u32, results: anytype, timeout_ms: u32) !bool {
    if (hashes.len > 8192 or results.options.max_results > 1024) return false; // FPX_MAX_QUERY_TERMS / FPX_MAX_RESULTS
    const opts = SearchOpts{
        .max_results = results.options.max_results,
        .min_score = results.options.min_score, // already resolved by MultiIndex.zig:302-306
        .min_score_pct = results.options.min_score_pct,
    };
    var ids: [1024]u32 = undefined;
    var scores: [1024]u32 = undefined;
    var n: u32 = 0;
    const st = fpx_batcher_search(batcher, hashes.ptr, hashes.len, &opts, timeout_ms, &ids, &scores, opts.max_results, &n);
    if (st == 6 or st == 7) return false;
    try toError(st);
    try results.results.ensureTotalCapacity(results.allocator, n);
    results.results.clearRetainingCapacity();
    for (0..n) |i| results.results.appendAssumeCapacity(.{ .id = ids[i], .score = scores[i] });
    return true; // finish() is already applied (common.zig:131-167)
}
