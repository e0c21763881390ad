// ubench_gather.cu — how fast can one B200 gather short posting rows (mean ~290 B, 16-byte aligned, random
// places in a multi-GB array) into the SMs?  Three mechanisms, no counting work, so this is the ceiling of the
// gather stage of the search kernel.
//   A  cp.async.bulk (TMA, UBLKCP) row -> shared-memory stage, mbarrier pipeline, 1 producer warp per CTA
//   B  LDG.128 straight to registers, a warp per row, two rows in flight per warp
//   C  cp.async 16 B (LDGSTS) by all warps into a double-buffered stage
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_gather ubench_gather.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int ROWS_PER_Q = 100;
constexpr int STAGE_U4 = 2304; // 36 KB stage: one query's rows

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
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- A: TMA bulk copies
template <int STAGES, int CONSUMER_WARPS>
__global__ void __launch_bounds__((CONSUMER_WARPS + 1) * 32) gather_tma(const uint4 *data, const uint2 *rows, int n_queries,
                                                                         unsigned long long *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint4 *stage = reinterpret_cast<uint4 *>(smem);
    __shared__ uint64_t full[STAGES], empty[STAGES];
    __shared__ uint32_t s_total[STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == CONSUMER_WARPS) { // producer
        int it = 0;
        for (int q = blockIdx.x; q < n_queries; q += gridDim.x, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            if (it >= STAGES) mbar_wait(&empty[s], ph ^ 1);
            // row descriptors: 100 rows -> lanes take 4 each (rows lane, lane+32, ...)
            uint2 r[4];
            uint32_t n4[4], sum = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = lane + 32 * j;
                r[j] = idx < ROWS_PER_Q ? rows[(size_t)q * ROWS_PER_Q + idx] : make_uint2(0, 0);
                n4[j] = (r[j].y + 3) >> 2;
                sum += n4[j];
            }
            // exclusive offsets: order rows as (j, lane)
            uint32_t off[4], base = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t x = n4[j];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
                    if (lane >= d) x += y;
                }
                off[j] = base + x - n4[j];
                base += __shfl_sync(0xFFFFFFFFu, x, 31);
            }
            if (lane == 0) {
                s_total[s] = base;
                mbar_expect_tx(&full[s], base * 16);
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n4[j]) bulk_g2s(stage + (size_t)s * STAGE_U4 + off[j], data + r[j].x, n4[j] * 16, &full[s]);
        }
    } else { // consumers
        unsigned long long acc = 0;
        int it = 0;
        for (int q = blockIdx.x; q < n_queries; q += gridDim.x, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&full[s], ph);
            const uint32_t total = s_total[s];
            for (uint32_t i = threadIdx.x; i < total; i += CONSUMER_WARPS * 32) {
                const uint4 v = stage[(size_t)s * STAGE_U4 + i];
                acc += v.x ^ v.y ^ v.z ^ v.w;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (acc == 0x123456789ull) sink[0] = acc;
    }
}


// ---------------------------------------------------------------- A2: TMA, several producer warps share each query
template <int STAGES, int CONSUMER_WARPS, int PRODUCER_WARPS>
__global__ void __launch_bounds__((CONSUMER_WARPS + PRODUCER_WARPS) * 32)
gather_tma_split(const uint4 *data, const uint2 *rows, int n_queries, unsigned long long *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint4 *stage = reinterpret_cast<uint4 *>(smem);
    __shared__ uint64_t full[STAGES], empty[STAGES];
    __shared__ uint32_t s_total[STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], PRODUCER_WARPS);
            mbar_init(&empty[s], CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp >= CONSUMER_WARPS) { // producers: every warp scans all row descriptors, issues every PW-th chunk
        const int p = warp - CONSUMER_WARPS;
        int it = 0;
        for (int q = blockIdx.x; q < n_queries; q += gridDim.x, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            if (it >= STAGES) mbar_wait(&empty[s], ph ^ 1);
            uint32_t base = 0, mine = 0;
            uint2 r[4];
            uint32_t n4[4], off[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = lane + 32 * j;
                r[j] = idx < ROWS_PER_Q ? rows[(size_t)q * ROWS_PER_Q + idx] : make_uint2(0, 0);
                n4[j] = (r[j].y + 3) >> 2;
                uint32_t x = n4[j];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
                    if (lane >= d) x += y;
                }
                off[j] = base + x - n4[j];
                const uint32_t chunk_total = __shfl_sync(0xFFFFFFFFu, x, 31);
                if (j % PRODUCER_WARPS == p) mine += chunk_total;
                base += chunk_total;
            }
            if (lane == 0) {
                if (p == 0) s_total[s] = base;
                mbar_expect_tx(&full[s], mine * 16);
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j % PRODUCER_WARPS == p && n4[j])
                    bulk_g2s(stage + (size_t)s * STAGE_U4 + off[j], data + r[j].x, n4[j] * 16, &full[s]);
        }
    } else {
        unsigned long long acc = 0;
        int it = 0;
        for (int q = blockIdx.x; q < n_queries; q += gridDim.x, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&full[s], ph);
            const uint32_t total = s_total[s];
            for (uint32_t i = threadIdx.x; i < total; i += CONSUMER_WARPS * 32) {
                const uint4 v = stage[(size_t)s * STAGE_U4 + i];
                acc += v.x ^ v.y ^ v.z ^ v.w;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (acc == 0x123456789ull) sink[0] = acc;
    }
}

// ---------------------------------------------------------------- B: LDG.128 to registers
template <int UNROLL>
__global__ void __launch_bounds__(256) gather_ldg(const uint4 *data, const uint2 *rows, int n_queries, unsigned long long *sink) {
    __shared__ uint2 rows_s[ROWS_PER_Q];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long acc = 0;
    for (int q = blockIdx.x; q < n_queries; q += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < ROWS_PER_Q) rows_s[threadIdx.x] = rows[(size_t)q * ROWS_PER_Q + threadIdx.x];
        __syncthreads();
        for (int r = warp * UNROLL; r < ROWS_PER_Q; r += 8 * UNROLL) {
            uint4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                v[u] = make_uint4(0, 0, 0, 0);
                if (r + u < ROWS_PER_Q) {
                    const uint2 d = rows_s[r + u];
                    const uint32_t n4 = (d.y + 3) >> 2;
                    for (uint32_t i = lane; i < n4; i += 32) {
                        const uint4 t = __ldg(data + d.x + i);
                        v[u].x ^= t.x; v[u].y ^= t.y; v[u].z ^= t.z; v[u].w ^= t.w;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
        }
    }
    if (acc == 0x123456789ull) sink[0] = acc;
}

// ---------------------------------------------------------------- C: cp.async (LDGSTS) double buffer
__global__ void __launch_bounds__(256) gather_cpasync(const uint4 *data, const uint2 *rows, int n_queries, unsigned long long *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint4 *stage = reinterpret_cast<uint4 *>(smem);
    __shared__ uint2 rows_s[2][ROWS_PER_Q];
    __shared__ uint32_t offs[2][ROWS_PER_Q + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long acc = 0;
    auto issue = [&](int q, int s) {
        if (threadIdx.x < ROWS_PER_Q) rows_s[s][threadIdx.x] = rows[(size_t)q * ROWS_PER_Q + threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t o = 0;
            for (int r = 0; r < ROWS_PER_Q; ++r) {
                offs[s][r] = o;
                o += (rows_s[s][r].y + 3) >> 2;
            }
            offs[s][ROWS_PER_Q] = o;
        }
        __syncthreads();
        for (int r = warp; r < ROWS_PER_Q; r += 8) {
            const uint2 d = rows_s[s][r];
            const uint32_t n4 = (d.y + 3) >> 2, o = offs[s][r];
            for (uint32_t i = lane; i < n4; i += 32) {
                const uint32_t dst = smem_u32(stage + (size_t)s * STAGE_U4 + o + i);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(data + d.x + i) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int it = 0;
    int q = blockIdx.x;
    if (q < n_queries) issue(q, 0);
    for (; q < n_queries; q += gridDim.x, ++it) {
        const int s = it & 1;
        const int qn = q + gridDim.x;
        if (qn < n_queries) {
            issue(qn, s ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const uint32_t total = offs[s][ROWS_PER_Q];
        for (uint32_t i = threadIdx.x; i < total; i += 256) {
            const uint4 v = stage[(size_t)s * STAGE_U4 + i];
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
        __syncthreads();
    }
    if (acc == 0x123456789ull) sink[0] = acc;
}

static uint64_t sm64(uint64_t &x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const size_t data_u4 = (size_t)4 << 30 >> 4; // 4 GiB of postings
    const int nq = 200000;
    uint4 *data;
    uint2 *rows;
    unsigned long long *sink;
    cudaMalloc(&data, data_u4 * 16);
    cudaMemset(data, 1, data_u4 * 16);
    cudaMalloc(&sink, 8);
    std::vector<uint2> h((size_t)nq * ROWS_PER_Q);
    uint64_t st = 42, bytes = 0;
    for (auto &r : h) {
        uint32_t len = 40 + (uint32_t)(sm64(st) % 64); // 40..103 docids, mean 71.5
        r.y = len;
        r.x = (uint32_t)(sm64(st) % (data_u4 - 64));
        bytes += (uint64_t)((len + 3) / 4) * 16;
    }
    cudaMalloc(&rows, h.size() * sizeof(uint2));
    cudaMemcpy(rows, h.data(), h.size() * sizeof(uint2), cudaMemcpyHostToDevice);
    printf("%s, %d SMs; %d queries x %d rows, %.2f GB gathered per launch (padded rows)\n", p.name, sms, nq, ROWS_PER_Q, bytes / 1e9);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    auto report = [&](const char *name, float ms) {
        printf("%-52s %8.3f ms  %8.1f GB/s  %6.1f Mq/s  err=%s\n", name, ms, bytes / ms / 1e6, nq / ms / 1e3,
               cudaGetErrorString(cudaGetLastError()));
    };
#define TIME(name, launch)                \
    do {                                  \
        launch;                           \
        cudaDeviceSynchronize();          \
        cudaEventRecord(a);               \
        launch;                           \
        cudaEventRecord(b);               \
        cudaEventSynchronize(b);          \
        float ms;                         \
        cudaEventElapsedTime(&ms, a, b);  \
        report(name, ms);                 \
    } while (0)

    {
        auto k = gather_tma<2, 4>;
        int sm = 2 * STAGE_U4 * 16;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        TIME("A tma 2 stages, 4 cons warps, 1 CTA/SM", (k<<<sms, 160, sm>>>(data, rows, nq, sink)));
        TIME("A tma 2 stages, 4 cons warps, 2 CTA/SM", (k<<<sms * 2, 160, sm>>>(data, rows, nq, sink)));
        TIME("A tma 2 stages, 4 cons warps, 3 CTA/SM", (k<<<sms * 3, 160, sm>>>(data, rows, nq, sink)));
    }
    {
        auto k = gather_tma<3, 8>;
        int sm = 3 * STAGE_U4 * 16;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        TIME("A tma 3 stages, 8 cons warps, 1 CTA/SM", (k<<<sms, 288, sm>>>(data, rows, nq, sink)));
        TIME("A tma 3 stages, 8 cons warps, 2 CTA/SM", (k<<<sms * 2, 288, sm>>>(data, rows, nq, sink)));
    }
    {
        auto k = gather_tma<6, 8>;
        int sm = 6 * STAGE_U4 * 16;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        TIME("A tma 6 stages, 8 cons warps, 1 CTA/SM", (k<<<sms, 288, sm>>>(data, rows, nq, sink)));
    }

    {
        auto k = gather_tma_split<2, 8, 2>;
        int sm = 2 * STAGE_U4 * 16;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        TIME("A2 tma split 2 stages, 8 cons, 2 prod, 1 CTA/SM", (k<<<sms, 320, sm>>>(data, rows, nq, sink)));
        TIME("A2 tma split 2 stages, 8 cons, 2 prod, 2 CTA/SM", (k<<<sms * 2, 320, sm>>>(data, rows, nq, sink)));
    }
    {
        auto k = gather_tma_split<2, 8, 4>;
        int sm = 2 * STAGE_U4 * 16;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        TIME("A2 tma split 2 stages, 8 cons, 4 prod, 1 CTA/SM", (k<<<sms, 384, sm>>>(data, rows, nq, sink)));
        TIME("A2 tma split 2 stages, 8 cons, 4 prod, 2 CTA/SM", (k<<<sms * 2, 384, sm>>>(data, rows, nq, sink)));
        TIME("A2 tma split 2 stages, 8 cons, 4 prod, 3 CTA/SM", (k<<<sms * 3, 384, sm>>>(data, rows, nq, sink)));
    }
    {
        auto k = gather_tma_split<4, 8, 4>;
        int sm = 4 * STAGE_U4 * 16;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        TIME("A2 tma split 4 stages, 8 cons, 4 prod, 1 CTA/SM", (k<<<sms, 384, sm>>>(data, rows, nq, sink)));
    }
    {
        auto k = gather_tma_split<6, 4, 4>;
        int sm = 6 * STAGE_U4 * 16;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        TIME("A2 tma split 6 stages, 4 cons, 4 prod, 1 CTA/SM", (k<<<sms, 256, sm>>>(data, rows, nq, sink)));
    }
    TIME("B ldg128 warp/row unroll 1, 3 CTA/SM", (gather_ldg<1><<<sms * 3, 256>>>(data, rows, nq, sink)));
    TIME("B ldg128 warp/row unroll 2, 3 CTA/SM", (gather_ldg<2><<<sms * 3, 256>>>(data, rows, nq, sink)));
    TIME("B ldg128 warp/row unroll 4, 3 CTA/SM", (gather_ldg<4><<<sms * 3, 256>>>(data, rows, nq, sink)));
    TIME("B ldg128 warp/row unroll 4, 6 CTA/SM", (gather_ldg<4><<<sms * 6, 256>>>(data, rows, nq, sink)));
    TIME("B ldg128 warp/row unroll 4, 8 CTA/SM", (gather_ldg<4><<<sms * 8, 256>>>(data, rows, nq, sink)));
    {
        int sm = 2 * STAGE_U4 * 16;
        cudaFuncSetAttribute(gather_cpasync, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        TIME("C cp.async double buffer, 1 CTA/SM", (gather_cpasync<<<sms, 256, sm>>>(data, rows, nq, sink)));
        TIME("C cp.async double buffer, 2 CTA/SM", (gather_cpasync<<<sms * 2, 256, sm>>>(data, rows, nq, sink)));
        TIME("C cp.async double buffer, 3 CTA/SM", (gather_cpasync<<<sms * 3, 256, sm>>>(data, rows, nq, sink)));
    }
    return 0;
}
