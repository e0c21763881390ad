// ubench_smem.cu — shared-memory atomic / random-access throughput on B200, the numbers that size the
// per-query counting table.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_smem ubench_smem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int LOG = 14, SLOTS = 1 << LOG, THREADS = 256;

template <int MODE>
__global__ void __launch_bounds__(THREADS, 3) k(uint32_t *out, int iters, uint32_t seed) {
    extern __shared__ uint32_t tab[];
    for (int i = threadIdx.x; i < SLOTS; i += THREADS) tab[i] = 0;
    __syncthreads();
    uint32_t x = (blockIdx.x * THREADS + threadIdx.x) * 2654435761u + seed;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        uint32_t s = (x * 0x9E3779B1u) >> (32 - LOG);
        if (MODE == 0) acc += atomicCAS(&tab[s], 0u, x | 1u);            // CAS, result used
        if (MODE == 1) atomicAdd(&tab[s], 1u);                           // add, result unused
        if (MODE == 2) acc += atomicAdd(&tab[s], 1u);                    // add, result used
        if (MODE == 3) acc += tab[s];                                    // random LDS
        if (MODE == 4) tab[s] = x;                                       // random STS
        if (MODE == 5) { uint32_t v = tab[s]; tab[s] = v + 1; acc += v; } // non-atomic RMW (racy; cost only)
        if (MODE == 6) atomicOr(&tab[s >> 5], 1u << (s & 31));           // bitmap set, result unused
        if (MODE == 7) { // sorted-lane pattern: consecutive lanes -> consecutive slots (conflict-free atomics)
            uint32_t s2 = ((x >> 8) * 32 + (threadIdx.x & 31)) & (SLOTS - 1);
            atomicAdd(&tab[s2], 1u);
        }
    }
    __syncthreads();
    if (acc == 0x12345678u) out[0] = acc + tab[threadIdx.x];
}

template <int MODE> void run(const char *name, int sms) {
    uint32_t *out;
    cudaMalloc(&out, 4);
    const int iters = 4096, grid = sms * 3;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SLOTS * 4);
    k<MODE><<<grid, THREADS, SLOTS * 4>>>(out, 64, 1);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<grid, THREADS, SLOTS * 4>>>(out, iters, 7);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double ops = (double)grid * THREADS * iters;
    printf("%-34s %8.3f ms  %8.1f Gops/s chip  %6.2f ops/clk/SM @1.9GHz  err=%s\n", name, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / sms / 1.9e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, 3 CTAs x 256 threads per SM, table 2^14 words\n", p.name, p.multiProcessorCount);
    int sms = p.multiProcessorCount;
    run<0>("atomicCAS random (used)", sms);
    run<1>("atomicAdd random (unused)", sms);
    run<2>("atomicAdd random (used)", sms);
    run<3>("LDS random", sms);
    run<4>("STS random", sms);
    run<5>("LDS+STS random rmw", sms);
    run<6>("atomicOr bitmap (unused)", sms);
    run<7>("atomicAdd conflict-free (unused)", sms);
    return 0;
}
