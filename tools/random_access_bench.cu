// Random 4-byte gathers from a table far larger than L2: the DRAM random-access ceiling the walker's
// collection -> item hop runs against (development tool; profiles/r2*_random_access.txt).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/rab tools/random_access_bench.cu && /tmp/rab
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int U>
__global__ void gather_kernel(const uint32_t* __restrict__ table, uint32_t n_words, uint32_t iters, uint32_t* __restrict__ out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0, s = mix(tid * 2654435761u + 1u);
    for (uint32_t it = 0; it < iters; ++it) {
        uint32_t v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = s * 1664525u + 1013904223u;
            v[u] = __ldg(table + __umulhi(mix(s), n_words));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u];
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// dependent chain of two gathers (index -> index), like hop -> hop
__global__ void chase_kernel(const uint32_t* __restrict__ table, uint32_t n_words, uint32_t iters, uint32_t* __restrict__ out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0, s = mix(tid * 2654435761u + 1u);
    for (uint32_t it = 0; it < iters; ++it) {
        s = s * 1664525u + 1013904223u;
        const uint32_t a = __ldg(table + __umulhi(mix(s), n_words));
        const uint32_t b = __ldg(table + __umulhi(mix(a ^ s), n_words));
        acc += b;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

int main() {
    const size_t sizes_mb[] = {160, 320, 2048};
    uint32_t* out; cudaMalloc(&out, 4);
    for (int gran : {0, 32}) {
        if (gran) {
            cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
            size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
            printf("# cudaLimitMaxL2FetchGranularity set to %d: %s, now %zu\n", gran, cudaGetErrorString(e), g);
        } else {
            size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
            printf("# default cudaLimitMaxL2FetchGranularity = %zu\n", g);
        }
        for (size_t mb : sizes_mb) {
            const uint32_t n_words = static_cast<uint32_t>(mb * 1024 * 1024 / 4);
            uint32_t* table; cudaMalloc(&table, size_t(n_words) * 4);
            cudaMemset(table, 0x5a, size_t(n_words) * 4);
            for (int warps_per_sm : {16, 32, 64}) {
                const int threads = 256, blocks = 148 * warps_per_sm * 32 / threads;
                const uint32_t iters = 64;
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                auto run = [&](int which) {
                    float best = 1e30f;
                    for (int rep = 0; rep < 4; ++rep) {
                        cudaEventRecord(e0);
                        if (which == 1) gather_kernel<1><<<blocks, threads>>>(table, n_words, iters * 4, out);
                        else if (which == 4) gather_kernel<4><<<blocks, threads>>>(table, n_words, iters, out);
                        else chase_kernel<<<blocks, threads>>>(table, n_words, iters * 2, out);
                        cudaEventRecord(e1); cudaEventSynchronize(e1);
                        float ms; cudaEventElapsedTime(&ms, e0, e1);
                        if (rep > 0 && ms < best) best = ms;
                    }
                    return best;
                };
                const double n_acc = double(blocks) * threads * iters * 4;
                const float t1 = run(1), t4 = run(4), tc = run(0);
                printf("table %5zu MB  %2d warps/SM  1-in-flight: %6.2f G acc/s  4-in-flight: %6.2f G acc/s  2-chain: %6.2f G acc/s   (x32 B = %5.0f / %5.0f GB/s, x64 B = %5.0f / %5.0f GB/s)\n",
                       mb, warps_per_sm, n_acc / t1 / 1e6, n_acc / t4 / 1e6, n_acc / tc / 1e6,
                       n_acc / t1 / 1e6 * 32, n_acc / t4 / 1e6 * 32, n_acc / t1 / 1e6 * 64, n_acc / t4 / 1e6 * 64);
            }
            cudaFree(table);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("# %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
