// Shared helpers for libpinsage_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#define PS_OK 0
#define PS_ERR_INVALID (-1)
#define PS_ERR_CUDA (-2)
#define PS_ERR_GRAPH (-3)
#define PS_ERR_UNSUPPORTED (-4)
#define PS_ERR_NOSPACE (-5)
#define PS_ERR_RANGE (-6)

#define PS_LEAKY_SLOPE 0.01f

// thread-local last-error text, returned by ps_last_error()
char* ps_err_buf();
int ps_fail(int code, const char* fmt, ...);

#define PS_REQUIRE(cond, ...)                                 \
    do {                                                      \
        if (!(cond)) return ps_fail(PS_ERR_INVALID, __VA_ARGS__); \
    } while (0)

#define PS_CUDA_CHECK(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess)                                                           \
            return ps_fail(PS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define PS_LAUNCH_CHECK() PS_CUDA_CHECK(cudaGetLastError())

static inline int64_t ps_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float ps_leaky(float x) { return x > 0.f ? x : PS_LEAKY_SLOPE * x; }
__device__ __forceinline__ float ps_leaky_grad_from_out(float y) { return y > 0.f ? 1.f : PS_LEAKY_SLOPE; }

__device__ __forceinline__ float ps_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float4 ps_ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
