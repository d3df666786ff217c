// Ingest on the device: edge list -> CSR and per-column feature standardisation
// (reference spotify_graph.py:48-63 builds a DGLGraph from Python lists of edge endpoints;
// :77-79 standardises the stacked feature matrix on the CPU).
//
//   ps_csr_build     edges (src, dst) as listed in graph.json (both directions present, duplicates kept) ->
//                    indptr int64 [n_nodes + 1], indices int32 [E].  One stable LSD radix sort of (src, dst) pairs over the
//                    significant bits of the node id keeps the listed order inside a row, which is the successor
//                    order DGL returns (insertion order); row offsets by binary search over the sorted keys.
//   ps_standardize   x[:, j] = (x[:, j] - mean_j) / (std_j + eps), std unbiased (N - 1), two passes over the column
//                    (mean, then centred squares) accumulated in fp64 so the result does not depend on the
//                    summation order beyond fp32 rounding of the final quotient.
// Both are HBM-bound streaming passes; scratch comes from cudaMallocAsync (ingest runs once, not per step).
#include "common.cuh"
#include "../../include/pinsage_b200.h"
#include <cub/cub.cuh>

namespace {

inline unsigned grid_for(int64_t n, int block = 256) { return static_cast<unsigned>(ps_ceil_div(n > 0 ? n : 1, block)); }

__global__ void narrow_edges_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t n_edges,
                                    int64_t n_nodes, int32_t* __restrict__ k, int32_t* __restrict__ v,
                                    unsigned long long* __restrict__ bad) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_edges) return;
    const int64_t s = src[i], d = dst[i];
    if (s < 0 || s >= n_nodes || d < 0 || d >= n_nodes) {
        atomicAdd(bad, 1ull);
        k[i] = 0; v[i] = 0;
        return;
    }
    k[i] = static_cast<int32_t>(s);
    v[i] = static_cast<int32_t>(d);
}

// indptr[u] = number of sorted keys < u  (lower bound), u in [0, n_nodes]
__global__ void row_offsets_kernel(const int32_t* __restrict__ keys, int64_t n_edges, int64_t n_nodes, int64_t* __restrict__ indptr) {
    const int64_t u = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (u > n_nodes) return;
    int64_t lo = 0, hi = n_edges;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (static_cast<int64_t>(__ldg(keys + mid)) < u) lo = mid + 1; else hi = mid;
    }
    indptr[u] = lo;
}

// column sums in fp64: block = 32 columns x 8 row groups, each block walks a row range; pass 0 sums x, pass 1 sums
// (x - mean)^2
template <int kPass>
__global__ void __launch_bounds__(256)
col_stats_kernel(const float* __restrict__ x, int64_t ld, int64_t n, int d, int64_t rows_per_block,
                 const double* __restrict__ mean_sum, double* __restrict__ out) {
    __shared__ double part[8][33];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int col = blockIdx.y * 32 + lx;
    const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
    const int64_t r1 = min(n, r0 + rows_per_block);
    double s = 0.0;
    if (col < d) {
        const double mu = kPass == 1 ? mean_sum[col] / static_cast<double>(n) : 0.0;
        for (int64_t r = r0 + ly; r < r1; r += 8) {
            const double v = static_cast<double>(__ldg(x + r * ld + col));
            s += kPass == 1 ? (v - mu) * (v - mu) : v;
        }
    }
    part[ly][lx] = s;
    __syncthreads();
    if (ly == 0 && col < d) {
#pragma unroll
        for (int k = 1; k < 8; ++k) s += part[k][lx];
        atomicAdd(out + col, s);
    }
}

__global__ void standardize_kernel(float* __restrict__ x, int64_t ld, int64_t n, int d, const double* __restrict__ sum,
                                   const double* __restrict__ sq, double eps, float* __restrict__ mean_out,
                                   float* __restrict__ std_out) {
    // the reference computes mean / std in fp32 tensors and then (x - mean) / (std + eps) in fp32
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n * d) return;
    const int64_t r = i / d;
    const int c = static_cast<int>(i - r * d);
    const float mu = static_cast<float>(sum[c] / static_cast<double>(n));
    const float sd = static_cast<float>(sqrt(sq[c] / static_cast<double>(n - 1)));
    const float den = sd + static_cast<float>(eps);
    x[r * ld + c] = (x[r * ld + c] - mu) / den;
    if (r == 0) {
        if (mean_out) mean_out[c] = mu;
        if (std_out) std_out[c] = den;
    }
}

}  // namespace

extern "C" int ps_csr_build(const int64_t* src, const int64_t* dst, int64_t n_edges, int64_t n_nodes,
                            int64_t* indptr, int32_t* indices, ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(indptr != nullptr && (n_edges == 0 || (src && dst && indices)), "null pointer");
    PS_REQUIRE(n_nodes > 0 && n_nodes < (1ll << 31), "node ids must fit in 31 bits");
    PS_REQUIRE(n_edges >= 0 && n_edges < (1ll << 31), "edge list too long for one call (< 2^31 entries)");
    int end_bit = 1;
    while ((1ll << end_bit) < n_nodes && end_bit < 31) ++end_bit;
    int32_t *k_in = nullptr, *v_in = nullptr, *k_out = nullptr;
    unsigned long long* bad = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    const size_t eb = static_cast<size_t>(n_edges > 0 ? n_edges : 1) * sizeof(int32_t);
    PS_CUDA_CHECK(cudaMallocAsync(&k_in, eb, stream));
    PS_CUDA_CHECK(cudaMallocAsync(&v_in, eb, stream));
    PS_CUDA_CHECK(cudaMallocAsync(&k_out, eb, stream));
    PS_CUDA_CHECK(cudaMallocAsync(&bad, sizeof(unsigned long long), stream));
    PS_CUDA_CHECK(cudaMemsetAsync(bad, 0, sizeof(unsigned long long), stream));
    int rc = PS_OK;
    unsigned long long h_bad = 0;
    if (n_edges > 0) {
        narrow_edges_kernel<<<grid_for(n_edges), 256, 0, stream>>>(src, dst, n_edges, n_nodes, k_in, v_in, bad);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess)
            e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, indices, static_cast<int>(n_edges), 0, end_bit, stream);
        if (e == cudaSuccess) e = cudaMallocAsync(&tmp, tmp_bytes > 0 ? tmp_bytes : 1, stream);
        if (e == cudaSuccess)
            e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, indices, static_cast<int>(n_edges), 0, end_bit, stream);
        if (e != cudaSuccess) rc = ps_fail(PS_ERR_CUDA, "ps_csr_build: %s", cudaGetErrorString(e));
    }
    if (rc == PS_OK) {
        row_offsets_kernel<<<grid_for(n_nodes + 1), 256, 0, stream>>>(k_out, n_edges, n_nodes, indptr);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(&h_bad, bad, sizeof(h_bad), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) rc = ps_fail(PS_ERR_CUDA, "ps_csr_build: %s", cudaGetErrorString(e));
    }
    cudaFreeAsync(k_in, stream); cudaFreeAsync(v_in, stream); cudaFreeAsync(k_out, stream); cudaFreeAsync(bad, stream);
    if (tmp) cudaFreeAsync(tmp, stream);
    if (rc != PS_OK) return rc;
    if (h_bad != 0) return ps_fail(PS_ERR_RANGE, "%llu edge endpoint(s) out of range", h_bad);
    return PS_OK;
}

extern "C" int ps_standardize(float* x, int64_t ld, int64_t n, int d, double eps, float* mean_out, float* std_out,
                              ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(x != nullptr, "null pointer");
    PS_REQUIRE(n >= 0 && d > 0 && ld >= d, "bad shape");
    if (n == 0) return PS_OK;
    double* acc = nullptr;
    PS_CUDA_CHECK(cudaMallocAsync(&acc, static_cast<size_t>(2 * d) * sizeof(double), stream));
    PS_CUDA_CHECK(cudaMemsetAsync(acc, 0, static_cast<size_t>(2 * d) * sizeof(double), stream));
    const int col_tiles = static_cast<int>(ps_ceil_div(d, 32));
    int64_t row_blocks = ps_ceil_div(148 * 8, col_tiles);
    if (row_blocks > ps_ceil_div(n, 64)) row_blocks = ps_ceil_div(n, 64);
    if (row_blocks < 1) row_blocks = 1;
    const int64_t rpb = ps_ceil_div(n, row_blocks);
    dim3 grid(static_cast<unsigned>(ps_ceil_div(n, rpb)), static_cast<unsigned>(col_tiles));
    col_stats_kernel<0><<<grid, 256, 0, stream>>>(x, ld, n, d, rpb, nullptr, acc);
    PS_LAUNCH_CHECK();
    col_stats_kernel<1><<<grid, 256, 0, stream>>>(x, ld, n, d, rpb, acc, acc + d);
    PS_LAUNCH_CHECK();
    standardize_kernel<<<grid_for(n * d), 256, 0, stream>>>(x, ld, n, d, acc, acc + d, eps, mean_out, std_out);
    PS_LAUNCH_CHECK();
    PS_CUDA_CHECK(cudaFreeAsync(acc, stream));
    return PS_OK;
}
