// K1 + K2: restart random walks over the bipartite track-collection CSR, fused with the
// visit-count -> top-T reduction.  Replaces the Python double loop of
// do_random_walks (reference pinsage_model.py:32-53) and the dense [n, N+C] float64
// histogram + torch.topk of sample_neighborhood[_topt] (:88-107).
//
// Design (B200): one warp per source node.  The chain of a source is a sequence of i.i.d.
// segments that all start at the source; the restart flag of step j depends only on
// Philox(seed, source, j), so the 32 lanes find the segment starts of a 32-step chunk
// with one shuffle and walk the segments independently.  The trace of the source lives
// in shared memory (n_hops x 4 B per warp) and never goes to HBM; the histogram is a
// warp-local bitonic sort + run-length count, the top-T a second sort of (count, run
// head) keys -- the dense row of the reference is never materialised.
// HBM traffic per step is the two CSR hops (indptr pair + one neighbour id each).
#include "common.cuh"
#include "../../include/pinsage_b200.h"

struct ps_graph {
    const int64_t* indptr;
    const int32_t* indices;
    int64_t n_tracks, n_cols, n_entries;
};

namespace {

constexpr int kMaxWarpsPerCta = 8;  // fewer when the per-warp trace (8 B x pow2(n_hops)) is large
constexpr uint32_t kPad = 0xFFFFFFFFu;

struct Philox {
    uint32_t x0, x1, x2, x3;
};

__device__ __forceinline__ Philox philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;  // the bump after the last round is dead code
    }
    return {c0, c1, c2, c3};
}

// one CSR hop: uniform successor of `node` picked by x
__device__ __forceinline__ uint32_t hop(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                        uint32_t node, uint32_t x) {
    const int64_t beg = __ldg(indptr + node);
    const int64_t end = __ldg(indptr + node + 1);
    const uint32_t deg = static_cast<uint32_t>(end - beg);
    if (deg == 0) return node;  // unreachable after ps_graph_create's degree check
    return static_cast<uint32_t>(__ldg(indices + beg + __umulhi(x, deg)));
}

// warp-cooperative bitonic sort of a[0..P) in shared memory, P a power of two
template <bool kDescending>
__device__ __forceinline__ void warp_bitonic_sort(uint32_t* a, int P, int lane) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (P >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool up = ((i & k) == 0) != kDescending;
                const uint32_t x = a[i], y = a[p];
                if ((x > y) == up) { a[i] = y; a[p] = x; }
            }
            __syncwarp();
        }
    }
}

// kFromTrace: read the steps from a caller-supplied trace instead of walking (parity hook)
template <bool kFromTrace>
__global__ void __launch_bounds__(kMaxWarpsPerCta * 32)
walk_topt_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                 const int64_t* __restrict__ sources, const int64_t* __restrict__ in_trace,
                 int64_t n, int n_hops, int P, uint64_t restart_thr, int fixed_len, int T,
                 uint32_t k0, uint32_t k1,
                 int64_t* __restrict__ out_nodes64, double* __restrict__ out_w64,
                 int32_t* __restrict__ out_nodes32, float* __restrict__ out_w32,
                 int32_t* __restrict__ out_trace) {
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint32_t* ids = smem + static_cast<size_t>(warp) * 2 * P;  // trace, then sorted ids
    uint32_t* keys = ids + P;                                 // (count << 16) | (0xFFFF - head position)

    const int warps_per_cta = blockDim.x >> 5;
    for (int64_t s = static_cast<int64_t>(blockIdx.x) * warps_per_cta + warp; s < n;
         s += static_cast<int64_t>(gridDim.x) * warps_per_cta) {
        const uint32_t src = static_cast<uint32_t>(sources[s]);

        if (kFromTrace) {
            for (int j = lane; j < P; j += 32)
                ids[j] = j < n_hops ? static_cast<uint32_t>(in_trace[s * n_hops + j]) : kPad;
        } else {
            bool carry = true;  // "the step before this chunk restarted" (step 0 starts at the source)
            for (int base = 0; base < n_hops; base += 32) {
                const int j = base + lane;
                const bool valid = j < n_hops;
                Philox d = philox4x32_10(static_cast<uint32_t>(j), src, 0u, 0u, k0, k1);
                bool r = !valid || (fixed_len > 0 ? ((j + 1) % fixed_len == 0)
                                                  : (static_cast<uint64_t>(d.x2) < restart_thr));
                bool prev = __shfl_up_sync(0xffffffffu, r, 1);
                if (lane == 0) prev = carry;
                carry = __shfl_sync(0xffffffffu, r, 31);
                if (valid && prev) {  // this lane owns the segment that starts at step j
                    uint32_t item = src;
                    int jj = j;
                    while (true) {
                        const uint32_t col = hop(indptr, indices, item, d.x0);
                        item = hop(indptr, indices, col, d.x1);
                        ids[jj] = item;
                        if (r || jj + 1 >= n_hops) break;
                        ++jj;
                        d = philox4x32_10(static_cast<uint32_t>(jj), src, 0u, 0u, k0, k1);
                        r = fixed_len > 0 ? ((jj + 1) % fixed_len == 0)
                                          : (static_cast<uint64_t>(d.x2) < restart_thr);
                    }
                }
            }
            for (int j = n_hops + lane; j < P; j += 32) ids[j] = kPad;
        }
        __syncwarp();
        if (out_trace != nullptr)
            for (int j = lane; j < n_hops; j += 32) out_trace[s * n_hops + j] = static_cast<int32_t>(ids[j]);
        __syncwarp();

        // histogram by sorting: equal ids become runs, run length = visit count
        warp_bitonic_sort<false>(ids, P, lane);
        for (int i = lane; i < P; i += 32) {
            const uint32_t v = ids[i];
            uint32_t key = 0;
            if (v != kPad && v != src && (i == 0 || ids[i - 1] != v)) {
                int lo = i + 1, hi = P;  // first index in (i, P] whose id exceeds v
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (ids[mid] <= v) lo = mid + 1; else hi = mid;
                }
                key = (static_cast<uint32_t>(lo - i) << 16) | (0xFFFFu - static_cast<uint32_t>(i));
            }
            keys[i] = key;
        }
        __syncwarp();
        // (count desc, id asc): ids are sorted ascending, so a smaller head position is a smaller id
        warp_bitonic_sort<true>(keys, P, lane);

        for (int t = lane; t < T; t += 32) {
            const uint32_t key = t < P ? keys[t] : 0u;
            uint32_t node = src;
            uint32_t count = 0;
            if (key != 0u) {
                count = key >> 16;
                node = ids[0xFFFFu - (key & 0xFFFFu)];
            }
            const int64_t o = s * T + t;
            if (out_nodes64) out_nodes64[o] = static_cast<int64_t>(node);
            if (out_w64) out_w64[o] = static_cast<double>(count) / static_cast<double>(n_hops);
            if (out_nodes32) out_nodes32[o] = static_cast<int32_t>(node);
            if (out_w32) out_w32[o] = static_cast<float>(static_cast<double>(count) / static_cast<double>(n_hops));
        }
        __syncwarp();
    }
}

__global__ void count_zero_degree_kernel(const int64_t* __restrict__ indptr, int64_t n_nodes, int64_t n_entries,
                                         unsigned long long* __restrict__ bad) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const int64_t a = indptr[i], b = indptr[i + 1];
    if (b <= a || a < 0 || b > n_entries) atomicAdd(bad, 1ull);
}

int next_pow2(int x) {
    int p = 32;
    while (p < x) p <<= 1;
    return p;
}

template <bool kFromTrace>
int launch_walk(const int64_t* indptr, const int32_t* indices, const int64_t* sources, const int64_t* trace,
                int64_t n, int n_hops, double alpha, int fixed_len, int T, uint64_t seed,
                int64_t* on64, double* ow64, int32_t* on32, float* ow32, int32_t* otrace, cudaStream_t stream) {
    PS_REQUIRE(n >= 0 && n_hops > 0 && n_hops <= 16384, "n_hops must be in [1, 16384] (got %d)", n_hops);
    PS_REQUIRE(T > 0, "T must be positive");
    PS_REQUIRE(alpha >= 0.0 && alpha <= 1.0, "alpha must be in [0, 1]");
    if (n == 0) return PS_OK;
    const int P = next_pow2(n_hops);
    const size_t per_warp = 2 * static_cast<size_t>(P) * sizeof(uint32_t);
    int warps = static_cast<int>((200 * 1024) / per_warp);
    if (warps > kMaxWarpsPerCta) warps = kMaxWarpsPerCta;
    PS_REQUIRE(warps >= 1, "n_hops=%d too large for the shared-memory trace (max 16384)", n_hops);
    const size_t smem = static_cast<size_t>(warps) * per_warp;
    auto kern = walk_topt_kernel<kFromTrace>;
    if (smem > 48 * 1024)
        PS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int dev = 0, sms = 148, occ = 1;
    PS_CUDA_CHECK(cudaGetDevice(&dev));
    PS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    PS_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, smem));
    if (occ < 1) occ = 1;
    int64_t blocks = ps_ceil_div(n, warps);
    const int64_t resident = static_cast<int64_t>(sms) * occ;  // persistent: grid-stride over sources
    if (blocks > resident * 4) blocks = resident * 4;
    const uint64_t thr = static_cast<uint64_t>(alpha * 4294967296.0);
    kern<<<static_cast<unsigned>(blocks), warps * 32, smem, stream>>>(
        indptr, indices, sources, trace, n, n_hops, P, thr, fixed_len, T,
        static_cast<uint32_t>(seed & 0xFFFFFFFFull), static_cast<uint32_t>(seed >> 32), on64, ow64, on32, ow32, otrace);
    PS_LAUNCH_CHECK();
    return PS_OK;
}

}  // namespace

extern "C" int ps_graph_create(const int64_t* indptr, const int32_t* indices, int64_t n_tracks, int64_t n_cols,
                               int64_t n_entries, ps_graph_t** out, ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(out != nullptr && indptr != nullptr && indices != nullptr, "null pointer");
    PS_REQUIRE(n_tracks > 0 && n_cols > 0 && n_entries > 0, "empty graph");
    PS_REQUIRE(n_tracks + n_cols < (1ll << 31), "node ids must fit in 31 bits");
    unsigned long long* bad = nullptr;
    PS_CUDA_CHECK(cudaMallocAsync(&bad, sizeof(unsigned long long), stream));
    PS_CUDA_CHECK(cudaMemsetAsync(bad, 0, sizeof(unsigned long long), stream));
    const int64_t n_nodes = n_tracks + n_cols;
    count_zero_degree_kernel<<<static_cast<unsigned>(ps_ceil_div(n_nodes, 256)), 256, 0, stream>>>(indptr, n_nodes, n_entries, bad);
    PS_LAUNCH_CHECK();
    unsigned long long h_bad = 0;
    PS_CUDA_CHECK(cudaMemcpyAsync(&h_bad, bad, sizeof(h_bad), cudaMemcpyDeviceToHost, stream));
    PS_CUDA_CHECK(cudaStreamSynchronize(stream));
    PS_CUDA_CHECK(cudaFreeAsync(bad, stream));
    if (h_bad != 0)
        return ps_fail(PS_ERR_GRAPH, "%llu node(s) have no successors or a malformed indptr; the reference's walker raises on them (pinsage_model.py:42)", h_bad);
    ps_graph* g = new ps_graph{indptr, indices, n_tracks, n_cols, n_entries};
    *out = g;
    return PS_OK;
}

extern "C" int ps_graph_destroy(ps_graph_t* g) {
    delete g;
    return PS_OK;
}

extern "C" int ps_walk_topt(const ps_graph_t* g, const int64_t* sources, int64_t n, int n_hops, double alpha,
                            int fixed_len, int T, uint64_t seed, int64_t* out_nodes_i64, double* out_w_f64,
                            int32_t* out_nodes_i32, float* out_w_f32, int32_t* out_trace, ps_stream_t stream) {
    PS_REQUIRE(g != nullptr && (sources != nullptr || n == 0), "null pointer");
    PS_REQUIRE(fixed_len >= 0, "fixed_len must be >= 0");
    return launch_walk<false>(g->indptr, g->indices, sources, nullptr, n, n_hops, alpha, fixed_len, T, seed,
                              out_nodes_i64, out_w_f64, out_nodes_i32, out_w_f32, out_trace,
                              static_cast<cudaStream_t>(stream));
}

extern "C" int ps_trace_topt(const int64_t* trace, const int64_t* sources, int64_t n, int n_hops, int T,
                             int64_t* out_nodes_i64, double* out_w_f64, int32_t* out_nodes_i32, float* out_w_f32,
                             ps_stream_t stream) {
    PS_REQUIRE((trace != nullptr && sources != nullptr) || n == 0, "null pointer");
    return launch_walk<true>(nullptr, nullptr, sources, trace, n, n_hops, 0.0, 0, T, 0ull,
                             out_nodes_i64, out_w_f64, out_nodes_i32, out_w_f32, nullptr,
                             static_cast<cudaStream_t>(stream));
}
