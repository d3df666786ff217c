// K14: per-row top-k of a similarity tile, the selection half of the
// cosine kNN search of the reference's evaluation (baselines.py:91-103: cosine_sim.topk(k+1, dim=1), k = 1000
// over N columns).  The similarity tile itself is a ps_gemm (tcgen05); this kernel replaces torch.topk.
//
// One CTA per row.  The k-th largest value is found by an MSB-first radix select over the order-preserving
// 32-bit image of the floats (3 passes with 12 / 12 / 8-bit digits and shared-memory histograms -- 4096 bins in
// the first pass so that values sharing an exponent still spread over 16 bins -- each pass streams the row with
// 128-bit loads); one more pass collects the elements above the threshold plus the lowest-index elements equal
// to it, and the <= k winners are sorted in shared memory (bitonic, 64-bit keys = value | ~index), i.e.
// descending by value, ascending by column on ties.  HBM traffic: 4 reads of the row; nothing else is written
// than the k results.  NaNs order above +inf (as the largest values), like torch.topk.
#include "common.cuh"
#include "../../include/pinsage_b200.h"

namespace {

constexpr int kThreads = 512;
constexpr int kEqCap = 1024;  // elements equal to the k-th value that are ranked by column; beyond that, first come

__device__ __forceinline__ uint32_t order_key(float v) {
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone: larger float <-> larger key
}
__device__ __forceinline__ float key_value(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

__global__ void __launch_bounds__(kThreads)
topk_rows_kernel(const float* __restrict__ x, int64_t ld, int64_t n_cols, int k, int kp2,
                 const int32_t* __restrict__ col_ids, const int32_t* __restrict__ row_counts,
                 float* __restrict__ out_val, int64_t* __restrict__ out_idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* sel = reinterpret_cast<uint64_t*>(smem_raw);           // [kp2] winners
    uint32_t* eq_idx = reinterpret_cast<uint32_t*>(sel + kp2);         // [kEqCap] columns equal to the threshold
    uint32_t* hist = eq_idx + kEqCap;                                  // [4096]
    __shared__ uint32_t s_prefix, s_need, s_nsel, s_neq;
    const int tid = threadIdx.x;
    const float* row = x + static_cast<int64_t>(blockIdx.x) * ld;
    // candidate lists (ps_topk_rows_mapped): only the first row_counts[row] entries are valid and entry i stands for
    // column col_ids[row, i] (ranking on ties and the reported index use that id)
    const int32_t* ids = col_ids ? col_ids + static_cast<int64_t>(blockIdx.x) * ld : nullptr;
    if (row_counts != nullptr) n_cols = min(n_cols, static_cast<int64_t>(max(__ldg(row_counts + blockIdx.x), 0)));
    const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const int64_t n4 = vec ? n_cols / 4 : 0;

    if (tid == 0) { s_prefix = 0; s_need = static_cast<uint32_t>(k); s_nsel = 0; s_neq = 0; }
    __syncthreads();
    // ---- radix select of the k-th largest key: digits of 12, 12 and 8 bits
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = pass == 0 ? 20 : (pass == 1 ? 8 : 0);
        const int bins = pass == 2 ? 256 : 4096;
        for (int b = tid; b < bins; b += kThreads) hist[b] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t hi_mask = pass == 0 ? 0u : (pass == 1 ? 0xFFF00000u : 0xFFFFFF00u);
        const uint32_t dmask = static_cast<uint32_t>(bins - 1);
        auto count = [&](float v) {
            const uint32_t key = order_key(v);
            if ((key & hi_mask) == prefix) atomicAdd(hist + ((key >> shift) & dmask), 1u);
        };
        for (int64_t i = tid; i < n4; i += kThreads) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(row) + i);
            count(v.x); count(v.y); count(v.z); count(v.w);
        }
        for (int64_t i = n4 * 4 + tid; i < n_cols; i += kThreads) count(__ldg(row + i));
        __syncthreads();
        if (tid < 32) {  // one warp walks the bins from the top: the bin that holds the s_need-th element
            const int lane = tid, per = bins / 32;
            uint32_t sum = 0;
            for (int q = 0; q < per; ++q) sum += hist[bins - 1 - per * lane - q];
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t excl = incl - sum, need = s_need;
            if (excl < need && need <= incl) {
                uint32_t run = excl;
                for (int q = 0; q < per; ++q) {
                    const uint32_t c = hist[bins - 1 - per * lane - q];
                    if (run < need && need <= run + c) {
                        s_prefix = prefix | (static_cast<uint32_t>(bins - 1 - per * lane - q) << shift);
                        s_need = need - run;
                    }
                    run += c;
                }
            }
        }
        __syncthreads();
    }
    const uint32_t thr = s_prefix;  // key of the k-th largest; s_need = how many elements equal to it are wanted
    // ---- collect: everything above the threshold, and the columns that equal it
    auto take = [&](float v, int64_t col) {
        const uint32_t key = order_key(v);
        if (ids != nullptr) col = __ldg(ids + col);
        if (key > thr) {
            const uint32_t p = atomicAdd(&s_nsel, 1u);
            sel[p] = (static_cast<uint64_t>(key) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(col));
        } else if (key == thr) {
            const uint32_t p = atomicAdd(&s_neq, 1u);
            if (p < kEqCap) eq_idx[p] = static_cast<uint32_t>(col);
        }
    };
    for (int64_t i = tid; i < n4; i += kThreads) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row) + i);
        take(v.x, 4 * i); take(v.y, 4 * i + 1); take(v.z, 4 * i + 2); take(v.w, 4 * i + 3);
    }
    for (int64_t i = n4 * 4 + tid; i < n_cols; i += kThreads) take(__ldg(row + i), i);
    __syncthreads();
    // the wanted equals are the ones with the smallest columns: rank the (few) candidates by counting
    const uint32_t n_gt = s_nsel, need_eq = s_need;
    const uint32_t n_eq = s_neq < static_cast<uint32_t>(kEqCap) ? s_neq : static_cast<uint32_t>(kEqCap);
    for (uint32_t i = tid; i < n_eq; i += kThreads) {
        const uint32_t c = eq_idx[i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n_eq; ++j) rank += eq_idx[j] < c;
        if (rank < need_eq) sel[n_gt + rank] = (static_cast<uint64_t>(thr) << 32) | (0xFFFFFFFFu - c);
    }
    uint32_t total = n_gt + (need_eq < n_eq ? need_eq : n_eq);
    for (uint32_t i = total + tid; i < static_cast<uint32_t>(kp2); i += kThreads) sel[i] = 0ull;
    __syncthreads();
    // ---- bitonic sort, descending
    for (int kk = 2; kk <= kp2; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (kp2 >> 1); t += kThreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool down = (i & kk) == 0;
                const uint64_t a = sel[i], b = sel[p];
                if ((a < b) == down) { sel[i] = b; sel[p] = a; }
            }
            __syncthreads();
        }
    }
    for (int t = tid; t < k; t += kThreads) {
        const uint64_t e = sel[t];
        const int64_t o = static_cast<int64_t>(blockIdx.x) * k + t;
        out_val[o] = e ? key_value(static_cast<uint32_t>(e >> 32)) : -INFINITY;
        out_idx[o] = e ? static_cast<int64_t>(0xFFFFFFFFu - static_cast<uint32_t>(e)) : -1;
    }
}

}  // namespace

extern "C" int ps_topk_rows(const float* x, int64_t ld, int64_t n_rows, int64_t n_cols, int k,
                            float* out_val, int64_t* out_idx, ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(x && out_val && out_idx, "null pointer");
    PS_REQUIRE(k > 0 && k <= 8192 && k <= n_cols, "k must be in [1, min(8192, n_cols)] (got %d)", k);
    PS_REQUIRE(n_cols < (1ll << 32) && ld >= n_cols, "bad row shape");
    if (n_rows == 0) return PS_OK;
    int kp2 = 32;
    while (kp2 < k) kp2 <<= 1;
    const size_t smem = static_cast<size_t>(kp2) * 8 + kEqCap * 4 + 4096 * 4;
    if (smem > 48 * 1024)
        PS_CUDA_CHECK(cudaFuncSetAttribute(topk_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    topk_rows_kernel<<<static_cast<unsigned>(n_rows), kThreads, smem, stream>>>(x, ld, n_cols, k, kp2, nullptr, nullptr, out_val, out_idx);
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// The same selection over per-row candidate lists (the output of ps_gemm_filter): row r holds row_counts[r] valid
// (value, column id) pairs; rows with fewer than k candidates are padded with (-inf, -1).
extern "C" int ps_topk_rows_mapped(const float* x, const int32_t* col_ids, const int32_t* row_counts, int64_t ld,
                                   int64_t n_rows, int k, float* out_val, int64_t* out_idx, ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(x && col_ids && row_counts && out_val && out_idx, "null pointer");
    PS_REQUIRE(k > 0 && k <= 8192 && k <= ld, "k must be in [1, min(8192, ld)] (got %d)", k);
    if (n_rows == 0) return PS_OK;
    int kp2 = 32;
    while (kp2 < k) kp2 <<= 1;
    const size_t smem = static_cast<size_t>(kp2) * 8 + kEqCap * 4 + 4096 * 4;
    if (smem > 48 * 1024)
        PS_CUDA_CHECK(cudaFuncSetAttribute(topk_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    topk_rows_kernel<<<static_cast<unsigned>(n_rows), kThreads, smem, stream>>>(x, ld, ld, k, kp2, col_ids, row_counts, out_val, out_idx);
    PS_LAUNCH_CHECK();
    return PS_OK;
}
