// Frontier plans on the device: the index-only work of relevant_nodes_per_layer[_precomp]
// (reference pinsage_model.py:142-168) and of the backward's (target, slot) -> z-row transpose, as a handful of
// kernels per layer instead of ~45 framework launches (cat / unique / fancy indexing / sort / searchsorted ...).
//
//   ps_plan_layer      given the neighbour ids nb[n, T] of a layer's targets `cur` (sorted, distinct), builds the
//                      next frontier uniq = sorted distinct(nb [u cur]) through a dense flag map over the id space
//                      (mark -> exclusive scan -> compact) and the positions nbz = pos(nb), self_rows = pos(cur).
//   ps_plan_transpose  sorts the pairs q = i*T + t by z-row (radix sort over only the significant key bits; stable,
//                      so q ascends inside a segment and the backward sums are reproducible), and derives the
//                      segment offsets, the 64-pair work chunks of ps_aggregate_bwd and the chunk -> row map.
// Scratch (flag / prefix arrays, CUB temporaries) lives in per-(device, stream) buffers owned by the library.
#include "common.cuh"
#include "../../include/pinsage_b200.h"
#include <cub/cub.cuh>
#include <mutex>

namespace {

struct Scratch { int dev; cudaStream_t stream; int slot; void* ptr; size_t bytes; unsigned long long used; };
Scratch g_scratch[64] = {};
unsigned long long g_scratch_clock = 0;

// grow-only buffer per (device, stream, slot); reuse across calls is safe by stream order
std::mutex g_scratch_mutex;

int scratch(cudaStream_t stream, int slot_id, size_t bytes, void** out) {
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    int dev = 0;
    PS_CUDA_CHECK(cudaGetDevice(&dev));
    Scratch* s = nullptr;
    for (auto& e : g_scratch)
        if (e.ptr != nullptr && e.dev == dev && e.stream == stream && e.slot == slot_id) { s = &e; break; }
    if (s == nullptr)
        for (auto& e : g_scratch)
            if (e.ptr == nullptr) { s = &e; break; }
    if (s == nullptr) {  // every entry is taken (streams come and go): recycle the least recently used one
        s = &g_scratch[0];
        for (auto& e : g_scratch)
            if (e.used < s->used) s = &e;
        PS_CUDA_CHECK(cudaFree(s->ptr));  // waits for whatever still reads it
        s->ptr = nullptr; s->bytes = 0;
    }
    s->used = ++g_scratch_clock;
    if (s->ptr == nullptr || s->bytes < bytes) {
        if (s->ptr != nullptr) PS_CUDA_CHECK(cudaFree(s->ptr));
        s->ptr = nullptr;
        const size_t want = bytes + bytes / 4 + 256;
        PS_CUDA_CHECK(cudaMalloc(&s->ptr, want));
        s->dev = dev; s->stream = stream; s->slot = slot_id; s->bytes = want;
    }
    *out = s->ptr;
    return PS_OK;
}

__global__ void mark_kernel(const int32_t* __restrict__ nb, int64_t n_nb, const int64_t* __restrict__ cur, int64_t n_cur,
                            int with_self, int64_t n_ids, int32_t* __restrict__ flag) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    // ids are range-checked when the table is built (NeighborTable._sanitize); an id that is out of range anyway is
    // skipped here instead of written out of bounds
    if (i < n_nb && static_cast<uint32_t>(nb[i]) < static_cast<uint32_t>(n_ids)) flag[nb[i]] = 1;
    if (with_self && i < n_cur && static_cast<uint64_t>(cur[i]) < static_cast<uint64_t>(n_ids)) flag[cur[i]] = 1;
}

__global__ void compact_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ pos, int64_t n_ids,
                               int64_t* __restrict__ uniq64, int32_t* __restrict__ uniq32) {
    const int64_t id = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (id < n_ids && flag[id]) {
        const int32_t p = pos[id];
        if (uniq64) uniq64[p] = id;
        if (uniq32) uniq32[p] = static_cast<int32_t>(id);
    }
}

__global__ void inverse_kernel(const int32_t* __restrict__ pos, const int32_t* __restrict__ nb, int64_t n_nb,
                               const int64_t* __restrict__ cur, int64_t n_cur, int with_self, int64_t n_ids,
                               int32_t* __restrict__ nbz, int32_t* __restrict__ self_rows) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n_nb) nbz[i] = static_cast<uint32_t>(nb[i]) < static_cast<uint32_t>(n_ids) ? pos[nb[i]] : 0;
    if (with_self && self_rows != nullptr && i < n_cur)
        self_rows[i] = static_cast<uint64_t>(cur[i]) < static_cast<uint64_t>(n_ids) ? pos[cur[i]] : 0;
}

__global__ void iota_kernel(int32_t* __restrict__ v, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) v[i] = static_cast<int32_t>(i);
}

// seg_off[u] = first position whose key >= u  (u in [0, nz]); chunks[u] = ceil(len(u) / chunk_pairs) for u < nz
__global__ void seg_kernel(const int32_t* __restrict__ keys, int64_t n_pairs, int64_t nz, int32_t* __restrict__ seg_off) {
    const int64_t u = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (u > nz) return;
    int64_t lo = 0, hi = n_pairs;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < u) lo = mid + 1; else hi = mid;
    }
    seg_off[u] = static_cast<int32_t>(lo);
}

__global__ void chunk_count_kernel(const int32_t* __restrict__ seg_off, int64_t nz, int chunk_pairs, int32_t* __restrict__ cnt) {
    const int64_t u = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (u > nz) return;
    cnt[u] = u < nz ? (seg_off[u + 1] - seg_off[u] + chunk_pairs - 1) / chunk_pairs : 0;
}

// chunk_row[c] = last u with chunk_off[u] <= c  (u in [0, nz])
__global__ void chunk_row_kernel(const int32_t* __restrict__ chunk_off, int64_t nz, int64_t max_chunks, int32_t* __restrict__ chunk_row) {
    const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= max_chunks) return;
    int64_t lo = 0, hi = nz + 1;  // first u with chunk_off[u] > c
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (chunk_off[mid] <= c) lo = mid + 1; else hi = mid;
    }
    chunk_row[c] = static_cast<int32_t>(lo - 1);
}

inline unsigned blocks_for(int64_t n) { return static_cast<unsigned>(ps_ceil_div(n > 0 ? n : 1, 256)); }

}  // namespace

extern "C" int ps_plan_layer(const int32_t* nb, int64_t n, int T, const int64_t* cur, int with_self, int64_t n_ids,
                             int64_t* uniq_i64, int32_t* uniq_i32, int32_t* nbz, int32_t* self_rows, int32_t* count_out,
                             ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(nb && cur && nbz && count_out, "null pointer");
    PS_REQUIRE(n >= 0 && T > 0 && n_ids > 0 && n_ids < (1ll << 31) && n * T < (1ll << 31), "bad shape");
    void* p = nullptr;
    int rc = scratch(stream, 0, static_cast<size_t>(2 * (n_ids + 1)) * sizeof(int32_t), &p);
    if (rc != PS_OK) return rc;
    int32_t* flag = static_cast<int32_t*>(p);
    int32_t* pos = flag + (n_ids + 1);
    PS_CUDA_CHECK(cudaMemsetAsync(flag, 0, static_cast<size_t>(n_ids + 1) * sizeof(int32_t), stream));
    const int64_t n_nb = n * T;
    if (n > 0) {
        mark_kernel<<<blocks_for(n_nb), 256, 0, stream>>>(nb, n_nb, cur, n, with_self, n_ids, flag);
        PS_LAUNCH_CHECK();
    }
    size_t tmp_bytes = 0;
    PS_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flag, pos, static_cast<int>(n_ids + 1), stream));
    void* tmp = nullptr;
    rc = scratch(stream, 1, tmp_bytes, &tmp);
    if (rc != PS_OK) return rc;
    PS_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flag, pos, static_cast<int>(n_ids + 1), stream));
    compact_kernel<<<blocks_for(n_ids), 256, 0, stream>>>(flag, pos, n_ids, uniq_i64, uniq_i32);
    PS_LAUNCH_CHECK();
    if (n > 0) {
        inverse_kernel<<<blocks_for(n_nb), 256, 0, stream>>>(pos, nb, n_nb, cur, n, with_self, n_ids, nbz, self_rows);
        PS_LAUNCH_CHECK();
    }
    PS_CUDA_CHECK(cudaMemcpyAsync(count_out, pos + n_ids, sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
    return PS_OK;
}

extern "C" int ps_plan_transpose(const int32_t* nbz, int64_t n_pairs, int64_t nz, int chunk_pairs,
                                 int32_t* pair_q, int32_t* seg_off, int32_t* chunk_off, int32_t* chunk_row, int64_t max_chunks,
                                 ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(nbz && pair_q && seg_off && chunk_off && chunk_row, "null pointer");
    PS_REQUIRE(n_pairs >= 0 && n_pairs < (1ll << 31) && nz >= 0 && nz < (1ll << 31) && chunk_pairs > 0, "bad shape");
    int end_bit = 1;
    while ((1ll << end_bit) < nz && end_bit < 31) ++end_bit;
    void* p = nullptr;
    int rc = scratch(stream, 2, static_cast<size_t>(2 * n_pairs + nz + 2) * sizeof(int32_t), &p);
    if (rc != PS_OK) return rc;
    int32_t* iota = static_cast<int32_t*>(p);
    int32_t* keys_sorted = iota + n_pairs;
    int32_t* cnt = keys_sorted + n_pairs;
    if (n_pairs > 0) {
        iota_kernel<<<blocks_for(n_pairs), 256, 0, stream>>>(iota, n_pairs);
        PS_LAUNCH_CHECK();
        size_t tmp_bytes = 0;
        PS_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, nbz, keys_sorted, iota, pair_q, static_cast<int>(n_pairs), 0, end_bit, stream));
        void* tmp = nullptr;
        rc = scratch(stream, 3, tmp_bytes, &tmp);
        if (rc != PS_OK) return rc;
        PS_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, nbz, keys_sorted, iota, pair_q, static_cast<int>(n_pairs), 0, end_bit, stream));
    }
    seg_kernel<<<blocks_for(nz + 1), 256, 0, stream>>>(keys_sorted, n_pairs, nz, seg_off);
    PS_LAUNCH_CHECK();
    chunk_count_kernel<<<blocks_for(nz + 1), 256, 0, stream>>>(seg_off, nz, chunk_pairs, cnt);
    PS_LAUNCH_CHECK();
    size_t tmp_bytes = 0;
    PS_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, chunk_off, static_cast<int>(nz + 1), stream));
    void* tmp = nullptr;
    rc = scratch(stream, 1, tmp_bytes, &tmp);
    if (rc != PS_OK) return rc;
    PS_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, chunk_off, static_cast<int>(nz + 1), stream));
    if (max_chunks > 0) {
        chunk_row_kernel<<<blocks_for(max_chunks), 256, 0, stream>>>(chunk_off, nz, max_chunks, chunk_row);
        PS_LAUNCH_CHECK();
    }
    return PS_OK;
}
