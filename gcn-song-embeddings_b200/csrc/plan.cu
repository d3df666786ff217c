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

// ---- kernels of ps_prepare_plan
__global__ void mark_ids64_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t n_ids, int32_t* __restrict__ flag,
                                  int32_t* __restrict__ bad) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t id = ids[i];
    if (static_cast<uint64_t>(id) < static_cast<uint64_t>(n_ids)) flag[id] = 1;
    else atomicAdd(bad, 1);
}

__global__ void positions64_kernel(const int32_t* __restrict__ pos, const int64_t* __restrict__ ids, int64_t n, int32_t* __restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = pos[ids[i]];
}

// nb[i, t] = table_nodes[cur[i], t], w[i, t] = table_w[cur[i], t] for t < T   (NeighborTable.lookup)
__global__ void lookup_kernel(const int64_t* __restrict__ cur, int64_t n, const int32_t* __restrict__ tab_nodes,
                              const float* __restrict__ tab_w, int Tp, int T, int32_t* __restrict__ nb, float* __restrict__ w) {
    const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n * T) return;
    const int64_t i = e / T;
    const int t = static_cast<int>(e - i * T);
    const int64_t src = cur[i] * Tp + t;
    nb[e] = __ldg(tab_nodes + src);
    w[e] = __ldg(tab_w + src);
}

__global__ void narrow_ids_kernel(const int64_t* __restrict__ in, int64_t n, int32_t* __restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = static_cast<int32_t>(in[i]);
}

struct PlanArena {
    char* base; int64_t cap, off;
    template <typename T> bool take(int64_t count, T** out, int64_t* off_out) {
        const int64_t bytes = (count * static_cast<int64_t>(sizeof(T)) + 255) & ~static_cast<int64_t>(255);
        *off_out = off;
        *out = reinterpret_cast<T*>(base + off);
        off += bytes;
        return off <= cap;
    }
};

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


// The whole index-only preparation of a training batch in ONE host call (Engine.prepare composed it from ~25 calls):
//   top      = sorted distinct node ids of batch [B,3]           (the shared frontier's top layer)
//   triples  = position of every batch entry in `top`, counts = per-column occurrences (the duplicate-node factor)
//   per layer l = L-1 .. 0 (relevant_nodes_per_layer_precomp, pinsage_model.py:156-168): neighbourhood lookup of the
//   layer's targets, next frontier through the dense flag map, positions, and the backward transpose.
// Sizes are only known on the device, so the call synchronises `stream` L+1 times (one 4-byte read each) and lays the
// outputs out back to back in the caller's arena; `out` receives sizes and byte offsets.  Returns PS_ERR_NOSPACE (with
// out->bytes_needed = a size that is certainly enough for the part reached) when the arena is too small: grow and retry.
namespace {
// where a layer's neighbourhoods come from: the precomputed table (reference default) or the walker run on the
// layer's targets (online sampling, pinsage_model.py:142-154)
struct NeighbourSource {
    const int32_t* tab_nodes; const float* tab_w; int Tp;         // table
    const ps_graph_t* graph; int n_hops; double alpha; uint64_t seed;  // walker (graph != nullptr)
};
}  // namespace

static int prepare_plan_impl(const int64_t* batch, int64_t B, const NeighbourSource& nsrc, int64_t n_ids,
                             int T, int n_layers, int need_backward, void* arena_, int64_t arena_bytes, ps_plan_desc* out,
                             ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int32_t* tab_nodes = nsrc.tab_nodes;
    const float* tab_w = nsrc.tab_w;
    const int Tp = nsrc.graph != nullptr ? T : nsrc.Tp;
    PS_REQUIRE(batch && (nsrc.graph != nullptr || (tab_nodes && tab_w)) && arena_ && out, "null pointer");
    PS_REQUIRE(B > 0 && T > 0 && T <= Tp && n_layers >= 1 && n_layers <= PS_MAX_LAYERS, "bad shape (T must not exceed the table width)");
    PS_REQUIRE(n_ids > 0 && n_ids < (1ll << 31), "bad id space");
    memset(out, 0, sizeof(*out));
    PlanArena ar{static_cast<char*>(arena_), arena_bytes, 0};
    void* p = nullptr;
    int rc = scratch(stream, 0, static_cast<size_t>(2 * (n_ids + 1) + 2) * sizeof(int32_t), &p);
    if (rc != PS_OK) return rc;
    int32_t* flag = static_cast<int32_t*>(p);
    int32_t* pos = flag + (n_ids + 1);
    int32_t* bad = pos + (n_ids + 1);
    size_t scan_bytes = 0;
    PS_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, flag, pos, static_cast<int>(n_ids + 1), stream));
    void* scan_tmp = nullptr;
    rc = scratch(stream, 1, scan_bytes, &scan_tmp);
    if (rc != PS_OK) return rc;
    int32_t h_count[2] = {0, 0};
#define PS_NOSPACE(need_more)                                                   \
    do {                                                                        \
        out->bytes_needed = ar.off + static_cast<int64_t>(need_more) + 4096;    \
        return PS_ERR_NOSPACE;                                                  \
    } while (0)

    // ---- top = unique(batch), triples, counts
    const int64_t nb3 = 3 * B;
    PS_CUDA_CHECK(cudaMemsetAsync(flag, 0, static_cast<size_t>(2 * (n_ids + 1) + 2) * sizeof(int32_t), stream));
    mark_ids64_kernel<<<blocks_for(nb3), 256, 0, stream>>>(batch, nb3, n_ids, flag, bad);
    PS_LAUNCH_CHECK();
    PS_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, flag, pos, static_cast<int>(n_ids + 1), stream));
    PS_CUDA_CHECK(cudaMemcpyAsync(&h_count[0], pos + n_ids, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    PS_CUDA_CHECK(cudaMemcpyAsync(&h_count[1], bad, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    PS_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (h_count[1] != 0) return ps_fail(PS_ERR_RANGE, "node id out of range");
    const int64_t U = h_count[0];
    out->U = U;
    int64_t* top = nullptr; int32_t* triples = nullptr; int32_t* counts = nullptr;
    bool ok = ar.take<int64_t>(U, &top, &out->off_top);
    ok = ar.take<int32_t>(nb3, &triples, &out->off_triples) && ok;
    ok = ar.take<int32_t>(3 * U, &counts, &out->off_counts) && ok;
    if (!ok) PS_NOSPACE(0);
    compact_kernel<<<blocks_for(n_ids), 256, 0, stream>>>(flag, pos, n_ids, top, nullptr);
    PS_LAUNCH_CHECK();
    positions64_kernel<<<blocks_for(nb3), 256, 0, stream>>>(pos, batch, nb3, triples);
    PS_LAUNCH_CHECK();
    rc = ps_count_triples(triples, B, U, counts, stream_);
    if (rc != PS_OK) return rc;

    // ---- layers, top down
    const int64_t* cur = top;
    int64_t n = U;
    for (int l = n_layers - 1; l >= 0; --l) {
        ps_plan_desc_layer& d = out->layers[l];
        const bool with_self = l > 0;
        const int64_t n_nb = n * T;
        PS_REQUIRE(n_nb < (1ll << 31), "too many (target, slot) pairs in one layer");
        int32_t* nb = nullptr; float* w = nullptr; int32_t* nbz = nullptr; int32_t* self_rows = nullptr;
        int64_t off_nb_tmp = 0;
        d.n = n;
        d.off_nodes = reinterpret_cast<const char*>(cur) - ar.base;
        ok = ar.take<float>(n_nb, &w, &d.off_w);
        ok = ar.take<int32_t>(n_nb, &nbz, &d.off_nbz) && ok;
        ok = ar.take<int32_t>(n, &self_rows, &d.off_self_rows) && ok;
        if (!ok) PS_NOSPACE(n_nb * 16);
        // raw neighbour ids: a temporary at the arena's current end (overwritten by what is carved next)
        const int64_t save = ar.off;
        ok = ar.take<int32_t>(n_nb, &nb, &off_nb_tmp);
        if (!ok) PS_NOSPACE(n_nb * 12);
        if (nsrc.graph != nullptr) {  // fresh neighbourhoods of this layer's targets: top-T visit counts of n_hops walk steps each
            rc = ps_walk_topt(nsrc.graph, cur, n, nsrc.n_hops, nsrc.alpha, 0, T, nsrc.seed, nullptr, nullptr, nb, w, nullptr, stream_);
            if (rc != PS_OK) return rc;
        } else {
            lookup_kernel<<<blocks_for(n_nb), 256, 0, stream>>>(cur, n, tab_nodes, tab_w, Tp, T, nb, w);
            PS_LAUNCH_CHECK();
        }
        PS_CUDA_CHECK(cudaMemsetAsync(flag, 0, static_cast<size_t>(n_ids + 1) * sizeof(int32_t), stream));
        mark_kernel<<<blocks_for(n_nb), 256, 0, stream>>>(nb, n_nb, cur, n, with_self ? 1 : 0, n_ids, flag);
        PS_LAUNCH_CHECK();
        PS_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, flag, pos, static_cast<int>(n_ids + 1), stream));
        inverse_kernel<<<blocks_for(n_nb), 256, 0, stream>>>(pos, nb, n_nb, cur, n, with_self ? 1 : 0, n_ids, nbz, self_rows);
        PS_LAUNCH_CHECK();
        if (!with_self) {
            narrow_ids_kernel<<<blocks_for(n), 256, 0, stream>>>(cur, n, self_rows);  // layer 0 reads the feature table by node id
            PS_LAUNCH_CHECK();
        }
        PS_CUDA_CHECK(cudaMemcpyAsync(&h_count[0], pos + n_ids, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        PS_CUDA_CHECK(cudaStreamSynchronize(stream));
        const int64_t nz = h_count[0];
        d.nz = nz;
        ar.off = save;  // the raw ids are dead once nbz exists
        int64_t* uniq64 = nullptr; int32_t* uniq32 = nullptr;
        d.off_zrows = -1;
        if (with_self) {
            int64_t off_next = 0;
            if (!ar.take<int64_t>(nz, &uniq64, &off_next)) PS_NOSPACE(nz * 64);
        } else {
            if (!ar.take<int32_t>(nz, &uniq32, &d.off_zrows)) PS_NOSPACE(nz * 64);
        }
        compact_kernel<<<blocks_for(n_ids), 256, 0, stream>>>(flag, pos, n_ids, uniq64, uniq32);
        PS_LAUNCH_CHECK();
        d.off_seg_off = d.off_pair_q = d.off_chunk_off = d.off_chunk_row = -1;
        if (need_backward) {
            const int64_t max_chunks = n_nb / PS_AGG_BWD_CHUNK + nz;
            int32_t *pair_q = nullptr, *seg_off = nullptr, *chunk_off = nullptr, *chunk_row = nullptr;
            ok = ar.take<int32_t>(n_nb, &pair_q, &d.off_pair_q);
            ok = ar.take<int32_t>(nz + 1, &seg_off, &d.off_seg_off) && ok;
            ok = ar.take<int32_t>(nz + 1, &chunk_off, &d.off_chunk_off) && ok;
            ok = ar.take<int32_t>(max_chunks > 1 ? max_chunks : 1, &chunk_row, &d.off_chunk_row) && ok;
            if (!ok) PS_NOSPACE(0);
            rc = ps_plan_transpose(nbz, n_nb, nz, PS_AGG_BWD_CHUNK, pair_q, seg_off, chunk_off, chunk_row, max_chunks, stream_);
            if (rc != PS_OK) return rc;
        }
        cur = uniq64;
        n = nz;
    }
#undef PS_NOSPACE
    out->bytes_used = ar.off;
    out->bytes_needed = ar.off;
    return PS_OK;
}

extern "C" int ps_prepare_plan(const int64_t* batch, int64_t B, const int32_t* tab_nodes, const float* tab_w, int64_t n_ids, int Tp,
                               int T, int n_layers, int need_backward, void* arena_, int64_t arena_bytes, ps_plan_desc* out,
                               ps_stream_t stream_) {
    PS_REQUIRE(tab_nodes && tab_w, "null pointer");
    const NeighbourSource nsrc{tab_nodes, tab_w, Tp, nullptr, 0, 0.0, 0ull};
    return prepare_plan_impl(batch, B, nsrc, n_ids, T, n_layers, need_backward, arena_, arena_bytes, out, stream_);
}

extern "C" int ps_prepare_plan_online(const int64_t* batch, int64_t B, const ps_graph_t* graph, int64_t n_items, int n_hops,
                                      double alpha, uint64_t seed, int T, int n_layers, int need_backward, void* arena_,
                                      int64_t arena_bytes, ps_plan_desc* out, ps_stream_t stream_) {
    PS_REQUIRE(graph != nullptr, "null pointer");
    const NeighbourSource nsrc{nullptr, nullptr, 0, graph, n_hops, alpha, seed};
    return prepare_plan_impl(batch, B, nsrc, n_items, T, n_layers, need_backward, arena_, arena_bytes, out, stream_);
}
