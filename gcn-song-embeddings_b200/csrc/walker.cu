// K1 + K2: restart random walks over the bipartite track-collection CSR, fused with the
// visit-count -> top-T reduction.  Replaces the Python double loop of
// do_random_walks (reference pinsage_model.py:32-53) and the dense [n, N+C] float64
// histogram + torch.topk of sample_neighborhood[_topt] (:88-107).
//
// Design (B200): one warp per source node, 32 steps per pass, one step per lane.
//  * Walk.  The chain of a source is a sequence of i.i.d. segments that all start at the
//    source; the restart flag of step j depends only on Philox(seed, source, j).  A ballot
//    of the 32 flags gives every lane its position inside its segment; round k advances
//    the lanes at position k from their left neighbour's item (one shuffle), so a pass
//    costs max-segment-length rounds (about 3 at alpha = 0.85) and every lane computes
//    exactly one Philox block.  A segment that crosses the pass boundary is carried in a
//    register.  The source's own adjacency row bounds are loaded once.
//  * Histogram.  Visited ids go into a per-warp open-addressed hash table in shared
//    memory (id -> 16-bit count); the dense row of the reference is never materialised
//    and the trace never goes to HBM (unless the caller asks for it).
//  * Top-T.  The table is compacted in place; a warp-wide MSB-first radix select over the
//    48-bit key (count, ~id) finds the T-th largest entry in <= 6 passes; only the
//    selected <= T entries are sorted (bitonic, 64-bit keys) -> canonical order
//    (count desc, id asc); weight = count / n_hops in IEEE double.
// HBM traffic per step is the two CSR hops (indptr pair + one neighbour id each).
#include "common.cuh"
#include "../../include/pinsage_b200.h"

struct ps_graph {
    const int64_t* indptr;
    const int32_t* indices;
    int64_t n_tracks, n_cols, n_entries;
    uint32_t* indptr32;  // owned compact copy of indptr (n_entries < 2^32), halves the bytes per hop
    bool use32;          // ps_graph_use_indptr32
};

namespace {

constexpr int kMaxWarpsPerCta = 8;
constexpr uint32_t kEmpty = 0xFFFFFFFFu;
constexpr int kHistBins = 256;
constexpr int kChunks = 4;  // 32-step chunks walked per pass (independent loads in flight per lane)

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
template <typename PtrT>
__device__ __forceinline__ uint32_t hop(const PtrT* __restrict__ indptr, const int32_t* __restrict__ indices,
                                        uint32_t node, uint32_t x) {
    const PtrT beg = __ldg(indptr + node);
    const PtrT end = __ldg(indptr + node + 1);
    const uint32_t deg = static_cast<uint32_t>(end - beg);
    if (deg == 0) return node;  // unreachable after ps_graph_create's degree check
    return static_cast<uint32_t>(__ldg(indices + beg + __umulhi(x, deg)));
}

// warp-cooperative bitonic sort (descending) of a[0..P) in shared memory, P a power of two >= 32
__device__ __forceinline__ void warp_bitonic_sort_desc(uint64_t* a, int P, int lane) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (P >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool down = (i & k) == 0;
                const uint64_t x = a[i], y = a[p];
                const uint64_t mx = x > y ? x : y, mn = x > y ? y : x;
                a[i] = down ? mx : mn;
                a[p] = down ? mn : mx;
            }
            __syncwarp();
        }
    }
}

__device__ __forceinline__ uint64_t entry_key(const uint32_t* ids, const uint16_t* cnt, int i) {
    return (static_cast<uint64_t>(cnt[i]) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - ids[i]);
}

// Per-warp shared memory (32-bit words): keys[cap] | counts[cap/2] (two 16-bit counters per
// word) | optional extra; hist (256 words) and sel (2*Tp2 words) live at hist_off / sel_off,
// either in the dead tail of keys (after the in-place compaction) or in the extra region.
// kFromTrace: read the steps from a caller-supplied trace instead of walking (parity hook)
template <typename PtrT, bool kFromTrace>
__global__ void __launch_bounds__(kMaxWarpsPerCta * 32)
walk_topt_kernel(const PtrT* __restrict__ indptr, const int32_t* __restrict__ indices,
                 const int64_t* __restrict__ sources, const int64_t* __restrict__ in_trace,
                 int64_t n, int n_hops, int cap, int hash_shift, int per_warp_words, int hist_off, int sel_off, int Tp2,
                 uint64_t restart_thr, int fixed_len, int T, uint32_t k0, uint32_t k1,
                 int64_t* __restrict__ out_nodes64, double* __restrict__ out_w64,
                 int32_t* __restrict__ out_nodes32, float* __restrict__ out_w32,
                 int32_t* __restrict__ out_trace) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint32_t* keys = smem + static_cast<size_t>(warp) * per_warp_words;
    uint32_t* cnt32 = keys + cap;
    uint16_t* cnt16 = reinterpret_cast<uint16_t*>(cnt32);
    uint32_t* hist = keys + hist_off;
    uint64_t* sel = reinterpret_cast<uint64_t*>(keys + sel_off);
    const uint32_t cap_mask = static_cast<uint32_t>(cap - 1);

    const int warps_per_cta = blockDim.x >> 5;
    for (int64_t s = static_cast<int64_t>(blockIdx.x) * warps_per_cta + warp; s < n;
         s += static_cast<int64_t>(gridDim.x) * warps_per_cta) {
        const uint32_t src = static_cast<uint32_t>(sources[s]);

        {   // empty table
            uint4* k4 = reinterpret_cast<uint4*>(keys);
            for (int i = lane; i < (cap >> 2); i += 32) k4[i] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
            uint4* c4 = reinterpret_cast<uint4*>(cnt32);
            for (int i = lane; i < (cap >> 3); i += 32) c4[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncwarp();

        PtrT sbeg = 0;   // the source's own row bounds: 85 % of the steps start there
        uint32_t sdeg = 0;
        if (!kFromTrace) {
            sbeg = __ldg(indptr + src);
            sdeg = static_cast<uint32_t>(__ldg(indptr + src + 1) - sbeg);
        }
        uint32_t carry_item = src;  // where step 0 of the next pass starts from

        for (int base = 0; base < n_hops; base += 32 * kChunks) {
            uint32_t item[kChunks], x0[kChunks], x1[kChunks], rm[kChunks];
            int pos[kChunks];
            bool valid[kChunks];
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                const int j = base + c * 32 + lane;
                valid[c] = j < n_hops;
                item[c] = src;
                if (kFromTrace) {
                    if (valid[c]) item[c] = static_cast<uint32_t>(in_trace[s * n_hops + j]);
                } else {
                    const Philox d = philox4x32_10(static_cast<uint32_t>(j), src, 0u, 0u, k0, k1);
                    x0[c] = d.x0; x1[c] = d.x1;
                    const bool r = !valid[c] || (fixed_len > 0 ? ((j + 1) % fixed_len == 0)
                                                               : (static_cast<uint64_t>(d.x2) < restart_thr));
                    rm[c] = __ballot_sync(0xffffffffu, r);
                }
            }
            if (!kFromTrace) {
                // position of every step inside its segment: step i starts one iff i == 0 or step i-1 restarted
                int last_start = 0, maxpos = 0;
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    const uint32_t starts = (rm[c] << 1) | (c == 0 ? 1u : (rm[c - 1] >> 31));
                    const uint32_t le = starts & (0xFFFFFFFFu >> (31 - lane));
                    pos[c] = c * 32 + lane - (le ? c * 32 + 31 - __clz(le) : last_start);
                    if (starts) last_start = c * 32 + 31 - __clz(starts);
                    if (valid[c]) maxpos = max(maxpos, pos[c]);
                }
                maxpos = __reduce_max_sync(0xffffffffu, maxpos);
                if (lane == 0) item[0] = carry_item;
                for (int k = 0; k <= maxpos; ++k) {
                    // round k advances the steps at position k from their predecessor's item
                    // the loads of the kChunks independent steps of a lane are issued stage by stage so they overlap
                    uint32_t cur[kChunks], deg[kChunks];
                    PtrT beg[kChunks];
                    bool act[kChunks];
#pragma unroll
                    for (int c = 0; c < kChunks; ++c) {
                        uint32_t prev = __shfl_up_sync(0xffffffffu, item[c], 1);
                        if (c > 0) {
                            const uint32_t tail = __shfl_sync(0xffffffffu, item[c - 1], 31);
                            if (lane == 0) prev = tail;
                        }
                        act[c] = valid[c] && pos[c] == k;
                        cur[c] = k == 0 ? item[c] : prev;
                    }
#pragma unroll
                    for (int c = 0; c < kChunks; ++c) {
                        beg[c] = sbeg; deg[c] = sdeg;  // 85 % of the steps start at the source: its row bounds are in registers
                        if (act[c] && cur[c] != src) {
                            beg[c] = __ldg(indptr + cur[c]);
                            deg[c] = static_cast<uint32_t>(__ldg(indptr + cur[c] + 1) - beg[c]);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < kChunks; ++c)
                        if (act[c] && deg[c] != 0) cur[c] = static_cast<uint32_t>(__ldg(indices + beg[c] + __umulhi(x0[c], deg[c])));
#pragma unroll
                    for (int c = 0; c < kChunks; ++c) {
                        deg[c] = 0;
                        if (act[c]) {
                            beg[c] = __ldg(indptr + cur[c]);
                            deg[c] = static_cast<uint32_t>(__ldg(indptr + cur[c] + 1) - beg[c]);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < kChunks; ++c) {
                        if (act[c] && deg[c] != 0) cur[c] = static_cast<uint32_t>(__ldg(indices + beg[c] + __umulhi(x1[c], deg[c])));
                        if (act[c]) item[c] = cur[c];
                    }
                }
                carry_item = (rm[kChunks - 1] >> 31) ? src : __shfl_sync(0xffffffffu, item[kChunks - 1], 31);
            }
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                const int j = base + c * 32 + lane;
                if (out_trace != nullptr && valid[c]) out_trace[s * n_hops + j] = static_cast<int32_t>(item[c]);
                if (valid[c] && item[c] != src) {  // the self entry is zeroed by the reference (pinsage_model.py:98-99)
                    uint32_t slot = (item[c] * 0x9E3779B1u) >> hash_shift;
                    while (true) {
                        const uint32_t old = atomicCAS(keys + slot, kEmpty, item[c]);
                        if (old == kEmpty || old == item[c]) {
                            atomicAdd(cnt32 + (slot >> 1), 1u << ((slot & 1u) << 4));
                            break;
                        }
                        slot = (slot + 1) & cap_mask;
                    }
                }
            }
        }
        __syncwarp();

        // in-place compaction of the table: ids -> keys[0, D), counts -> cnt16[0, D)
        int D = 0;
        for (int b = 0; b < cap; b += 32) {
            const uint32_t id = keys[b + lane];
            const uint32_t c = cnt16[b + lane];
            const uint32_t m = __ballot_sync(0xffffffffu, id != kEmpty);
            __syncwarp();
            if (id != kEmpty) {
                const int p = D + __popc(m & ((1u << lane) - 1u));
                keys[p] = id;
                cnt16[p] = static_cast<uint16_t>(c);
            }
            D += __popc(m);
            __syncwarp();
        }

        // radix select, MSB first, of the T-th largest (count, ~id): digits 0-1 are the count bytes, 2-5 the
        // bytes of ~id among the entries whose count equals the boundary count
        uint64_t thr_key = 0;  // D <= T: everything is selected
        if (D > T) {
            uint32_t cpre = 0, ipre = 0;  // fixed high bytes of the count / of ~id
            int R = T;
            for (int p = n_hops >= 256 ? 0 : 1; p < 6; ++p) {
                for (int b = lane; b < kHistBins; b += 32) hist[b] = 0;
                __syncwarp();
                const int sh = p < 2 ? 8 * (1 - p) : 8 * (5 - p);
                for (int i = lane; i < D; i += 32) {
                    const uint32_t c = cnt16[i];
                    const uint32_t nid = ~keys[i];
                    bool match;
                    uint32_t v;
                    if (p < 2) { v = c; match = p == 0 || (c >> 8) == cpre; }
                    else { v = nid; match = c == cpre && (p == 2 || (nid >> (sh + 8)) == ipre); }
                    if (match) atomicAdd(hist + ((v >> sh) & 255u), 1u);
                }
                __syncwarp();
                int loc[8], sum = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) { loc[q] = static_cast<int>(hist[255 - 8 * lane - q]); sum += loc[q]; }
                int incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                const int excl = incl - sum;
                const bool mine = excl < R && R <= incl;
                int bsel = 0, above = 0;
                if (mine) {
                    int run = excl;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (run < R && R <= run + loc[q]) { bsel = 255 - 8 * lane - q; above = run; }
                        run += loc[q];
                    }
                }
                const int owner = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
                bsel = __shfl_sync(0xffffffffu, bsel, owner);
                above = __shfl_sync(0xffffffffu, above, owner);
                R -= above;
                if (p < 2) cpre = (cpre << 8) | static_cast<uint32_t>(bsel);
                else ipre = (ipre << 8) | static_cast<uint32_t>(bsel);
                __syncwarp();
            }
            thr_key = (static_cast<uint64_t>(cpre) << 32) | ipre;
        }

        // gather the selected entries, sort them (count desc, id asc)
        int n_sel = 0;
        for (int b = 0; b < D; b += 32) {
            const int i = b + lane;
            const uint64_t key = i < D ? entry_key(keys, cnt16, i) : 0ull;
            const bool take = i < D && key >= thr_key;
            const uint32_t m = __ballot_sync(0xffffffffu, take);
            if (take) sel[n_sel + __popc(m & ((1u << lane) - 1u))] = key;
            n_sel += __popc(m);
        }
        for (int i = n_sel + lane; i < Tp2; i += 32) sel[i] = 0ull;
        __syncwarp();
        warp_bitonic_sort_desc(sel, Tp2, lane);

        for (int t = lane; t < T; t += 32) {
            const uint64_t key = sel[t];
            uint32_t node = src;
            uint32_t count = 0;
            if (key != 0ull) {
                count = static_cast<uint32_t>(key >> 32);
                node = 0xFFFFFFFFu - static_cast<uint32_t>(key);
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

__global__ void narrow_indptr_kernel(const int64_t* __restrict__ indptr, uint32_t* __restrict__ out, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = static_cast<uint32_t>(indptr[i]);
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

template <typename PtrT, bool kFromTrace>
int launch_walk(const PtrT* indptr, const int32_t* indices, const int64_t* sources, const int64_t* trace,
                int64_t n, int n_hops, double alpha, int fixed_len, int T, uint64_t seed,
                int64_t* on64, double* ow64, int32_t* on32, float* ow32, int32_t* otrace, cudaStream_t stream) {
    PS_REQUIRE(n >= 0 && n_hops > 0 && n_hops <= 16384, "n_hops must be in [1, 16384] (got %d)", n_hops);
    PS_REQUIRE(T > 0 && T <= 8192, "T must be in [1, 8192]");
    PS_REQUIRE(alpha >= 0.0 && alpha <= 1.0, "alpha must be in [0, 1]");
    if (n == 0) return PS_OK;
    const int cap = next_pow2(2 * n_hops < 64 ? 64 : 2 * n_hops);  // load factor <= 1/2
    int log2cap = 0;
    while ((1 << log2cap) < cap) ++log2cap;
    const int Tp2 = next_pow2(T);
    const int need = kHistBins + 2 * Tp2;  // hist + 64-bit sort keys, in words
    int per_warp_words = cap + cap / 2, hist_off;
    if (cap - n_hops >= need) {
        hist_off = cap - need;  // dead tail of the key array once the table is compacted
    } else {
        hist_off = per_warp_words;
        per_warp_words += need;
    }
    const size_t per_warp = static_cast<size_t>(per_warp_words) * sizeof(uint32_t);
    int warps = static_cast<int>((200 * 1024) / per_warp);
    if (warps > kMaxWarpsPerCta) warps = kMaxWarpsPerCta;
    PS_REQUIRE(warps >= 1, "n_hops=%d / T=%d too large for the shared-memory visit table", n_hops, T);
    const size_t smem = static_cast<size_t>(warps) * per_warp;
    auto kern = walk_topt_kernel<PtrT, kFromTrace>;
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
        indptr, indices, sources, trace, n, n_hops, cap, 32 - log2cap, per_warp_words, hist_off, hist_off + kHistBins, Tp2,
        thr, fixed_len, T, static_cast<uint32_t>(seed & 0xFFFFFFFFull), static_cast<uint32_t>(seed >> 32),
        on64, ow64, on32, ow32, otrace);
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
    ps_graph* g = new ps_graph{indptr, indices, n_tracks, n_cols, n_entries, nullptr, true};
    if (n_entries < (1ll << 32)) {  // 4-byte row offsets for the walker
        if (cudaMalloc(&g->indptr32, static_cast<size_t>(n_nodes + 1) * sizeof(uint32_t)) != cudaSuccess) {
            g->indptr32 = nullptr;
            (void)cudaGetLastError();
        } else {
            narrow_indptr_kernel<<<static_cast<unsigned>(ps_ceil_div(n_nodes + 1, 256)), 256, 0, stream>>>(indptr, g->indptr32, n_nodes + 1);
            PS_LAUNCH_CHECK();
            PS_CUDA_CHECK(cudaStreamSynchronize(stream));
        }
    }
    *out = g;
    return PS_OK;
}

extern "C" int ps_graph_use_indptr32(ps_graph_t* g, int on) {
    PS_REQUIRE(g != nullptr, "null pointer");
    const int old = g->use32 ? 1 : 0;
    g->use32 = on != 0;
    return old;
}

extern "C" int ps_graph_destroy(ps_graph_t* g) {
    if (g != nullptr && g->indptr32 != nullptr) cudaFree(g->indptr32);
    delete g;
    return PS_OK;
}

extern "C" int ps_walk_topt(const ps_graph_t* g, const int64_t* sources, int64_t n, int n_hops, double alpha,
                            int fixed_len, int T, uint64_t seed, int64_t* out_nodes_i64, double* out_w_f64,
                            int32_t* out_nodes_i32, float* out_w_f32, int32_t* out_trace, ps_stream_t stream) {
    PS_REQUIRE(g != nullptr && (sources != nullptr || n == 0), "null pointer");
    PS_REQUIRE(fixed_len >= 0, "fixed_len must be >= 0");
    if (g->indptr32 != nullptr && g->use32)
        return launch_walk<uint32_t, false>(g->indptr32, g->indices, sources, nullptr, n, n_hops, alpha, fixed_len, T, seed,
                                            out_nodes_i64, out_w_f64, out_nodes_i32, out_w_f32, out_trace,
                                            static_cast<cudaStream_t>(stream));
    return launch_walk<int64_t, false>(g->indptr, g->indices, sources, nullptr, n, n_hops, alpha, fixed_len, T, seed,
                                       out_nodes_i64, out_w_f64, out_nodes_i32, out_w_f32, out_trace,
                                       static_cast<cudaStream_t>(stream));
}

extern "C" int ps_trace_topt(const int64_t* trace, const int64_t* sources, int64_t n, int n_hops, int T,
                             int64_t* out_nodes_i64, double* out_w_f64, int32_t* out_nodes_i32, float* out_w_f32,
                             ps_stream_t stream) {
    PS_REQUIRE((trace != nullptr && sources != nullptr) || n == 0, "null pointer");
    return launch_walk<uint32_t, true>(nullptr, nullptr, sources, trace, n, n_hops, 0.0, 0, T, 0ull,
                             out_nodes_i64, out_w_f64, out_nodes_i32, out_w_f32, nullptr,
                             static_cast<cudaStream_t>(stream));
}
