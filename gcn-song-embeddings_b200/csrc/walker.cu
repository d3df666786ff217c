// K1 + K2: restart random walks over the bipartite track-collection CSR, fused with the
// visit-count -> top-T reduction.  Replaces the Python double loop of
// do_random_walks (reference pinsage_model.py:32-53) and the dense [n, N+C] float64
// histogram + torch.topk of sample_neighborhood[_topt] (:88-107).
//
// Design (B200): one warp per source node; draws keyed by Philox4x32-10(counter = (step, source, 0, 0), key = seed), so
// the result does not depend on the launch shape and a CPU restatement reproduces it bit for bit.
//  * Walk (walk_source).  The chain of a source is a sequence of i.i.d. segments that all start at the source; whether
//    step j ends its segment depends only on Philox(j, source).  Steps that START a segment (85 % at alpha = 0.85) are
//    walked in place, 32 * U at a time, from the source's row bounds held in registers; a step whose successor continues
//    the segment appends (j + 1, item) to a list in shared memory, and list rounds advance the pending continuations 32
//    at a time.  Every lane computes one Philox block per step it walks.
//  * Top-T, two kernels with identical results:
//      walk_sort_kernel (n_hops <= 512, T <= 256): the trace stays in shared memory (4 B per step), is sorted in registers
//        (lane-major bitonic network), run lengths of the sorted trace are the visit counts, the boundary count class
//        comes from packed tallies and the "R smallest ids of that class" from prefix ballots;
//      walk_topt_kernel (everything else): per-warp open-addressed hash table (id -> 16-bit count), in-place
//        compaction, class tallies, MSB-first radix passes over the significant id bits of the boundary class.
//    Both end with the <= T selected keys sorted (count desc, id asc) and weight = count / n_hops in IEEE double; the
//    dense row of the reference is never materialised and the trace never goes to HBM (unless the caller asks for it).
// HBM traffic per step is the two CSR hops (row bounds + one neighbour id each): in practice one random 64-byte DRAM
// burst per step, which is what bounds the kernel.
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
#ifndef PS_WALK_UNROLL
#define PS_WALK_UNROLL 4
#endif
constexpr int kWalkUnroll = PS_WALK_UNROLL;      // 32-step chunks walked together (independent loads in flight per lane)
constexpr int kListCap = 64 + 32 * kWalkUnroll;  // capacity of the per-warp continuation list
#ifndef PS_SORT_UNROLL
#define PS_SORT_UNROLL 2
#endif
constexpr int kSortUnroll = PS_SORT_UNROLL;      // the same two constants for the sort-based kernel
constexpr int kSortListCap = 64 + 32 * kSortUnroll;
#ifndef PS_SORT_CTAS
#define PS_SORT_CTAS 4
#endif
constexpr int kSortCtasPerSm = PS_SORT_CTAS;     // register budget of the sort-based kernel: 65536 / (256 * CTAs) per thread

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

// warp-cooperative bitonic sort (descending) of a[0..P) in shared memory, P a power of two >= 32 (T > 256)
template <typename K>
__device__ __forceinline__ void warp_bitonic_sort_desc(K* a, int P, int lane) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (P >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool down = (i & k) == 0;
                const K x = a[i], y = a[p];
                const K mx = x > y ? x : y, mn = x > y ? y : x;
                a[i] = down ? mx : mn;
                a[p] = down ? mn : mx;
            }
            __syncwarp();
        }
    }
}

// Bitonic sort (descending) of 32*R keys, R per lane in registers; on return the key of rank t sits in lane t / R,
// register t % R (lane-major), whatever the initial placement was.  Lane-major order keeps every stage whose stride is
// below R inside the lane (one min + one max per pair, no shuffle), and the "flip" formulation (the first stage of a
// phase pairs e with e ^ (k - 1), the others e with e ^ j) makes the lower index keep the maximum in every stage, so
// a shuffle stage costs one predicate for the whole stage, then SHFL + min/max per key.  For R = 16 (512 keys): ~1200
// instructions against ~1900 for the row-major network with per-key direction tests.
template <typename K, int R>
__device__ __forceinline__ void warp_sort_desc_lane_major(K (&a)[R], int lane) {
    constexpr uint32_t kFull = 0xffffffffu;
    auto ce = [](K& hi, K& lo) {  // compare-exchange inside the lane
        const K x = hi, y = lo;
        hi = x > y ? x : y;
        lo = x > y ? y : x;
    };
    // phases within a lane: sorted runs of k = 2 .. R keys
#pragma unroll
    for (int k = 2; k <= R; k <<= 1) {
#pragma unroll
        for (int r = 0; r < R; ++r)
            if ((r ^ (k - 1)) > r) ce(a[r], a[r ^ (k - 1)]);
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1)
#pragma unroll
            for (int r = 0; r < R; ++r)
                if ((r ^ j) > r) ce(a[r], a[r ^ j]);
    }
    // phases across lanes: runs of kl = 2 .. 32 lanes
#pragma unroll 1
    for (int kl = 2; kl <= 32; kl <<= 1) {
        {   // flip: (lane, r) with (lane ^ (kl - 1), R - 1 - r)
            const bool keep_max = (lane & (kl >> 1)) == 0;
#pragma unroll
            for (int r = 0; r < (R + 1) / 2; ++r) {
                const K y0 = __shfl_xor_sync(kFull, a[R - 1 - r], kl - 1);
                const K y1 = __shfl_xor_sync(kFull, a[r], kl - 1);
                a[r] = keep_max ? (a[r] > y0 ? a[r] : y0) : (a[r] > y0 ? y0 : a[r]);
                if (R - 1 - r != r) a[R - 1 - r] = keep_max ? (a[R - 1 - r] > y1 ? a[R - 1 - r] : y1) : (a[R - 1 - r] > y1 ? y1 : a[R - 1 - r]);
            }
        }
#pragma unroll 1
        for (int jl = kl >> 2; jl > 0; jl >>= 1) {  // strides of jl lanes
            const bool keep_max = (lane & jl) == 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const K y = __shfl_xor_sync(kFull, a[r], jl);
                a[r] = keep_max ? (a[r] > y ? a[r] : y) : (a[r] > y ? y : a[r]);
            }
        }
#pragma unroll
        for (int j = R >> 1; j > 0; j >>= 1)  // strides inside the lane
#pragma unroll
            for (int r = 0; r < R; ++r)
                if ((r ^ j) > r) ce(a[r], a[r ^ j]);
    }
}

// one visit of `item` in the per-warp table (id -> 16-bit count, two counters per word)
__device__ __forceinline__ void visit(uint32_t* keys, uint32_t* cnt32, uint32_t cap_mask, int hash_shift, uint32_t item) {
    uint32_t slot = (item * 0x9E3779B1u) >> hash_shift;
    while (true) {
        const uint32_t old = atomicCAS(keys + slot, kEmpty, item);
        if (old == kEmpty || old == item) {
            atomicAdd(cnt32 + (slot >> 1), 1u << ((slot & 1u) << 4));
            break;
        }
        slot = (slot + 1) & cap_mask;
    }
}

// One MSB-first radix-select pass over the compacted entries: 256-bin histogram of the digit (v >> sh) & 255 of the
// entries that match (kId: v = ~id among the entries whose count is `cmatch`; else v = count), bins scanned from the
// highest down.  Returns the bin that holds the R-th largest matching entry (bsel), R reduced by the population above
// that bin, and the bin's population q.
template <bool kId>
__device__ __forceinline__ void radix_pass(uint32_t* hist, const uint32_t* keys, const uint16_t* cnt16, int D, int lane,
                                           uint32_t cmatch, bool has_pre, uint32_t pre, int pre_sh, int sh,
                                           int& R, uint32_t& bsel_out, int& q_out) {
    for (int b = lane; b < kHistBins; b += 32) hist[b] = 0;
    __syncwarp();
    for (int i = lane; i < D; i += 32) {
        const uint32_t c = cnt16[i];
        const uint32_t v = kId ? ~keys[i] : c;
        const bool match = (!kId || c == cmatch) && (!has_pre || (v >> pre_sh) == pre);
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
    int bsel = 0, above = 0, pop = 0;
    if (mine) {
        int run = excl;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (run < R && R <= run + loc[q]) { bsel = 255 - 8 * lane - q; above = run; pop = loc[q]; }
            run += loc[q];
        }
    }
    const int owner = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
    bsel_out = static_cast<uint32_t>(__shfl_sync(0xffffffffu, bsel, owner));
    R -= __shfl_sync(0xffffffffu, above, owner);
    q_out = __shfl_sync(0xffffffffu, pop, owner);
    __syncwarp();
}

// sel[0, n_sel) (K keys: uint32_t = count << id_bits | low id_bits of ~id, uint64_t = count << 32 | ~id) -> canonical
// order (count desc, id asc) -> outputs.  R keys per lane in registers when T <= 256, else (R = 0) the network runs in
// shared memory.
template <typename K, int R>
__device__ __forceinline__ void sort_emit(K* sel, int n_sel, int Tp2, int id_bits, int lane, uint32_t src, int64_t s, int T, int n_hops,
                                          int64_t* __restrict__ out_nodes64, double* __restrict__ out_w64,
                                          int32_t* __restrict__ out_nodes32, float* __restrict__ out_w32) {
    const uint32_t id_mask = id_bits >= 32 ? 0xFFFFFFFFu : ((1u << id_bits) - 1u);
    for (int i = n_sel + lane; i < Tp2; i += 32) sel[i] = 0;
    __syncwarp();
    auto emit = [&](int t, K key) {
        if (t >= T) return;
        uint32_t node = src, count = 0;
        if (key != 0) {
            if (sizeof(K) == 4) {
                count = static_cast<uint32_t>(key) >> id_bits;
                node = ~static_cast<uint32_t>(key) & id_mask;
            } else {
                count = static_cast<uint32_t>(static_cast<uint64_t>(key) >> 32);
                node = ~static_cast<uint32_t>(key);
            }
        }
        const int64_t o = s * T + t;
        const double w = static_cast<double>(count) / static_cast<double>(n_hops);
        if (out_nodes64) out_nodes64[o] = static_cast<int64_t>(node);
        if (out_w64) out_w64[o] = w;
        if (out_nodes32) out_nodes32[o] = static_cast<int32_t>(node);
        if (out_w32) out_w32[o] = static_cast<float>(w);
    };
    if (R > 0) {
        constexpr int RR = R > 0 ? R : 1;
        K a[RR];
#pragma unroll
        for (int r = 0; r < RR; ++r) a[r] = sel[r * 32 + lane];
        __syncwarp();
        warp_sort_desc_lane_major<K, RR>(a, lane);
#pragma unroll
        for (int r = 0; r < RR; ++r) sel[lane * RR + r] = a[r];  // rank t = lane * RR + r
        __syncwarp();
    } else {
        warp_bitonic_sort_desc<K>(sel, Tp2, lane);
    }
#pragma unroll 1
    for (int t = lane; t < T; t += 32) emit(t, sel[t]);
}

// hash-table path: gather the entries at or above the threshold key, then sort_emit
template <typename K, int R>
__device__ __forceinline__ void emit_sorted(const uint32_t* keys, const uint16_t* cnt16, int D, uint64_t thr_key, K* sel, int Tp2,
                                            int id_bits, int lane, uint32_t src, int64_t s, int T, int n_hops,
                                            int64_t* __restrict__ out_nodes64, double* __restrict__ out_w64,
                                            int32_t* __restrict__ out_nodes32, float* __restrict__ out_w32) {
    const uint32_t id_mask = id_bits >= 32 ? 0xFFFFFFFFu : ((1u << id_bits) - 1u);
    int n_sel = 0;
    for (int b = 0; b < D; b += 32) {
        const int i = b + lane;
        const uint32_t c = i < D ? cnt16[i] : 0u;
        const uint32_t v = i < D ? ~keys[i] : 0u;
        const uint64_t key = (static_cast<uint64_t>(c) << 32) | v;
        const bool take = i < D && key >= thr_key;
        const uint32_t m = __ballot_sync(0xffffffffu, take);
        if (take) sel[n_sel + __popc(m & ((1u << lane) - 1u))] = sizeof(K) == 4 ? static_cast<K>((c << id_bits) | (v & id_mask)) : static_cast<K>(key);
        n_sel += __popc(m);
    }
    sort_emit<K, R>(sel, n_sel, Tp2, id_bits, lane, src, s, T, n_hops, out_nodes64, out_w64, out_nodes32, out_w32);
}

// The walk of one source by one warp.  The chain of a source is a sequence of i.i.d. segments that all start at the
// source; whether step j ends its segment depends only on Philox(j, source).  Phase 0 walks, 32*U steps at a time, every
// step that STARTS a segment (85 % of them at alpha = 0.85: both hops leave from the source, whose row bounds sit in
// registers); a step whose successor continues the segment appends (j + 1, item) to a small list in shared memory, and
// list rounds advance all pending continuations by one step each, 32 at a time (dense lanes instead of the ~15 %
// occupancy of walking them in place).  record(j, item) receives every step exactly once, in no particular order.
// list_item / list_step: per-warp shared memory, 64 + 32*U entries.
template <typename PtrT, int U, typename Rec>
__device__ __forceinline__ void walk_source(const PtrT* __restrict__ indptr, const int32_t* __restrict__ indices, uint32_t src,
                                            int n_hops, uint64_t restart_thr, int fixed_len, uint32_t k0, uint32_t k1,
                                            uint32_t* list_item, uint16_t* list_step, int lane, Rec record) {
    constexpr uint32_t kFull = 0xffffffffu;
    constexpr int kCap = 64 + 32 * U;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const PtrT sbeg = __ldg(indptr + src);
    const uint32_t sdeg = static_cast<uint32_t>(__ldg(indptr + src + 1) - sbeg);
    int list_n = 0;
    auto restarts = [&](int j, uint32_t x2) {
        return fixed_len > 0 ? ((j + 1) % fixed_len == 0) : (static_cast<uint64_t>(x2) < restart_thr);
    };
    auto list_round = [&]() {  // every pending continuation advances one step; survivors are compacted in place
        __syncwarp();
        const int n_in = list_n;
        int n_out = 0;
        for (int b = 0; b < n_in; b += 32) {
            const bool has = b + lane < n_in;
            const int j = has ? list_step[b + lane] : 0;
            uint32_t cur = has ? list_item[b + lane] : 0u;
            __syncwarp();  // the batch is in registers before survivors overwrite slots <= b + 31
            bool cont = false;
            if (has) {
                const Philox d = philox4x32_10(static_cast<uint32_t>(j), src, 0u, 0u, k0, k1);
                PtrT beg = __ldg(indptr + cur);
                uint32_t deg = static_cast<uint32_t>(__ldg(indptr + cur + 1) - beg);
                if (deg != 0) cur = static_cast<uint32_t>(__ldg(indices + beg + __umulhi(d.x0, deg)));
                beg = __ldg(indptr + cur);
                deg = static_cast<uint32_t>(__ldg(indptr + cur + 1) - beg);
                if (deg != 0) cur = static_cast<uint32_t>(__ldg(indices + beg + __umulhi(d.x1, deg)));
                record(j, cur);
                cont = !restarts(j, d.x2) && j + 1 < n_hops;
            }
            const uint32_t m = __ballot_sync(kFull, cont);
            if (cont) {
                const int slot = n_out + __popc(m & lt_mask);
                list_item[slot] = cur;
                list_step[slot] = static_cast<uint16_t>(j + 1);
            }
            n_out += __popc(m);
        }
        list_n = n_out;
    };

    uint32_t prev_restart = 1u;  // warp-uniform: the step before this chunk ended its segment (step -1 did)
    for (int base = 0; base < n_hops; base += 32 * U) {
        while (list_n > kCap - 32 * U) list_round();  // room for this iteration's continuations
        uint32_t x1[U], cur[U];
        bool go[U], cont[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = base + u * 32 + lane;
            const bool valid = j < n_hops;
            const Philox d = philox4x32_10(static_cast<uint32_t>(j), src, 0u, 0u, k0, k1);
            const bool r = !valid || restarts(j, d.x2);
            const uint32_t rm = __ballot_sync(kFull, r);
            const uint32_t starts = (rm << 1) | prev_restart;  // bit l: step l of the chunk starts a segment
            prev_restart = rm >> 31;
            go[u] = valid && ((starts >> lane) & 1u);
            cont[u] = go[u] && !r && j + 1 < n_hops;
            x1[u] = d.x1;
            cur[u] = src;
            if (go[u] && sdeg != 0) cur[u] = static_cast<uint32_t>(__ldg(indices + sbeg + __umulhi(d.x0, sdeg)));
        }
        PtrT beg[U];
        uint32_t deg[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            deg[u] = 0;
            if (go[u]) {
                beg[u] = __ldg(indptr + cur[u]);
                deg[u] = static_cast<uint32_t>(__ldg(indptr + cur[u] + 1) - beg[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (go[u] && deg[u] != 0) cur[u] = static_cast<uint32_t>(__ldg(indices + beg[u] + __umulhi(x1[u], deg[u])));
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = base + u * 32 + lane;
            if (go[u]) record(j, cur[u]);
            const uint32_t m = __ballot_sync(kFull, cont[u]);
            if (cont[u]) {
                const int slot = list_n + __popc(m & lt_mask);
                list_item[slot] = cur[u];
                list_step[slot] = static_cast<uint16_t>(j + 1);
            }
            list_n += __popc(m);
        }
    }
    while (list_n > 0) list_round();
}

// Per-warp shared memory (32-bit words): keys[cap] | counts[cap/2] (two 16-bit counters per word) | extra.  The extra
// region holds the continuation list of the walk (kListCap node ids + kListCap 16-bit step numbers) and, when they do
// not fit into the dead tail of keys after the in-place compaction, hist (256 words) and sel (2*Tp2 words).
// kFromTrace: read the steps from a caller-supplied trace instead of walking (parity hook)
template <typename PtrT, bool kFromTrace>
__global__ void __launch_bounds__(kMaxWarpsPerCta * 32, 4)
walk_topt_kernel(const PtrT* __restrict__ indptr, const int32_t* __restrict__ indices,
                 const int64_t* __restrict__ sources, const int64_t* __restrict__ in_trace,
                 int64_t n, int n_hops, int cap, int hash_shift, int per_warp_words, int hist_off, int sel_off, int Tp2,
                 uint64_t restart_thr, int fixed_len, int T, uint32_t k0, uint32_t k1,
                 int64_t* __restrict__ out_nodes64, double* __restrict__ out_w64,
                 int32_t* __restrict__ out_nodes32, float* __restrict__ out_w32,
                 int32_t* __restrict__ out_trace) {
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr uint32_t kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* keys = smem + static_cast<size_t>(warp) * per_warp_words;
    uint32_t* cnt32 = keys + cap;
    uint16_t* cnt16 = reinterpret_cast<uint16_t*>(cnt32);
    uint32_t* list_item = cnt32 + (cap >> 1);
    uint16_t* list_step = reinterpret_cast<uint16_t*>(list_item + kListCap);
    uint32_t* hist = keys + hist_off;
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

        if (kFromTrace) {
            for (int j = lane; j < n_hops; j += 32) {
                const uint32_t item = static_cast<uint32_t>(in_trace[s * n_hops + j]);
                if (item != src) visit(keys, cnt32, cap_mask, hash_shift, item);  // the self entry is zeroed by the reference (pinsage_model.py:98-99)
            }
        } else {
            walk_source<PtrT, kWalkUnroll>(indptr, indices, src, n_hops, restart_thr, fixed_len, k0, k1, list_item, list_step, lane,
                                           [&](int j, uint32_t item) {
                                               if (out_trace != nullptr) out_trace[s * n_hops + j] = static_cast<int32_t>(item);
                                               if (item != src) visit(keys, cnt32, cap_mask, hash_shift, item);
                                           });
        }
        __syncwarp();

        // in-place compaction of the table: ids -> keys[0, D), counts -> cnt16[0, D)
        int D = 0;
        for (int b = 0; b < cap; b += 32) {
            const uint32_t id = keys[b + lane];
            const uint32_t c = cnt16[b + lane];
            const uint32_t m = __ballot_sync(kFull, id != kEmpty);
            __syncwarp();
            if (id != kEmpty) {
                const int p = D + __popc(m & lt_mask);
                keys[p] = id;
                cnt16[p] = static_cast<uint16_t>(c);
            }
            D += __popc(m);
            __syncwarp();
        }

        // statistics of the entries: largest id / count (the digits that matter) and the sizes of the count classes
        // 1..7 and ">= 8" (class 0), tallied in 16-bit fields (D <= n_hops <= 16384)
        uint32_t maxid = 0, maxc = 0;
        uint64_t ta = 0, tb = 0;
        for (int i = lane; i < D; i += 32) {
            const uint32_t c = cnt16[i];
            maxid = max(maxid, keys[i]);
            maxc = max(maxc, c);
            const uint32_t cls = c >= 8u ? 0u : c;
            const uint64_t inc = 1ull << ((cls & 3u) * 16u);
            if (cls & 4u) tb += inc; else ta += inc;
        }
        maxid = __reduce_max_sync(kFull, maxid);
        maxc = __reduce_max_sync(kFull, maxc);
        const int id_bits = 32 - __clz(maxid | 1u);

        // threshold (count, ~id) of the T-th largest entry: D <= T selects everything
        uint64_t thr_key = 0;
        if (D > T) {
            const uint32_t n01 = __reduce_add_sync(kFull, static_cast<uint32_t>(ta)), n23 = __reduce_add_sync(kFull, static_cast<uint32_t>(ta >> 32));
            const uint32_t n45 = __reduce_add_sync(kFull, static_cast<uint32_t>(tb)), n67 = __reduce_add_sync(kFull, static_cast<uint32_t>(tb >> 32));
            uint32_t cb = 0;  // count of the boundary class
            int R = T;        // how many entries of that class are selected ...
            int m = 0;        // ... out of m
            const int n_ge8 = static_cast<int>(n01 & 0xFFFFu);
            if (n_ge8 >= T) {  // the boundary count is 8 or more (small T): radix select over the count's 1-2 bytes
                uint32_t hi = 0, lo = 0;
                if (maxc >= 256u) radix_pass<false>(hist, keys, cnt16, D, lane, 0u, false, 0u, 0, 8, R, hi, m);
                radix_pass<false>(hist, keys, cnt16, D, lane, 0u, maxc >= 256u, hi, 8, 0, R, lo, m);
                cb = (hi << 8) | lo;
            } else {
                int above = n_ge8;
#pragma unroll
                for (int c = 7; c >= 1; --c) {
                    const uint32_t w = c >= 6 ? n67 : (c >= 4 ? n45 : (c >= 2 ? n23 : n01));
                    const int nc = static_cast<int>((c & 1) ? (w >> 16) : (w & 0xFFFFu));
                    if (cb == 0u) {
                        if (above + nc >= T) { cb = static_cast<uint32_t>(c); R = T - above; m = nc; }
                        else above += nc;
                    }
                }
            }
            // the R smallest ids of the boundary class = the R largest v = ~id: MSB-first radix passes over the
            // significant bits of the ids until the boundary bin holds at most 32 candidates, then rank those directly
            uint32_t thr_v = 0;  // R == m: the whole class
            if (R < m) {
                int sh = 32, q = m;
                uint32_t vpre = 0;
                while (q > 32 && sh > 0) {
                    const int nsh = sh == 32 ? max(id_bits - 8, 0) : max(sh - 8, 0);
                    uint32_t bsel = 0;
                    radix_pass<true>(hist, keys, cnt16, D, lane, cb, sh < 32, vpre, sh, nsh, R, bsel, q);
                    if (sh == 32) vpre = ((0xFFFFFFFFu >> (nsh + 8)) << 8) | bsel;
                    else vpre = (vpre << (sh - nsh)) | (bsel & ((1u << (sh - nsh)) - 1u));
                    sh = nsh;
                }
                uint32_t* cand = hist;
                int got = 0;
                for (int b = 0; b < D; b += 32) {
                    const int i = b + lane;
                    const uint32_t v = i < D ? ~keys[i] : 0u;
                    const bool ok = i < D && cnt16[i] == cb && (sh == 32 || (v >> sh) == vpre);
                    const uint32_t mk = __ballot_sync(kFull, ok);
                    if (ok) cand[got + __popc(mk & lt_mask)] = v;
                    got += __popc(mk);
                }
                __syncwarp();
                const uint32_t mine = lane < got ? cand[lane] : 0u;
                int rank = 0;
                for (int i = 0; i < got; ++i) rank += __shfl_sync(kFull, mine, i) > mine;
                const uint32_t who = __ballot_sync(kFull, lane < got && rank == R - 1);
                thr_v = __shfl_sync(kFull, mine, __ffs(who) - 1);
                __syncwarp();
            }
            thr_key = (static_cast<uint64_t>(cb) << 32) | thr_v;
        }

        // gather the selected entries, sort them (count desc, id asc), write the outputs
        const int count_bits = 32 - __clz(maxc | 1u);
        void* selp = keys + sel_off;
#define PS_EMIT(K, R) emit_sorted<K, R>(keys, cnt16, D, thr_key, static_cast<K*>(selp), Tp2, id_bits, lane, src, s, T, n_hops, out_nodes64, out_w64, out_nodes32, out_w32)
        if (count_bits + id_bits <= 32) {
            if (Tp2 == 32) PS_EMIT(uint32_t, 1);
            else if (Tp2 == 64) PS_EMIT(uint32_t, 2);
            else if (Tp2 == 128) PS_EMIT(uint32_t, 4);
            else if (Tp2 == 256) PS_EMIT(uint32_t, 8);
            else PS_EMIT(uint32_t, 0);
        } else {
            if (Tp2 == 32) PS_EMIT(uint64_t, 1);
            else if (Tp2 == 64) PS_EMIT(uint64_t, 2);
            else if (Tp2 == 128) PS_EMIT(uint64_t, 4);
            else if (Tp2 == 256) PS_EMIT(uint64_t, 8);
            else PS_EMIT(uint64_t, 0);
        }
#undef PS_EMIT
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Sort-based top-T (n_hops <= 1024, T <= 256): the trace of a source stays in shared memory (4 bytes per step instead
// of the 12 bytes per step of a half-empty hash table: 40-48 resident warps per SM instead of 32, and the walker is
// latency bound), is sorted in registers (bitonic network, R = n_hops/32 keys per lane: shuffles and min/max with 16-fold
// instruction-level parallelism instead of chains of shared-memory atomics), and run lengths of the sorted trace are the
// visit counts.  Equal ids end up adjacent and in ascending order, so "the R smallest ids of the boundary count class"
// is a prefix count over ballots: no radix select.
template <typename PtrT, bool kFromTrace, int R>
__global__ void __launch_bounds__(kMaxWarpsPerCta * 32, kSortCtasPerSm)
walk_sort_kernel(const PtrT* __restrict__ indptr, const int32_t* __restrict__ indices,
                 const int64_t* __restrict__ sources, const int64_t* __restrict__ in_trace,
                 int64_t n, int n_hops, int per_warp_words, int Tp2,
                 uint64_t restart_thr, int fixed_len, int T, uint32_t k0, uint32_t k1,
                 int64_t* __restrict__ out_nodes64, double* __restrict__ out_w64,
                 int32_t* __restrict__ out_nodes32, float* __restrict__ out_w32,
                 int32_t* __restrict__ out_trace) {
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr uint32_t kFull = 0xffffffffu;
    constexpr int kSteps = 32 * R;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // per warp: trace[kSteps] (v = ~id per step, 0 = nothing to count; `sorted` = the same words + 128 of padding, reused
    // once the trace is in registers) | cnt16[kSteps] | hm[R] nh[R + 1] | list_item[kSortListCap] list_step (16-bit);
    // sel (the selected keys) reuses the list's words after the walk
    uint32_t* trace = smem + static_cast<size_t>(warp) * per_warp_words;
    uint32_t* sorted = trace;
    uint16_t* cnt16 = reinterpret_cast<uint16_t*>(trace + kSteps + 128);
    uint32_t* hm = trace + kSteps + 128 + kSteps / 2;
    int* nh = reinterpret_cast<int*>(hm + R);
    uint32_t* list_item = hm + 2 * R + 2;
    uint16_t* list_step = reinterpret_cast<uint16_t*>(list_item + kSortListCap);
    uint32_t* sel = list_item;

    const int warps_per_cta = blockDim.x >> 5;
    for (int64_t s = static_cast<int64_t>(blockIdx.x) * warps_per_cta + warp; s < n;
         s += static_cast<int64_t>(gridDim.x) * warps_per_cta) {
        const uint32_t src = static_cast<uint32_t>(sources[s]);
        for (int j = n_hops + lane; j < kSteps; j += 32) trace[j] = 0u;  // steps beyond n_hops count nothing
        if (kFromTrace) {
            for (int j = lane; j < n_hops; j += 32) {
                const uint32_t item = static_cast<uint32_t>(in_trace[s * n_hops + j]);
                trace[j] = item != src ? ~item : 0u;  // the self entry is zeroed by the reference (pinsage_model.py:98-99)
            }
        } else {
            walk_source<PtrT, kSortUnroll>(indptr, indices, src, n_hops, restart_thr, fixed_len, k0, k1, list_item, list_step, lane,
                                           [&](int j, uint32_t item) {
                                               if (out_trace != nullptr) out_trace[s * n_hops + j] = static_cast<int32_t>(item);
                                               trace[j] = item != src ? ~item : 0u;
                                           });
        }
        __syncwarp();

        // ---- sort the trace: descending v = ascending id, the zeros (self visits, padding) last
        uint32_t a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = trace[r * 32 + lane];
        __syncwarp();
        warp_sort_desc_lane_major<uint32_t, R>(a, lane);

        // rank e = lane * R + r goes to sorted[e + 4 * (e / R)]: a lane stores R consecutive words, and the 4 words of padding
        // per lane keep those 128-bit stores free of bank conflicts; everything below is short rolled loops over `sorted`
        {
            uint32_t* mine = sorted + lane * (R + 4);
#pragma unroll
            for (int r = 0; r < R; r += 4) *reinterpret_cast<uint4*>(mine + r) = make_uint4(a[r], a[r + 1], a[r + 2], a[r + 3]);
        }
        __syncwarp();
        auto at = [&](int e) { return sorted[e + 4 * (e / R)]; };

        // ---- runs: element e starts a run iff it differs from element e - 1; hm[i] = head mask of elements 32 i ..
        int V = 0, D = 0;  // counted steps, distinct ids
        uint32_t carry = 0u;  // element 32 i - 1
#pragma unroll 1
        for (int i = 0; i < R; ++i) {
            const uint32_t v = at(i * 32 + lane);
            uint32_t prev = __shfl_up_sync(kFull, v, 1);
            if (lane == 0) prev = carry;
            carry = __shfl_sync(kFull, v, 31);
            const bool live = v != 0u;
            const uint32_t h = __ballot_sync(kFull, live && ((i == 0 && lane == 0) || v != prev));
            if (lane == 0) hm[i] = h;
            V += __popc(__ballot_sync(kFull, live));
            D += __popc(h);
        }
        // nh[i] = position of the first head at or after element 32 i (V if none): where a run that crosses a row ends
        if (lane == 0) {
            int next = V;
            nh[R] = next;
#pragma unroll 1
            for (int i = R - 1; i >= 0; --i) {
                const uint32_t h = hm[i];
                if (h) next = i * 32 + __ffs(h) - 1;
                nh[i] = next;
            }
        }
        __syncwarp();
        // visit counts (run lengths) at the heads, 0 elsewhere; largest count and the count-class sizes on the way
        const uint32_t minv = V > 0 ? at(V - 1) : 0xFFFFFFFFu;  // sorted descending: the last counted element
        const int id_bits = 32 - __clz(~minv | 1u);                  // largest id = ~(smallest v)
        uint32_t maxc = 1;
        uint64_t ta = 0, tb = 0;  // sizes of the count classes 1..7 and ">= 8" (class 0) in 16-bit fields
#pragma unroll 1
        for (int i = 0; i < R; ++i) {
            const uint32_t h = hm[i];
            uint32_t c = 0u;
            if ((h >> lane) & 1u) {
                const uint32_t above = (h >> lane) >> 1;
                const int next = above ? i * 32 + lane + __ffs(above) : nh[i + 1];
                c = static_cast<uint32_t>(next - (i * 32 + lane));
                maxc = max(maxc, c);
                const uint32_t cls = c >= 8u ? 0u : c;
                const uint64_t inc = 1ull << ((cls & 3u) * 16u);
                if (cls & 4u) tb += inc; else ta += inc;
            }
            cnt16[i * 32 + lane] = static_cast<uint16_t>(c);
        }
        maxc = __reduce_max_sync(kFull, maxc);
        __syncwarp();

        // ---- boundary of the top T: count cb of the class that is only partly selected, R_take of its members
        uint32_t cb = 0;
        int R_take = 0;
        if (D > T) {
            const uint32_t n01 = __reduce_add_sync(kFull, static_cast<uint32_t>(ta)), n23 = __reduce_add_sync(kFull, static_cast<uint32_t>(ta >> 32));
            const uint32_t n45 = __reduce_add_sync(kFull, static_cast<uint32_t>(tb)), n67 = __reduce_add_sync(kFull, static_cast<uint32_t>(tb >> 32));
            const int n_ge8 = static_cast<int>(n01 & 0xFFFFu);
            if (n_ge8 >= T) {
                // small T: the boundary count is 8 or more.  Largest c with #(count >= c) >= T by bisection over ballots.
                auto count_ge = [&](uint32_t c) {
                    int t = 0;
#pragma unroll 1
                    for (int i = 0; i < R; ++i) t += __popc(__ballot_sync(kFull, cnt16[i * 32 + lane] >= c));
                    return t;
                };
                uint32_t lo = 8u, hi = maxc;  // invariant: #(count >= lo) >= T
                while (lo < hi) {
                    const uint32_t mid = (lo + hi + 1u) >> 1;
                    if (count_ge(mid) >= T) lo = mid; else hi = mid - 1u;
                }
                cb = lo;
                R_take = T - count_ge(cb + 1u);
            } else {
                int above = n_ge8;
#pragma unroll
                for (int c = 7; c >= 1; --c) {
                    const uint32_t w = c >= 6 ? n67 : (c >= 4 ? n45 : (c >= 2 ? n23 : n01));
                    const int nc = static_cast<int>((c & 1) ? (w >> 16) : (w & 0xFFFFu));
                    if (cb == 0u) {
                        if (above + nc >= T) { cb = static_cast<uint32_t>(c); R_take = T - above; }
                        else above += nc;
                    }
                }
            }
        }

        // ---- selected runs -> sel (keys), in element order; of the boundary class the first R_take (smallest ids)
        const int count_bits = 32 - __clz(maxc);
        const bool pack32 = count_bits + id_bits <= 32;
        const uint32_t id_mask = id_bits >= 32 ? 0xFFFFFFFFu : ((1u << id_bits) - 1u);
        uint32_t* sel32 = sel;
        uint64_t* sel64 = reinterpret_cast<uint64_t*>(sel);
        int n_sel = 0, seen_cb = 0;
#pragma unroll 1
        for (int i = 0; i < R; ++i) {
            const uint32_t c = cnt16[i * 32 + lane];
            bool take = c != 0u;
            if (D > T) {
                const bool is_cb = c == cb;
                const uint32_t mcb = __ballot_sync(kFull, is_cb);
                take = c > cb || (is_cb && seen_cb + __popc(mcb & lt_mask) < R_take);
                seen_cb += __popc(mcb);
            }
            const uint32_t m = __ballot_sync(kFull, take);
            if (take) {
                const int slot = n_sel + __popc(m & lt_mask);
                const uint32_t v = at(i * 32 + lane);
                if (pack32) sel32[slot] = (c << id_bits) | (v & id_mask);
                else sel64[slot] = (static_cast<uint64_t>(c) << 32) | v;
            }
            n_sel += __popc(m);
        }
        // ---- canonical order (count desc, id asc) and outputs
#define PS_EMIT(K, RR) sort_emit<K, RR>(reinterpret_cast<K*>(sel), n_sel, Tp2, id_bits, lane, src, s, T, n_hops, out_nodes64, out_w64, out_nodes32, out_w32)
        if (pack32) {
            if (Tp2 == 32) PS_EMIT(uint32_t, 1);
            else if (Tp2 == 64) PS_EMIT(uint32_t, 2);
            else if (Tp2 == 128) PS_EMIT(uint32_t, 4);
            else PS_EMIT(uint32_t, 8);
        } else {
            if (Tp2 == 32) PS_EMIT(uint64_t, 1);
            else if (Tp2 == 64) PS_EMIT(uint64_t, 2);
            else if (Tp2 == 128) PS_EMIT(uint64_t, 4);
            else PS_EMIT(uint64_t, 8);
        }
#undef PS_EMIT
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

int g_walk_algo = 0;  // ps_walk_algo: 0 = sort-based kernel where it applies, 1 = hash-table kernel always

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
    const uint64_t thr = static_cast<uint64_t>(alpha * 4294967296.0);
    int dev = 0, sms = 148, occ = 1;
    PS_CUDA_CHECK(cudaGetDevice(&dev));
    PS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (g_walk_algo == 0 && n_hops <= 512 && T <= 256) {
        // sort-based kernel: the trace in shared memory (4 B per step), sorted in registers
        const int Tp2 = next_pow2(T);
        const int R = n_hops <= 128 ? 4 : (n_hops <= 256 ? 8 : 16);
        int tail_words = kSortListCap + kSortListCap / 2;       // the continuation list ...
        if (tail_words < 2 * Tp2) tail_words = 2 * Tp2;           // ... whose words sel (up to Tp2 64-bit keys) reuses after the walk
        int per_warp_words = 32 * R + 128 + 16 * R + 2 * R + 2 + tail_words;
        per_warp_words = (per_warp_words + 3) & ~3;
        const int warps = kMaxWarpsPerCta;
        const size_t smem = static_cast<size_t>(warps) * per_warp_words * sizeof(uint32_t);
        auto kern = R == 4 ? walk_sort_kernel<PtrT, kFromTrace, 4> : (R == 8 ? walk_sort_kernel<PtrT, kFromTrace, 8> : walk_sort_kernel<PtrT, kFromTrace, 16>);
        if (smem > 48 * 1024)
            PS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        PS_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, smem));
        if (occ < 1) occ = 1;
        int64_t blocks = ps_ceil_div(n, warps);
        const int64_t resident = static_cast<int64_t>(sms) * occ;  // persistent: grid-stride over sources
        if (blocks > resident * 4) blocks = resident * 4;
        kern<<<static_cast<unsigned>(blocks), warps * 32, smem, stream>>>(
            indptr, indices, sources, trace, n, n_hops, per_warp_words, Tp2, thr, fixed_len, T,
            static_cast<uint32_t>(seed & 0xFFFFFFFFull), static_cast<uint32_t>(seed >> 32), on64, ow64, on32, ow32, otrace);
        PS_LAUNCH_CHECK();
        return PS_OK;
    }
    const int cap = next_pow2(2 * n_hops < 64 ? 64 : 2 * n_hops);  // load factor <= 1/2
    int log2cap = 0;
    while ((1 << log2cap) < cap) ++log2cap;
    const int Tp2 = next_pow2(T);
    const int need = kHistBins + 2 * Tp2;  // hist + 64-bit sort keys, in words
    // per warp: keys | counts | extra; extra = the walk's continuation list, reused by hist / sel after the walk when
    // those do not fit into the dead tail of the key array
    const int list_words = kListCap + kListCap / 2;
    int per_warp_words = cap + cap / 2, hist_off;
    if (cap - n_hops >= need) {
        hist_off = cap - need;  // dead tail of the key array once the table is compacted
        per_warp_words += list_words;
    } else {
        hist_off = per_warp_words;
        per_warp_words += need > list_words ? need : list_words;
    }
    const size_t per_warp = static_cast<size_t>(per_warp_words) * sizeof(uint32_t);
    int warps = static_cast<int>((200 * 1024) / per_warp);
    if (warps > kMaxWarpsPerCta) warps = kMaxWarpsPerCta;
    PS_REQUIRE(warps >= 1, "n_hops=%d / T=%d too large for the shared-memory visit table", n_hops, T);
    const size_t smem = static_cast<size_t>(warps) * per_warp;
    auto kern = walk_topt_kernel<PtrT, kFromTrace>;
    if (smem > 48 * 1024)
        PS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    PS_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, smem));
    if (occ < 1) occ = 1;
    int64_t blocks = ps_ceil_div(n, warps);
    const int64_t resident = static_cast<int64_t>(sms) * occ;  // persistent: grid-stride over sources
    if (blocks > resident * 4) blocks = resident * 4;
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

extern "C" int ps_walk_algo(int mode) {
    const int old = g_walk_algo;
    if (mode == 0 || mode == 1) g_walk_algo = mode;
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
