// K12: training-batch construction on the device, one launch, no host round trip.
// Replaces sample_positives_with_rep + sample_easy_negatives of the reference
// (pinsage_training.py:53-77): B distinct random rows of `positives` (the law of
// randperm(P)[:B]) and B distinct random ids that occur in none of those pairs (the law of
// randperm over the masked id list, [:B]).
//
// Both are "first occurrences of an i.i.d. uniform candidate stream": candidate c of phase
// ph is floor(r64 * range / 2^64) with r64 = (x0 << 32 | x1) of
// Philox4x32-10(counter = (c, ph, step_lo, step_hi), key = seed); a candidate is kept iff no
// earlier candidate of the phase has the same value (and, for negatives, the value is not
// a node of the positive pairs); the first B kept candidates, in stream order, are the
// sample.  That is sequential rejection sampling, i.e. a uniformly random B-subset in
// uniformly random order, and it is launch-shape independent, so the CPU oracle
// (oracle.sample_batch_philox) reproduces it bit for bit.
//
// One small CTA (128 threads, no dynamic shared memory, so it fits beside a resident
// persistent GEMM CTA and never forces a shared-memory carve-out switch): candidates are
// inserted into a hash table in caller-provided global scratch (L2-resident; value ->
// smallest candidate index), kept flags are compacted in stream order with a block scan.
#include "common.cuh"
#include "../../include/pinsage_b200.h"

namespace {

constexpr int kThreads = 128;
constexpr uint32_t kEmpty = 0xFFFFFFFFu;

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t& o0, uint32_t& o1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    o0 = c0; o1 = c1;
}

__device__ __forceinline__ uint32_t candidate(uint32_t c, uint32_t phase, uint32_t s0, uint32_t s1, uint32_t k0, uint32_t k1,
                                              uint64_t range) {
    uint32_t x0, x1;
    philox4x32_10(c, phase, s0, s1, k0, k1, x0, x1);
    const uint64_t r64 = (static_cast<uint64_t>(x0) << 32) | x1;
    return static_cast<uint32_t>(__umul64hi(r64, range));
}

__device__ __forceinline__ uint32_t slot_of(uint32_t v, int hash_shift) { return (v * 0x9E3779B1u) >> hash_shift; }

// insert v (or find it); returns its slot
__device__ __forceinline__ uint32_t table_insert(uint32_t* tval, uint32_t v, uint32_t mask, int hash_shift) {
    uint32_t slot = slot_of(v, hash_shift);
    while (true) {
        const uint32_t old = atomicCAS(tval + slot, kEmpty, v);
        if (old == kEmpty || old == v) return slot;
        slot = (slot + 1) & mask;
    }
}

__device__ __forceinline__ uint32_t table_find(const uint32_t* tval, uint32_t v, uint32_t mask, int hash_shift) {
    uint32_t slot = slot_of(v, hash_shift);
    while (__ldcg(tval + slot) != v) slot = (slot + 1) & mask;  // v is always present; L2 reads: the table is written by atomics
    return slot;
}

__global__ void __launch_bounds__(kThreads)
sample_batch_kernel(const int64_t* __restrict__ positives, uint64_t P, const int64_t* __restrict__ all_ids, uint64_t n_items,
                    int B, int M, int cap, int hash_shift, int max_rounds, uint32_t k0, uint32_t k1, uint32_t s0, uint32_t s1,
                    uint32_t* __restrict__ table, int64_t* out, int* __restrict__ short_flag) {
    uint32_t* tval = table;
    uint32_t* tidx = table + cap;
    __shared__ int warp_sums[32];
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t mask = static_cast<uint32_t>(cap - 1);
    const int per = (M + kThreads - 1) / kThreads;  // candidates per thread, contiguous (stream order); <= 32
    if (tid < 32) warp_sums[tid] = 0;

    for (uint32_t phase = 0; phase < 2; ++phase) {
        const uint64_t range = phase == 0 ? P : n_items;
        for (int i = tid; i < cap; i += kThreads) { tval[i] = kEmpty; tidx[i] = kEmpty; }
        __syncthreads();
        if (phase == 1) {  // the nodes of the positive pairs can never be negatives (the reference masks positions of
                           // all_ids by node id, pinsage_training.py:69-72): index 0 beats every candidate
            for (int i = tid; i < 2 * B; i += kThreads) {
                const uint32_t v = static_cast<uint32_t>(out[(i >> 1) * 3 + (i & 1)]);
                atomicExch(tidx + table_insert(tval, v, mask, hash_shift), 0u);
            }
            __syncthreads();
        }
        int kept_total = 0;
        for (int round = 0; round < max_rounds && kept_total < B; ++round) {
            const uint32_t base = static_cast<uint32_t>(round) * static_cast<uint32_t>(M);
            for (int i = tid; i < M; i += kThreads) {
                const uint32_t v = candidate(base + i, phase, s0, s1, k0, k1, range);
                atomicMin(tidx + table_insert(tval, v, mask, hash_shift), base + i + 1u);
            }
            __syncthreads();
            uint32_t keep_mask = 0;
            int cnt = 0;
            for (int q = 0; q < per; ++q) {
                const int i = tid * per + q;
                if (i < M) {
                    const uint32_t v = candidate(base + i, phase, s0, s1, k0, k1, range);
                    if (__ldcg(tidx + table_find(tval, v, mask, hash_shift)) == base + i + 1u) { keep_mask |= 1u << q; ++cnt; }
                }
            }
            // exclusive block scan of cnt
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) warp_sums[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                int w = lane < kThreads / 32 ? warp_sums[lane] : 0;
                int wi = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, wi, o);
                    if (lane >= o) wi += t;
                }
                if (lane < kThreads / 32) warp_sums[lane] = wi - w;
                if (lane == 31) s_total = wi;
            }
            __syncthreads();
            int k = kept_total + warp_sums[warp] + incl - cnt;
            for (int q = 0; q < per; ++q) {
                if (keep_mask & (1u << q)) {
                    if (k < B) {
                        const int i = tid * per + q;
                        const uint32_t v = candidate(base + i, phase, s0, s1, k0, k1, range);
                        if (phase == 0) {
                            out[k * 3 + 0] = positives[2 * static_cast<uint64_t>(v)];
                            out[k * 3 + 1] = positives[2 * static_cast<uint64_t>(v) + 1];
                        } else {
                            out[k * 3 + 2] = all_ids ? all_ids[v] : static_cast<int64_t>(v);
                        }
                    }
                    ++k;
                }
            }
            kept_total += s_total;
            __syncthreads();
        }
        if (kept_total < B && tid == 0 && short_flag != nullptr) *short_flag = 1;  // cannot happen within the documented limits
        __syncthreads();
    }
}

}  // namespace

static int sampler_shape(int B, int* M, int* cap, int* log2cap) {
    *M = B + B / 2 + 32;  // candidates per round: > 1.25 B survive duplicates + exclusions at the limits below
    *cap = 1024; *log2cap = 10;
    while (*cap < 2 * (2 * B + 2 * *M)) { *cap <<= 1; ++*log2cap; }  // load factor <= 1/2 with two rounds
    return 0;
}

extern "C" int64_t ps_sample_batch_workspace(int B) {
    if (B <= 0) return 0;
    int M, cap, l2;
    sampler_shape(B, &M, &cap, &l2);
    return static_cast<int64_t>(cap) * 2 * sizeof(uint32_t);
}

extern "C" int ps_sample_batch(const int64_t* positives, int64_t P, const int64_t* all_ids, int64_t n_items, int B,
                               uint64_t seed, uint64_t step, int64_t* out_batch, void* workspace, int64_t workspace_bytes,
                               int* short_flag, ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(positives != nullptr && out_batch != nullptr && workspace != nullptr, "null pointer");
    PS_REQUIRE(B > 0 && B <= 2600, "ps_sample_batch supports 1 <= B <= 2600 (got %d)", B);
    PS_REQUIRE(P >= 16ll * B && P < 0xFFFFFFFFll, "ps_sample_batch needs 16*B <= P < 2^32 - 1 positives (P = %lld)", static_cast<long long>(P));
    PS_REQUIRE(n_items >= 16ll * B && n_items < (1ll << 31), "ps_sample_batch needs 16*B <= n_items < 2^31 (n_items = %lld)",
               static_cast<long long>(n_items));
    int M, cap, log2cap;
    sampler_shape(B, &M, &cap, &log2cap);
    PS_REQUIRE((M + kThreads - 1) / kThreads <= 32, "batch too large");
    PS_REQUIRE(workspace_bytes >= ps_sample_batch_workspace(B), "workspace too small (need ps_sample_batch_workspace(B) bytes)");
    const int max_rounds = (cap - 2 * B - cap / 8) / M;  // never fill the table
    sample_batch_kernel<<<1, kThreads, 0, stream>>>(positives, static_cast<uint64_t>(P), all_ids, static_cast<uint64_t>(n_items), B, M,
                                                    cap, 32 - log2cap, max_rounds, static_cast<uint32_t>(seed & 0xFFFFFFFFull),
                                                    static_cast<uint32_t>(seed >> 32), static_cast<uint32_t>(step & 0xFFFFFFFFull),
                                                    static_cast<uint32_t>(step >> 32), static_cast<uint32_t*>(workspace), out_batch, short_flag);
    PS_LAUNCH_CHECK();
    return PS_OK;
}
