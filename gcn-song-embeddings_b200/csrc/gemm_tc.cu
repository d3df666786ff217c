// tcgen05 (5th-gen tensor core) path of ps_gemm for sm_100a:
//     C[i,j] (+)= act( sum_r P(i,r) Q(j,r) + bias[j] )      fp32 in, fp32 out
// computed as an error-compensated 3xTF32 product (hi/lo split of both operands,
// a_lo*b_hi + a_hi*b_lo + a_hi*b_hi accumulated in fp32 in TMEM), which keeps the
// reference's fp32 parity (<= 1e-4) that plain TF32 cannot.
//
// Structure of one CTA (one 128 x BN output tile, BN = 128 or 256, optional split-K):
//   warps 8-15 producers: LDG.128 the operand rows (row gather folded in, K4), split each
//              fp32 into hi/lo, STS.128 into the UMMA canonical shared-memory layout
//              (SWIZZLE_128B K-major or SWIZZLE_128B_BASE32B MN-major), fence.proxy.async,
//              arrive on the stage's mbarrier;
//   warp 16    lane 0 issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8), 12 per
//              32-wide k-block, into one of two TMEM accumulators; tcgen05.commit frees the
//              smem stage and, every 2 k-blocks, publishes the accumulator;
//   warps 0-7  tcgen05.ld the finished accumulator (32x32b, one output row x BN/2 columns per
//              thread) and add it to fp32 registers -- the tensor core's TMEM adder truncates
//              (measured ~2e-8 relative per MMA, biased), so the truncating chain is kept to
//              24 MMAs and the long sum is round-to-nearest; then the epilogue from registers:
//              bias + leaky_relu + row L2-normalise and store, or red.global.add for split-K
//              weight gradients.  Registers are rebalanced with setmaxnreg (168 / 56 / 32 of the 96 per thread the CTA is launched with: the
//              pool only holds what the CTA's own warps give back).
// Large GEMMs run on PAIRS of CTAs (2-CTA clusters, two vertically adjacent tiles): one 256-row
// tcgen05.mma.cta_group::2 per step issued by the leader, each CTA staging its own A rows and half of the Q operand; the
// packed-weight kernels hand their output tiles to the TMA engine (cp.async.bulk.tensor stores of SWIZZLE_128B boxes).
// ps_gemm_tc_cluster selects the mode.
// Replaces nn.Linear / AddmmBackward of ConvLayer and the head (pinsage_model.py:201,
// 208-210, 259).
#include "common.cuh"
#include "../../include/pinsage_b200.h"
#include <cuda.h>  // CUtensorMap (the encode function is fetched through cudaGetDriverEntryPoint: no link against libcuda)
#include <mutex>

namespace {

constexpr int BM = 128;                 // UMMA M
constexpr int BK = 32;                  // floats per k-block = one 128-byte swizzle row
constexpr int kProducerThreads = 256;
constexpr int kThreads = 640;
constexpr uint32_t kHiMask = 0xFFFFE000u;  // keep sign, exponent and the 10 tf32 mantissa bits
constexpr int kEpiStride = 20;             // floats per staged row: 16 columns + 4 pad (conflict-free STS.128)
constexpr int kPrefetchKb = 6;             // weight-gradient producers: L2 prefetch distance in k-blocks
__device__ int g_tc_prefetch = 1;          // ps_gemm_tc_prefetch (development switch for the A/B comparison)

struct TcArgs {
    const float* P; int64_t ldp; const int32_t* p_rows;
    const float* Q; int64_t ldq; const int32_t* q_rows;
    float* C; int64_t ldc;
    int64_t M, N, K;
    const float* bias; float* norm_out;
    int act, l2norm, accumulate;
    int kb_per_split;   // k-blocks (of BK) per split
    int64_t mt, nt, zs;  // work grid: M tiles x N tiles x K splits
    const uint8_t* bpack;  // BPACK kernels: hi/lo images of the Q operand, one [2][BN x 128 B] block per (n-tile, k-block)
    uint32_t* mask; int64_t ldm;  // optional sign mask of the activation, one bit per output element (ldm in 32-bit words)
    unsigned long long* dbg;  // development: per-CTA cycle counters of the role waits (ps_gemm_tc_trace), nullptr = off
    int exp;                  // development (ps_gemm_tc_experiment; results are WRONG when set): bit 0 = the activation producers skip
                              // their global loads, bit 1 = the weight stream skips its bulk copies, bit 2 = the epilogue skips its stores
    // candidate filter (ps_gemm_filter, packed-weight kernels): no C store; element (i, j) >= filt_thr[j] appends
    // (value, i) to list j: filt_cnt[j]++ -> slot; filt_val / filt_row [j, slot] when slot < filt_cap
    const float* filt_thr; int32_t* filt_cnt; float* filt_val; int32_t* filt_row; int filt_cap;
    float* p_colsum;  // weight-gradient kernels (MN-major x MN-major): p_colsum[i] += sum_r P(i, r) -- the bias gradient, from the
                      // elements the producers already hold for the hi/lo split (replaces a separate pass over P)
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 2000000000ll) __trap();
    }
}
__device__ __forceinline__ void mbar_wait_timed(uint32_t bar, uint32_t parity, unsigned long long* dbg, long long& acc) {
    if (dbg == nullptr) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine, no tensor map): completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// the same, delivered to the same shared-memory offset of every CTA in `mask` (each one's mbarrier at that offset gets the bytes)
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// TMA tensor-map store of one [32 rows x 32 floats] box (SWIZZLE_128B in shared memory) to C at (col, row); rows / columns
// beyond the matrix are clipped by the hardware.  Bulk-group completion: wait_read<N> returns once all but the N most recent
// groups of this thread have finished READING shared memory.
__device__ __forceinline__ void tma_store_box(const CUtensorMap* map, uint32_t smem_src, int col, int row) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(col), "r"(row), "r"(smem_src) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// cta_group::2 forms (tile PAIRS on one 256-row MMA): allocation by one warp of EACH CTA of the pair, MMA and commits by the
// leader CTA only; a commit arrives on the barrier at this offset in both CTAs
template <int COLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit_multicast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(bar), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// arrive on the mbarrier at this offset in every CTA of `mask` once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 64-bit shared-memory matrix descriptor (sm_100 version bit set).  layout: 2 = SWIZZLE_128B
// (K-major operands), 1 = SWIZZLE_128B_BASE32B (the only layout tf32 MN-major operands may use).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
    const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

__device__ __forceinline__ void split_store(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, float4 v) {
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & kHiMask); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & kHiMask); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & kHiMask); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & kHiMask); l.w = v.w - h.w;
    *reinterpret_cast<float4*>(hi_base + off) = h;
    *reinterpret_cast<float4*>(lo_base + off) = l;
}

// Producer-side view of one operand: everything that does not change from k-block to k-block is resolved
// once per tile (row gather, row clamping, swizzled shared-memory offsets), so the per-k-block work of a
// thread is LDG.128 + hi/lo split + 2 x STS.128 per float4 and nothing else.
//   KMAJOR : element(i, r) = X[row(i)*ld + r]; SWIZZLE_128B: smem row i = 128 B, 16-B chunk c stored at c ^ (i % 8),
//            8-row groups 1024 B apart (SBO).  Thread t owns chunk c = t % 8 of rows (t / 8) + 32 j.
//   MNMAJOR: element(i, r) = X[row(r)*ld + i]; SWIZZLE_128B_BASE32B: atoms of 4 k-rows x 32 i (512 B), 32-B chunk c
//            of k-row r stored at c ^ (r % 4); atoms of consecutive i-chunks 512 B apart (LBO), 4-row k-groups
//            (ROWS/32)*512 B apart (SBO).  Thread t owns float4 c4 = t % (ROWS/4) of k-rows t / (ROWS/4) + RPP j.
// Rows / columns beyond the matrix edge are clamped to the last valid one: they only feed output rows /
// columns that are never stored.  Only the contraction tail (k >= K) must read as zero.
template <bool KMAJOR, int ROWS, bool PRECOMP>
struct Operand {
    static constexpr int NJ = KMAJOR ? ROWS / 32 : (BK * ROWS / 4) / kProducerThreads;  // float4 per thread per k-block
    static constexpr int C4 = ROWS / 4;
    static constexpr int RPP = kProducerThreads / C4;  // MN-major: k-rows per pass
    // K-major + PRECOMP: one resolved pointer per owned row (row gather / clamp done once per tile).
    // K-major, !PRECOMP: row pointers are rebuilt per k-block from (row0, ext, ld) -- 3 instructions, no registers.
    // MN-major: one column base pointer; the row is the contraction index.
    const float* base[(KMAJOR && PRECOMP) ? NJ : 1];
    const int32_t* rows;  // MN-major: optional gather on the contraction index
    int64_t ld;
    int row0, ext;        // K-major, !PRECOMP
    uint32_t off0;        // smem byte offset of the j = 0 element
    int kq;               // K-major: first k of this thread inside a k-block; MN-major: first k-row

    __device__ __forceinline__ void init(const float* X, int64_t ld_, const int32_t* rows_, int64_t i0, int64_t ext_i, int t) {
        ld = ld_;
        if (KMAJOR) {
            const int c = t & 7, tr = t >> 3;
            kq = c * 4;
            rows = nullptr;
            if (PRECOMP) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    int64_t i = i0 + tr + 32 * j;
                    if (i >= ext_i) i = ext_i - 1;
                    const int64_t row = rows_ ? static_cast<int64_t>(__ldg(rows_ + i)) : i;
                    base[j] = X + row * ld_ + kq;
                }
            } else {
                base[0] = X + kq;
                row0 = static_cast<int>(i0) + tr;
                ext = static_cast<int>(ext_i);
            }
            off0 = (tr >> 3) * 1024 + (tr & 7) * 128 + ((c ^ (tr & 7)) << 4);
        } else {
            const int c4 = t % C4, tk = t / C4;
            kq = tk;
            rows = rows_;
            int64_t i = i0 + c4 * 4;
            if (i >= ext_i) i = 0;
            base[0] = X + i;
            off0 = (tk >> 2) * (ROWS / 32) * 512 + (c4 >> 3) * 512 + (tk & 3) * 128 + ((((c4 & 7) >> 1) ^ (tk & 3)) << 5) + ((c4 & 1) << 4);
        }
    }
    // byte distance in smem between the j-th and (j+1)-th element of this thread
    static constexpr uint32_t kStep = KMAJOR ? 4096u : (RPP / 4) * (ROWS / 32) * 512u;

    // load + split + store elements [J0, J0 + CNT) of this thread for the k-block at k0
    template <int J0, int CNT>
    __device__ __forceinline__ void load(float4 (&v)[CNT], int64_t k0, int64_t K) const {
        if (KMAJOR) {
            const bool ok = k0 + kq < K;
#pragma unroll
            for (int j = 0; j < CNT; ++j) {
                v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) {
                    if (PRECOMP) v[j] = ps_ldg4(base[J0 + j] + k0);
                    else v[j] = ps_ldg4(base[0] + static_cast<int64_t>(min(row0 + 32 * (J0 + j), ext - 1)) * ld + k0);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < CNT; ++j) {
                const int64_t k = k0 + kq + RPP * (J0 + j);
                v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < K) {
                    const int64_t row = rows ? static_cast<int64_t>(__ldg(rows + k)) : k;
                    v[j] = ps_ldg4(base[0] + row * ld);
                }
            }
        }
    }
    // L2 prefetch of the elements this thread will load for the k-block at k0 (no registers held; the later
    // LDG then pays an L2 hit instead of an HBM round trip).  Gathered MN-major rows are skipped (their
    // address needs the index first).
    __device__ __forceinline__ void prefetch(int64_t k0, int64_t K) const {
        if (KMAJOR) {
            if (k0 + kq < K && (kq == 0 || kq == 28)) {  // the two ends of the 128-byte row segment cover its cache line(s)
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const float* p = PRECOMP ? base[j] + k0
                                             : base[0] + static_cast<int64_t>(min(row0 + 32 * j, ext - 1)) * ld + k0;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                }
            }
        } else if (rows == nullptr) {
            if ((threadIdx.x & 7) == 0) {  // one thread per 128 bytes of a k-row
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int64_t k = k0 + kq + RPP * j;
                    if (k < K) asm volatile("prefetch.global.L2 [%0];" ::"l"(base[0] + k * ld));
                }
            }
        }
    }
    template <int J0, int CNT>
    __device__ __forceinline__ void store(uint8_t* hi, uint8_t* lo, const float4 (&v)[CNT]) const {
#pragma unroll
        for (int j = 0; j < CNT; ++j) split_store(hi, lo, off0 + (J0 + j) * kStep, v[j]);
    }
};

// One unit of work of the persistent kernel: a 128 x BN output tile over a range of k-blocks.
struct Work {
    int64_t m0, n0;
    int kb_begin, num_kb;
};
__device__ __forceinline__ Work decode_work(const TcArgs& a, int64_t w, int bn) {
    const int64_t n_idx = w % a.nt, rest = w / a.nt;
    const int64_t m_idx = rest % a.mt, z = rest / a.mt;
    const int64_t nkb = (a.K + BK - 1) / BK;
    const int64_t b = z * a.kb_per_split;
    const int64_t e = min(nkb, b + a.kb_per_split);
    return {m_idx * BM, n_idx * static_cast<int64_t>(bn), static_cast<int>(b), static_cast<int>(e - b)};
}

// Load-side cursor of the MN-major x MN-major (weight-gradient) producers: the operand views of the k-block whose
// asynchronous copies are issued next, one k-block ahead of the split pass, across tile boundaries.
template <typename OpA, typename OpB>
struct LoadCursor {
    OpA a; OpB b;
    int64_t w; int kb; int nkb; int64_t k0; bool valid;
    int64_t m0, n0;  // origin of the cursor's work item (L2 prefetch addresses)
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Thread roles (20 warps): 0-7 accumulate + epilogue, 8-15 producers, 16 MMA issue (17-19 only give their
// registers away: setmaxnreg works on whole warpgroups).  Persistent: every role walks the same list of work
// items (blockIdx.x, + gridDim.x, ...), so the producers and the MMA warp run ahead into the next tile while
// the accumulate warps are still storing the previous one.
// BPACK: the Q operand (a weight matrix) was split into hi/lo and laid out as K-major SWIZZLE_128B tile images by
// pack_b_kernel; one thread streams the image of every k-block into the stage with ONE bulk copy, and the eight
// producer warps only move the activation operand, double-buffered in registers across k-blocks and tiles.
// CG2 (with CL == 2): the pair runs ONE 256-row tcgen05.mma.cta_group::2 per step instead of two 128-row ones.  Each CTA
// stages its own 128 rows of A and only HALF of the weight image (BN/2 columns); the tensor cores of the two SMs exchange the
// halves, so per k-block a CTA writes 64 KB and the MMAs read 96 KB of shared memory instead of 96 KB and 144 KB: the
// single-CTA 128 x 256 3xTF32 tile is bound by shared-memory bandwidth (240 KB per 1536-cycle k-block against 128 B/clk).
template <bool PK, bool QK, int BN, bool BPACK, int MASK, int CL, bool CG2 = false>  // MASK: 0 none, 1 = write sign bits (act 1), 2 = apply them (act 2); CL: CTAs per cluster
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(TcArgs a, const __grid_constant__ CUtensorMap cmap) {
    static_assert(!BPACK || (PK && QK), "packed-B kernels take a K-major P and lay Q out K-major");
    static_assert(CL == 1 || (BPACK && CL == 2) || (CG2 && CL == 2 && !PK && !QK), "clusters: pairs; packed-B kernels, or the cta_group::2 weight-gradient kernel");
    static_assert(!CG2 || CL == 2, "cta_group::2 needs the pair");
    constexpr int BNS = CG2 ? BN / 2 : BN;  // columns of the Q operand ONE CTA stages
    constexpr int STAGES = (CG2 && !BPACK) ? 3 : (BN == 256 ? 2 : 3);  // the cta_group::2 weight-gradient kernel: 64 KB stages, no store boxes
    constexpr int HALF = BN / 2;            // columns owned by one epilogue thread
    // k-blocks accumulated in TMEM before the fp32 register drain.  Measured with 4 for the packed-weight kernels: q_fwd_l0 0.96 ->
    // 0.90 ms (half the drains through the 64 B/clk TMEM read port), max error vs fp64 1.1e-6 instead of 6.2e-7 -- and the model-level
    // parity test at (512, 1024, 512) dims then misses 1e-4 on a gradient, so the chain stays at 24 MMAs.
    constexpr int CHUNK_KB = 2;
    constexpr uint32_t A_BYTES = BM * 128, B_BYTES = BN * 128;
    constexpr uint32_t B_STAGE = CG2 ? B_BYTES / 2 : B_BYTES;  // bytes of ONE of the two (hi / lo) weight tiles a CTA stages
    constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_STAGE;
    extern __shared__ uint8_t smem_raw[];
    // aligned by pointer arithmetic on the __shared__ array (not through an integer cast), so the compiler keeps the
    // shared address space and emits LDS / STS instead of generic loads / stores
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    float* ss_buf = reinterpret_cast<float*>(tmem_slot + 4);  // [2 parities][2 halves][128] partial sums of squares (l2norm)
    float* bias_s = ss_buf + 2 * 2 * 128;                     // [2 parities][BN] bias of the current tile
    float* epi_buf = bias_s + 2 * BN;                         // [8 warps][32 rows][kEpiStride] output staging (coalesced stores)
    // CG2: the staged C tile leaves through the TMA engine: per accumulate warp two [32 rows x 128 B] boxes (SWIZZLE_128B,
    // 1024-byte aligned), behind everything else
    uint8_t* tbox = smem + STAGES * STAGE_BYTES + 8192;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const uint32_t tfull0 = smem_u32(bars + 2 * STAGES), tempty0 = smem_u32(bars + 2 * STAGES + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // CL == 2: two CTAs of a cluster work on vertically adjacent tiles (m, m + 1) of the SAME n-tile in lock step; each
    // streams HALF of the packed weight image of a k-block and multicasts it into both CTAs' stages, so a weight byte
    // crosses L2 -> SM once per tile PAIR.  (Measured before: these kernels move ~7 TB/s of L2 traffic, most of it the
    // weight images re-streamed for every 128-row tile, and the tensor pipe idles 55-70 % of the time behind it.)
    uint32_t cta_rank = 0;
    if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    const int64_t w_first = CL > 1 ? blockIdx.x / CL : blockIdx.x;
    const int64_t w_step = CL > 1 ? gridDim.x / CL : gridDim.x;
    const int64_t m_pairs = (a.mt + CL - 1) / CL;
    const int64_t total_work = CL > 1 ? m_pairs * a.nt * (BPACK ? 1 : a.zs) : a.mt * a.nt * a.zs;
    auto decode = [&](int64_t w) -> Work {
        if (CL > 1 && BPACK) {  // n-tile fastest, then tile pairs; no split-K on this path; a pair's second tile may lie beyond M
            const int64_t n_idx = w % a.nt, mp = w / a.nt;
            return {(mp * CL + cta_rank) * BM, n_idx * static_cast<int64_t>(BN), 0, static_cast<int>((a.K + BK - 1) / BK)};
        }
        if (CL > 1) {  // weight gradients: tile pairs x n-tiles x k-splits
            const int64_t n_idx = w % a.nt, rest = w / a.nt;
            const int64_t mp = rest % m_pairs, z = rest / m_pairs;
            const int64_t nkb = (a.K + BK - 1) / BK;
            const int64_t b = z * a.kb_per_split;
            const int64_t e = min(nkb, b + a.kb_per_split);
            return {(mp * CL + cta_rank) * BM, n_idx * static_cast<int64_t>(BN), static_cast<int>(b), static_cast<int>(e - b)};
        }
        return decode_work(a, w, BN);
    };

    if (tid == 0) {
        // full: the producers + the weight stream's expect_tx arrive (+ with CG2, on the leader, the peer's forwarded "my half is staged");
        // empty: a stage is free when every MMA that reads it has completed: one commit per CTA of the cluster, or the leader's one commit
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, kProducerThreads + (BPACK ? 1 : 0) + ((CG2 && cta_rank == 0) ? 1 : 0));
            mbar_init(empty0 + 8 * s, CG2 ? 1 : CL);
        }
        // tempty: with CG2 the leader's MMA thread needs BOTH CTAs' accumulate warps to have drained the buffer
        for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, CG2 ? 16 : 8); }
        fence_barrier_init();
    }
    if (warp == 16) {
        if (CG2) tmem_alloc2<2 * BN>(smem_u32(tmem_slot));
        else tmem_alloc<2 * BN>(smem_u32(tmem_slot));
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) {  // the peer's barriers exist before anything is sent to them
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 16) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        // ------------------------------------------------ MMA issue (one lane of warp 16)
        if (warp == 16 && lane == 0 && (!CG2 || cta_rank == 0)) {
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((PK ? 0u : 1u) << 15) | ((QK ? 0u : 1u) << 16) |
                                       (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>((CG2 ? 2 * BM : BM) >> 4) << 24);
            constexpr uint32_t A_LBO = PK ? 16 : 512, A_SBO = PK ? 1024 : (BM / 32) * 512, A_STEP = PK ? 32 : 2 * (BM / 32) * 512;
            constexpr uint32_t B_LBO = QK ? 16 : 512, B_SBO = QK ? 1024 : (BNS / 32) * 512, B_STEP = QK ? 32 : 2 * (BNS / 32) * 512;
            constexpr uint32_t A_LAY = PK ? 2u : 1u, B_LAY = QK ? 2u : 1u;
            int stage = 0; uint32_t phase = 0;
            uint32_t gc = 0;  // chunks issued so far (selects the TMEM buffer and its barrier parity)
            long long d_full = 0, d_tempty = 0;
            if (a.exp & 0x18) {  // development: stagger the CTAs (pairs) in time so that their epilogue stores do not all burst at once
                const int phase_of = static_cast<int>((blockIdx.x / CL) % 4);
                const long long t_end = clock64() + static_cast<long long>(phase_of) * ((a.exp & 0x10) ? 16000 : 8000);
                while (clock64() < t_end) __nanosleep(200);
            }
            const long long d_t0 = clock64();
            for (int64_t w = w_first; w < total_work; w += w_step) {
                const Work wk = decode(w);
                int kb = 0;
                while (kb < wk.num_kb) {
                    const uint32_t buf = gc & 1;
                    mbar_wait_timed(tempty0 + 8 * buf, ((gc >> 1) & 1) ^ 1, a.dbg, d_tempty);  // the drain of this buffer's previous chunk is done
                    tc_fence_after();
                    const uint32_t tacc = tmem_base + buf * BN;
                    for (int q = 0; q < CHUNK_KB && kb < wk.num_kb; ++q, ++kb) {
                        mbar_wait_timed(full0 + 8 * stage, phase, a.dbg, d_full);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                        const uint32_t a_hi = sa, a_lo = sa + A_BYTES, b_hi = sa + 2 * A_BYTES, b_lo = b_hi + B_STAGE;
#pragma unroll
                        for (int s = 0; s < BK / 8; ++s) {
                            const uint64_t dah = make_desc(a_hi + s * A_STEP, A_LBO, A_SBO, A_LAY), dal = make_desc(a_lo + s * A_STEP, A_LBO, A_SBO, A_LAY);
                            const uint64_t dbh = make_desc(b_hi + s * B_STEP, B_LBO, B_SBO, B_LAY), dbl = make_desc(b_lo + s * B_STEP, B_LBO, B_SBO, B_LAY);
                            if (CG2) {
                                umma2_tf32(tacc, dal, dbh, idesc, (q | s) != 0);
                                umma2_tf32(tacc, dah, dbl, idesc, 1u);
                                umma2_tf32(tacc, dah, dbh, idesc, 1u);
                            } else {
                                umma_tf32(tacc, dal, dbh, idesc, (q | s) != 0);  // a chunk starts a fresh accumulator; small terms first
                                umma_tf32(tacc, dah, dbl, idesc, 1u);
                                umma_tf32(tacc, dah, dbh, idesc, 1u);
                            }
                        }
                        if (CG2) umma2_commit_multicast(empty0 + 8 * stage, 3);  // the stage is free in both CTAs
                        else if (CL > 1) umma_commit_multicast(empty0 + 8 * stage, static_cast<uint16_t>((1u << CL) - 1));  // ... in every CTA of the cluster
                        else umma_commit(empty0 + 8 * stage);  // frees the smem stage once these MMAs have read it
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    if (CG2) umma2_commit_multicast(tfull0 + 8 * buf, 3);  // publishes the chunk to the accumulate warps of both CTAs
                    else umma_commit(tfull0 + 8 * buf);  // publishes the chunk to the accumulate warps
                    ++gc;
                }
            }
            if (a.dbg) { a.dbg[blockIdx.x * 8 + 0] = clock64() - d_t0; a.dbg[blockIdx.x * 8 + 1] = d_full; a.dbg[blockIdx.x * 8 + 2] = d_tempty; }
        }
        // ------------------------------------------------ packed-B stream (one lane of warp 17)
        if (BPACK && warp == 17 && lane == 0) {
            const int64_t nkb = (a.K + BK - 1) / BK;
            int stage = 0; uint32_t phase = 0;
            long long d_empty = 0;
            for (int64_t w = w_first; w < total_work; w += w_step) {
                const Work wk = decode(w);
                const uint8_t* src = a.bpack + ((wk.n0 / BN) * nkb + wk.kb_begin) * static_cast<int64_t>(2 * B_BYTES);
                for (int kb = 0; kb < wk.num_kb; ++kb, src += 2 * B_BYTES) {
                    mbar_wait_timed(empty0 + 8 * stage, phase ^ 1, a.dbg, d_empty);
                    if (a.exp & 2) { mbar_arrive(full0 + 8 * stage); if (++stage == STAGES) { stage = 0; phase ^= 1; } continue; }
                    mbar_arrive_expect_tx(full0 + 8 * stage, 2 * B_STAGE);
                    if (CG2) {  // this CTA's BN/2 columns of the hi and of the lo tile, into its own stage only
                        const uint32_t dst = smem_u32(smem + stage * STAGE_BYTES + 2 * A_BYTES);
                        bulk_g2s(dst, src + cta_rank * B_STAGE, B_STAGE, full0 + 8 * stage);
                        bulk_g2s(dst + B_STAGE, src + B_BYTES + cta_rank * B_STAGE, B_STAGE, full0 + 8 * stage);
                    } else if (CL > 1) {  // this CTA's half of the image ([hi | lo]: rank 0 sends hi, rank 1 lo), into both CTAs
                        bulk_g2s_multicast(smem_u32(smem + stage * STAGE_BYTES + 2 * A_BYTES + cta_rank * B_BYTES), src + cta_rank * B_BYTES, B_BYTES,
                                           full0 + 8 * stage, static_cast<uint16_t>((1u << CL) - 1));
                    } else {
                        bulk_g2s(smem_u32(smem + stage * STAGE_BYTES + 2 * A_BYTES), src, 2 * B_BYTES, full0 + 8 * stage);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            if (a.dbg) a.dbg[blockIdx.x * 8 + 3] = d_empty;
        }
        // ------------------------------------------------ CG2, peer CTA: tell the leader's MMA thread when this CTA's half of a stage is staged
        if (CG2 && cta_rank != 0 && warp == 18 && lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t w = w_first; w < total_work; w += w_step) {
                const Work wk = decode(w);
                for (int kb = 0; kb < wk.num_kb; ++kb) {
                    mbar_wait(full0 + 8 * stage, phase);
                    mbar_arrive_remote(full0 + 8 * stage, 0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 8 && BPACK) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // ------------------------------------------------ producers, activation operand only (8 warps)
        const int t = tid - 8 * 32;
        using OpA = Operand<true, BM, true>;
        constexpr int NJ = OpA::NJ;
        OpA op;                       // the LOAD cursor: one k-block ahead of the stores, across tile boundaries
        int64_t lw = w_first;
        int lkb = 0;
        Work lwk = decode(lw < total_work ? lw : 0);
        float4 cur[NJ], nxt[NJ];
        if (lw < total_work) {
            op.init(a.P, a.ldp, a.p_rows, lwk.m0, a.M, t);
            op.template load<0, NJ>(cur, static_cast<int64_t>(lwk.kb_begin) * BK, a.K);
        }
        const uint32_t off0 = op.off0;  // shared-memory offsets depend on the thread only, not on the tile
        int stage = 0; uint32_t phase = 0;
        long long d_pempty = 0;
        for (int64_t w = w_first; w < total_work; w += w_step) {
            const Work wk = decode(w);
            // L2 prefetch of the NEXT work item's activation rows (one 128-byte line per row and k-block).  The register
            // double buffer keeps one k-block (16 KB per SM) in flight, which at HBM latency caps the operand stream near
            // 2 TB/s chip-wide (the kernel's top stall was long-scoreboard in the producers); a whole tile ahead the lines
            // arrive in L2 while the current tile is multiplied, and the loads below pay an L2 hit instead.
            if (g_tc_prefetch && w + w_step < total_work) {
                const Work nwk = decode(w + w_step);
                int64_t i = nwk.m0 + (t & (BM - 1));
                if (i >= a.M) i = a.M - 1;
                const int64_t row = a.p_rows ? static_cast<int64_t>(__ldg(a.p_rows + i)) : i;
                const float* src = a.P + row * a.ldp + static_cast<int64_t>(nwk.kb_begin) * BK;
                for (int j = t >> 7; j < nwk.num_kb; j += kProducerThreads / BM)
                    if (static_cast<int64_t>(nwk.kb_begin + j) * BK < a.K) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + j * BK));
            }
            for (int kb = 0; kb < wk.num_kb; ++kb) {
                // advance the load cursor and issue the next k-block's loads before waiting for the stage
                bool have_next = true;
                if (++lkb == lwk.num_kb) {
                    lw += w_step; lkb = 0;
                    have_next = lw < total_work;
                    if (have_next) { lwk = decode(lw); op.init(a.P, a.ldp, a.p_rows, lwk.m0, a.M, t); }
                }
                if (have_next && !(a.exp & 1)) op.template load<0, NJ>(nxt, static_cast<int64_t>(lwk.kb_begin + lkb) * BK, a.K);
                mbar_wait_timed(empty0 + 8 * stage, phase ^ 1, t == 0 ? a.dbg : nullptr, d_pempty);
                uint8_t* st = smem + stage * STAGE_BYTES;
#pragma unroll
                for (int j = 0; j < NJ; ++j) split_store(st, st + A_BYTES, off0 + j * OpA::kStep, cur[j]);
                fence_proxy_async();
                mbar_arrive(full0 + 8 * stage);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
#pragma unroll
                for (int j = 0; j < NJ; ++j) cur[j] = nxt[j];
            }
        }
        if (a.dbg && t == 0) a.dbg[blockIdx.x * 8 + 4] = d_pempty;
    } else if (warp >= 8 && !PK && !QK) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // ------------------------------------------------ producers, MN-major x MN-major (weight gradients)
        // The tensor core reads an fp32 word as tf32 by ignoring its low 13 mantissa bits, so the RAW operand is its
        // own "hi" tile: cp.async (LDGSTS, no registers held) copies the k-block straight into the hi tiles of the
        // stage, one k-block ahead of use, and the split pass only reads it back from shared memory, computes
        // lo = x - hi(x) and writes the lo tiles.  Loading through registers kept 8 x 16 B per thread in flight and
        // exposed ~4 memory latencies per k-block (index -> row, twice); this keeps the whole next k-block in flight.
        const int t = tid - 8 * 32;
        using OpA = Operand<false, BM, true>;
        using OpB = Operand<false, BNS, false>;  // CG2: this CTA's half of the Q columns
        const int64_t n_half = CG2 ? static_cast<int64_t>(cta_rank) * BNS : 0;
        LoadCursor<OpA, OpB> lc;
        lc.w = w_first; lc.kb = 0;
        lc.valid = lc.w < total_work;
        {
            const Work wk = decode(lc.valid ? lc.w : 0);
            lc.nkb = wk.num_kb;
            lc.k0 = static_cast<int64_t>(wk.kb_begin) * BK;
            lc.m0 = wk.m0; lc.n0 = wk.n0;
            lc.a.init(a.P, a.ldp, a.p_rows, wk.m0, a.M, t);
            lc.b.init(a.Q, a.ldq, a.q_rows, wk.n0 + n_half, a.N, t);
        }
        const uint32_t offa = lc.a.off0, offb = lc.b.off0;  // shared-memory offsets depend on the thread only
        const uint32_t smem0 = smem_u32(smem);
        auto issue = [&](int stage_) {  // asynchronous copies of the cursor's k-block into the hi tiles of stage_
            const uint32_t sa = smem0 + stage_ * STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < OpA::NJ; ++j) {
                const int64_t k = lc.k0 + lc.a.kq + OpA::RPP * j;
                const bool ok = k < a.K;
                const int64_t row = ok ? (lc.a.rows ? static_cast<int64_t>(__ldg(lc.a.rows + k)) : k) : 0;
                cp_async16(sa + offa + j * OpA::kStep, lc.a.base[0] + row * lc.a.ld, ok ? 16u : 0u);
            }
#pragma unroll
            for (int j = 0; j < OpB::NJ; ++j) {
                const int64_t k = lc.k0 + lc.b.kq + OpB::RPP * j;
                const bool ok = k < a.K;
                const int64_t row = ok ? (lc.b.rows ? static_cast<int64_t>(__ldg(lc.b.rows + k)) : k) : 0;
                cp_async16(sa + 2 * A_BYTES + offb + j * OpB::kStep, lc.b.base[0] + row * lc.b.ld, ok ? 16u : 0u);
            }
            cp_async_commit();
        };
        auto advance = [&]() {
            if (++lc.kb == lc.nkb) {
                lc.w += w_step; lc.kb = 0;
                lc.valid = lc.w < total_work;
                if (lc.valid) {
                    const Work wk = decode(lc.w);
                    lc.nkb = wk.num_kb;
                    lc.k0 = static_cast<int64_t>(wk.kb_begin) * BK;
                    lc.m0 = wk.m0; lc.n0 = wk.n0;
                    lc.a.init(a.P, a.ldp, a.p_rows, wk.m0, a.M, t);
                    lc.b.init(a.Q, a.ldq, a.q_rows, wk.n0 + n_half, a.N, t);
                }
            } else {
                lc.k0 += BK;
            }
        };
        int stage = 0; uint32_t phase = 0;
        if (lc.valid) {  // prologue: the first k-block (its stage is free)
            issue(0);
            advance();
        }
        float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);  // this thread's 4 P columns summed over the work item's k range
        for (int64_t w = w_first; w < total_work; w += w_step) {
            const Work wk = decode(w);
            for (int kb = 0; kb < wk.num_kb; ++kb) {
                // split pass of the current k-block: its copies were issued one iteration ago
                cp_async_wait_all();
                uint8_t* st = smem + stage * STAGE_BYTES;
#pragma unroll
                for (int j = 0; j < OpA::NJ; ++j) {
                    float4* hp = reinterpret_cast<float4*>(st + offa + j * OpA::kStep);
                    const float4 v = *hp;
                    csum.x += v.x; csum.y += v.y; csum.z += v.z; csum.w += v.w;  // rows k >= K were zero-filled
                    float4 l;
                    l.x = v.x - __uint_as_float(__float_as_uint(v.x) & kHiMask); l.y = v.y - __uint_as_float(__float_as_uint(v.y) & kHiMask);
                    l.z = v.z - __uint_as_float(__float_as_uint(v.z) & kHiMask); l.w = v.w - __uint_as_float(__float_as_uint(v.w) & kHiMask);
                    *reinterpret_cast<float4*>(st + A_BYTES + offa + j * OpA::kStep) = l;
                }
#pragma unroll
                for (int j = 0; j < OpB::NJ; ++j) {
                    float4* hp = reinterpret_cast<float4*>(st + 2 * A_BYTES + offb + j * OpB::kStep);
                    const float4 v = *hp;
                    float4 l;
                    l.x = v.x - __uint_as_float(__float_as_uint(v.x) & kHiMask); l.y = v.y - __uint_as_float(__float_as_uint(v.y) & kHiMask);
                    l.z = v.z - __uint_as_float(__float_as_uint(v.z) & kHiMask); l.w = v.w - __uint_as_float(__float_as_uint(v.w) & kHiMask);
                    *reinterpret_cast<float4*>(st + 2 * A_BYTES + B_STAGE + offb + j * OpB::kStep) = l;
                }
                fence_proxy_async();
                mbar_arrive(full0 + 8 * stage);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
                // copies of the next k-block into the stage that follows (free once the MMAs that read it are done)
                if (lc.valid) {
                    if (g_tc_prefetch) {  // L2 prefetch kPrefetchKb k-blocks ahead of the copies: 32 k-rows x (BM + BN) floats
                        const int64_t kp = lc.k0 + static_cast<int64_t>(kPrefetchKb) * BK + (t >> 3);
                        if (kp < a.K && lc.kb + kPrefetchKb < lc.nkb) {
                            const int line = t & 7;  // a Q row of BN floats = BN / 32 lines, a P row of BM floats = BM / 32 lines
                            const int64_t qrow = lc.b.rows ? static_cast<int64_t>(__ldg(lc.b.rows + kp)) : kp;
                            if (line < BNS / 32 && lc.n0 + n_half + line * 32 < a.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.Q + qrow * a.ldq + lc.n0 + n_half + line * 32));
                            if (line < BM / 32 && lc.m0 + line * 32 < a.M) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.P + kp * a.ldp + lc.m0 + line * 32));
                        }
                    }
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    issue(stage);
                    advance();
                }
            }
            if (a.p_colsum != nullptr) {  // every (m-tile, k-split) reads its P slice once per n-tile: count n-tile 0 only
                const int64_t i = wk.m0 + (t % OpA::C4) * 4;
                if (wk.n0 == 0 && i < a.M) {  // M % 4 == 0 on this path: a float4 of columns is all-or-nothing
                    atomicAdd(a.p_colsum + i, csum.x); atomicAdd(a.p_colsum + i + 1, csum.y);
                    atomicAdd(a.p_colsum + i + 2, csum.z); atomicAdd(a.p_colsum + i + 3, csum.w);
                }
                csum = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    } else if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // ------------------------------------------------ producers (8 warps)
        const int t = tid - 8 * 32;
        using OpA = Operand<PK, BM, true>;
        using OpB = Operand<QK, BN, false>;
        constexpr int HB = OpB::NJ / 2;
        int stage = 0; uint32_t phase = 0;
        for (int64_t w = w_first; w < total_work; w += w_step) {
            const Work wk = decode(w);
            OpA opa;
            OpB opb;
            opa.init(a.P, a.ldp, a.p_rows, wk.m0, a.M, t);
            opb.init(a.Q, a.ldq, a.q_rows, wk.n0, a.N, t);
            for (int kb = 0; kb < wk.num_kb; ++kb) {
                const int64_t k0 = static_cast<int64_t>(wk.kb_begin + kb) * BK;
                if (kb + 1 < wk.num_kb) { opa.prefetch(k0 + BK, a.K); opb.prefetch(k0 + BK, a.K); }
                float4 va[OpA::NJ];
                opa.template load<0, OpA::NJ>(va, k0, a.K);  // issued before the stage is free: the loads overlap the wait
                float4 vb[HB];
                opb.template load<0, HB>(vb, k0, a.K);
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                uint8_t* st = smem + stage * STAGE_BYTES;
                opa.template store<0, OpA::NJ>(st, st + A_BYTES, va);
                opb.template store<0, HB>(st + 2 * A_BYTES, st + 2 * A_BYTES + B_BYTES, vb);
                opb.template load<HB, HB>(vb, k0, a.K);
                opb.template store<HB, HB>(st + 2 * A_BYTES, st + 2 * A_BYTES + B_BYTES, vb);
                fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core's async proxy
                mbar_arrive(full0 + 8 * stage);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
        // ------------------------------------------------ accumulate (fp32, round-to-nearest, in registers) + epilogue
        // Each chunk of CHUNK_KB k-blocks (24 MMAs) is summed by the tensor core in TMEM, whose adder truncates;
        // draining it into registers keeps the truncating chain short (fp32-level accuracy).
        const int quarter = warp & 3, half = warp >> 2;
        const uint32_t tlane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + half * HALF;
        uint32_t gc = 0, tile_par = 0;
        float acc[HALF];
        long long d_tfull = 0, d_epi = 0, d_store = 0;
        for (int64_t w = w_first; w < total_work; w += w_step, tile_par ^= 1) {
            const Work wk = decode(w);
            const int n_chunks = (wk.num_kb + CHUNK_KB - 1) / CHUNK_KB;
#pragma unroll
            for (int j = 0; j < HALF; ++j) acc[j] = 0.f;
            // what the epilogue needs from global memory is requested BEFORE the accumulator drain, so its latency overlaps it
            const int64_t row = wk.m0 + quarter * 32 + lane;
            const int64_t col0 = wk.n0 + half * HALF;
            const bool row_ok = row < a.M;
            float* bs = bias_s + tile_par * BN;
            if (a.bias != nullptr && tid < BN) bs[tid] = wk.n0 + tid < a.N ? __ldg(a.bias + wk.n0 + tid) : 0.f;
            if (BPACK && MASK == 0 && a.filt_thr != nullptr && tid < BN) bs[tid] = wk.n0 + tid < a.N ? __ldg(a.filt_thr + wk.n0 + tid) : INFINITY;
            uint32_t mbits[HALF / 32];
            if (MASK == 2) {
                const uint32_t* mrow = a.mask + row * a.ldm + col0 / 32;
#pragma unroll
                for (int q = 0; q < HALF / 32; ++q) mbits[q] = (row_ok && col0 + 32 * q < a.N) ? __ldg(mrow + q) : 0u;
            }
            for (int c = 0; c < n_chunks; ++c, ++gc) {
                const uint32_t buf = gc & 1;
                mbar_wait_timed(tfull0 + 8 * buf, (gc >> 1) & 1, tid == 0 ? a.dbg : nullptr, d_tfull);
                tc_fence_after();
#pragma unroll
                for (int q = 0; q < HALF / 16; ++q) {
                    uint32_t v[16];
                    tmem_ld16(tlane + buf * BN + q * 16, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[q * 16 + j] += __uint_as_float(v[j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG2 && cta_rank != 0) mbar_arrive_remote(tempty0 + 8 * buf, 0);  // the MMA thread lives in the leader
                    else mbar_arrive(tempty0 + 8 * buf);
                }
            }
            // ---- epilogue on this thread's [row, n0 + half*HALF .. +HALF); the MMA warp is already on the next tile.
            // (Measured with ps_gemm_tc_trace: the accumulate warps spent 60-85 % of the kernel here and the MMA thread
            // 37-56 % waiting for them to free an accumulator, so this code is kept lean: nothing per element that the
            // call does not ask for, shared-memory accesses as LDS/STS, store mode chosen once per tile.)
            const long long d_e0 = (a.dbg && tid == 0) ? clock64() : 0;
            if (BPACK && MASK == 0 && a.filt_thr != nullptr) {
                // kNN candidate filter: nothing is stored but the (few) elements that reach their column's threshold
                asm volatile("bar.sync 1, 256;" ::: "memory");  // bs holds the thresholds of this tile's columns
                if (row_ok) {
                    const float* th = bs + half * HALF;
#pragma unroll
                    for (int j = 0; j < HALF; ++j) {
                        if (acc[j] >= th[j]) {
                            const int64_t col = col0 + j;
                            const int slot = atomicAdd(a.filt_cnt + col, 1);
                            if (slot < a.filt_cap) {
                                a.filt_val[col * a.filt_cap + slot] = acc[j];
                                a.filt_row[col * a.filt_cap + slot] = static_cast<int32_t>(row);
                            }
                        }
                    }
                }
                continue;
            }
            if (a.bias != nullptr) {
                asm volatile("bar.sync 1, 256;" ::: "memory");  // bs was filled before the drain
                const float4* b4 = reinterpret_cast<const float4*>(bs + half * HALF);
#pragma unroll
                for (int j = 0; j < HALF / 4; ++j) {
                    const float4 b = b4[j];
                    acc[4 * j] += b.x; acc[4 * j + 1] += b.y; acc[4 * j + 2] += b.z; acc[4 * j + 3] += b.w;
                }
            }
            if (a.act == 1) {
#pragma unroll
                for (int j = 0; j < HALF; ++j) acc[j] = fmaxf(acc[j], acc[j] * PS_LEAKY_SLOPE);  // leaky_relu, slope < 1
            }
            if (MASK == 1) {  // record sign(activation): the backward applies leaky' from these bits
                uint32_t* mrow = a.mask + row * a.ldm + col0 / 32;
#pragma unroll
                for (int q = 0; q < HALF / 32; ++q) {
                    uint32_t bits = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) bits |= (acc[32 * q + j] > 0.f ? 1u : 0u) << j;
                    if (row_ok && col0 + 32 * q < a.N) mrow[q] = bits;
                }
            } else if (MASK == 2) {  // act == 2: sum * leaky'(y), y > 0 recorded by the forward (words loaded before the drain)
#pragma unroll
                for (int q = 0; q < HALF / 32; ++q)
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[32 * q + j] *= ((mbits[q] >> j) & 1u) ? 1.f : PS_LEAKY_SLOPE;
            }
            float inv_norm = 1.f;
            if (a.l2norm) {  // the row is split over two threads (column halves): combine through shared memory
                float ss = 0.f;
#pragma unroll
                for (int j = 0; j < HALF; ++j) ss = fmaf(acc[j], (col0 + j < a.N) ? acc[j] : 0.f, ss);
                float* sb = ss_buf + tile_par * 256;
                sb[half * 128 + quarter * 32 + lane] = ss;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const float nrm = sqrtf(sb[quarter * 32 + lane] + sb[128 + quarter * 32 + lane]);
                inv_norm = 1.f / nrm;
                if (half == 0 && row_ok && a.norm_out) a.norm_out[row] = nrm;
#pragma unroll
                for (int j = 0; j < HALF; ++j) acc[j] *= inv_norm;
            }
            const long long d_e1 = (a.dbg && tid == 0) ? clock64() : 0;
            if (a.exp & 4) continue;
            // Stores go through a per-warp shared-memory transpose, 16 columns at a time: a thread owns one output
            // row, and writing it directly would touch 32 rows x 16 B per instruction (half sectors, 32 LSU
            // wavefronts).  Staged, an instruction writes 8 rows x 64 contiguous bytes (full sectors).
            if (CG2 && BPACK) {
                // The tile leaves through the TMA engine.  A thread owns one output row: per round it writes 32 columns of it
                // into its row of a [32 x 128 B] SWIZZLE_128B box (chunk c at c ^ (row % 8): conflict-free), lane 0 hands the
                // box to cp.async.bulk.tensor, and the warp goes on with the other box while the engine drains this one.
                // Two boxes per warp = half of the tile in flight when the warp returns to draining accumulators: the store
                // traffic (128 KB per tile at the ~30 GB/s one SM gets of the chip's write bandwidth) overlaps the next
                // tile's MMAs instead of holding the accumulator registers (measured: the register -> STG epilogue was 30 %
                // of the kernel, profiles/r2q_gemm_ablation.txt).
                const int row_base = static_cast<int>(wk.m0) + quarter * 32;
                uint8_t* boxes = tbox + warp * 8192;
#pragma unroll
                for (int r = 0; r < HALF / 32; ++r) {
                    uint8_t* box = boxes + (r & 1) * 4096;
                    if (lane == 0) bulk_wait_read<1>();  // the store that last read this box (two rounds ago) is done with it
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<float4*>(box + lane * 128 + ((c ^ (lane & 7)) << 4)) =
                            make_float4(acc[32 * r + 4 * c], acc[32 * r + 4 * c + 1], acc[32 * r + 4 * c + 2], acc[32 * r + 4 * c + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_box(&cmap, smem_u32(box), static_cast<int>(col0) + 32 * r, row_base);
                        bulk_commit();
                    }
                }
            } else
            {
                float* wb = epi_buf + warp * (32 * kEpiStride);
                const int rr = lane >> 2, cc = (lane & 3) * 4;
                const int64_t row_base = wk.m0 + quarter * 32;
                const int mode = a.accumulate ? 1 : ((MASK == 0 && a.act == 2) ? 2 : 0);  // uniform over the tile
                float* dst0 = a.C + (row_base + rr) * a.ldc + col0 + cc;
                const int64_t step8 = 8 * a.ldc;
                bool ok[4];
#pragma unroll
                for (int p8 = 0; p8 < 4; ++p8) ok[p8] = row_base + 8 * p8 + rr < a.M;
#pragma unroll
                for (int blk = 0; blk < HALF / 16; ++blk) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<float4*>(wb + lane * kEpiStride + 4 * q) =
                            make_float4(acc[16 * blk + 4 * q], acc[16 * blk + 4 * q + 1], acc[16 * blk + 4 * q + 2], acc[16 * blk + 4 * q + 3]);
                    __syncwarp();
                    const bool col_ok = col0 + 16 * blk + cc < a.N;  // N % 4 == 0: a float4 is all-or-nothing
                    float* dst = dst0 + 16 * blk;
                    float4 v[4];
#pragma unroll
                    for (int p8 = 0; p8 < 4; ++p8) v[p8] = *reinterpret_cast<const float4*>(wb + (8 * p8 + rr) * kEpiStride + cc);
                    if (mode == 0) {
#pragma unroll
                        for (int p8 = 0; p8 < 4; ++p8)
                            if (ok[p8] && col_ok) *reinterpret_cast<float4*>(dst + p8 * step8) = v[p8];
                    } else if (mode == 1) {  // split-K partials
#pragma unroll
                        for (int p8 = 0; p8 < 4; ++p8)
                            if (ok[p8] && col_ok) {
                                float* d = dst + p8 * step8;
                                atomicAdd(d + 0, v[p8].x); atomicAdd(d + 1, v[p8].y); atomicAdd(d + 2, v[p8].z); atomicAdd(d + 3, v[p8].w);
                            }
                    } else {  // C holds leaky_relu outputs y: C = acc * leaky'(y)
#pragma unroll
                        for (int p8 = 0; p8 < 4; ++p8)
                            if (ok[p8] && col_ok) {
                                float* d = dst + p8 * step8;
                                const float4 y = *reinterpret_cast<const float4*>(d);
                                *reinterpret_cast<float4*>(d) = make_float4(v[p8].x * ps_leaky_grad_from_out(y.x), v[p8].y * ps_leaky_grad_from_out(y.y),
                                                                            v[p8].z * ps_leaky_grad_from_out(y.z), v[p8].w * ps_leaky_grad_from_out(y.w));
                            }
                    }
                    __syncwarp();
                }
            }
            if (a.dbg && tid == 0) { const long long d_e2 = clock64(); d_epi += d_e2 - d_e0; d_store += d_e2 - d_e1; }
        }
        if (a.dbg && tid == 0) { a.dbg[blockIdx.x * 8 + 5] = d_tfull; a.dbg[blockIdx.x * 8 + 6] = d_epi; a.dbg[blockIdx.x * 8 + 7] = d_store; }
        if (CG2 && BPACK && lane == 0) bulk_wait<0>();  // the last boxes have left shared memory (and landed) before the CTA retires
        tc_fence_before();
    }
    __syncthreads();
    if (CL > 1) {  // nobody leaves while the peer may still signal its barriers or write its stages
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (warp == 16) {
        tc_fence_after();
        if (CG2) tmem_dealloc2<2 * BN>(tmem_base);
        else tmem_dealloc<2 * BN>(tmem_base);
    }
}

// Split a weight operand into hi/lo and write the K-major SWIZZLE_128B tile images the BPACK kernels stream:
// block (n-tile, k-block) -> [hi: BN rows x 128 B][lo: BN rows x 128 B]; row j = output column n0 + j, 16-byte
// chunk c (k0 + 4c .. +3) stored at c ^ (j % 8), 8-row groups 1024 B apart.  Works from either source layout.
template <int BN>
__global__ void __launch_bounds__(256) pack_b_kernel(const float* __restrict__ Q, int64_t ldq, int q_kmajor, int64_t N, int64_t K,
                                                     int64_t nkb, uint8_t* __restrict__ out) {
    const int64_t tile = blockIdx.x;
    const int64_t n0 = (tile / nkb) * BN, k0 = (tile % nkb) * BK;
    uint8_t* hi = out + tile * (2 * BN * 128);
    uint8_t* lo = hi + BN * 128;
    for (int g = threadIdx.x; g < BN * 8; g += 256) {
        const int j = q_kmajor ? g >> 3 : g % BN;
        const int c = q_kmajor ? g & 7 : g / BN;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        const int64_t n = n0 + j;
        if (n < N) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int64_t k = k0 + 4 * c + e;
                if (k < K) v[e] = __ldg(q_kmajor ? Q + n * ldq + k : Q + k * ldq + n);
            }
        }
        const uint32_t off = (j >> 3) * 1024 + (j & 7) * 128 + ((c ^ (j & 7)) << 4);
        split_store(hi, lo, off, make_float4(v[0], v[1], v[2], v[3]));
    }
}

// Scratch for the packed weight images: one buffer per (device, stream), grown on demand and reused by every later
// call on that stream (stream order makes the reuse safe; cudaFree of an outgrown buffer waits for its readers).
// cudaMallocAsync per call was measured to stall the host inside the training step.
static int g_tc_waves = 1;
static int g_tc_reserve = 0;

struct PackSlot { int dev; cudaStream_t stream; void* ptr; size_t bytes; };
static PackSlot g_pack_slots[16] = {};

static std::mutex g_pack_mutex;

static int pack_scratch(cudaStream_t stream, size_t bytes, void** out) {
    std::lock_guard<std::mutex> lock(g_pack_mutex);  // several host threads may run GEMMs (one stream each)
    int dev = 0;
    PS_CUDA_CHECK(cudaGetDevice(&dev));
    PackSlot* slot = nullptr;
    for (auto& s : g_pack_slots)
        if (s.ptr != nullptr && s.dev == dev && s.stream == stream) { slot = &s; break; }
    if (slot == nullptr)
        for (auto& s : g_pack_slots)
            if (s.ptr == nullptr) { slot = &s; break; }
    if (slot == nullptr) slot = &g_pack_slots[0];  // more than 16 (device, stream) pairs: recycle
    if (slot->ptr == nullptr || slot->dev != dev || slot->stream != stream || slot->bytes < bytes) {
        if (slot->ptr != nullptr) PS_CUDA_CHECK(cudaFree(slot->ptr));
        slot->ptr = nullptr;
        const size_t want = bytes < (4u << 20) ? (4u << 20) : bytes;
        PS_CUDA_CHECK(cudaMalloc(&slot->ptr, want));
        slot->dev = dev; slot->stream = stream; slot->bytes = want;
    }
    *out = slot->ptr;
    return PS_OK;
}

// 2-D tensor map of the output C [M, N] (row pitch ldc floats) for the TMA store of [32 x 32] fp32 boxes, SWIZZLE_128B
int make_c_map(CUtensorMap* map, float* C, int64_t M, int64_t N, int64_t ldc) {
    using Encode = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Encode encode = nullptr;
    if (encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        PS_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (fn == nullptr || q != cudaDriverEntryPointSuccess) return ps_fail(PS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        encode = reinterpret_cast<Encode>(fn);
    }
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(M)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ldc) * sizeof(float)};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, C, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return ps_fail(PS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(rc));
    return PS_OK;
}

template <bool PK, bool QK, int BN, bool BPACK, int MASK = 0, int CL = 1, bool CG2 = false>
int launch_tc(TcArgs a, int q_kmajor, cudaStream_t stream) {
    constexpr int STAGES = (CG2 && !BPACK) ? 3 : (BN == 256 ? 2 : 3);
    // CG2: half-width Q tiles per stage.  Packed-weight kernels: behind the barriers / bias the 64 KB of TMA store boxes replace the
    // 20 KB staging buffer; weight-gradient kernel: three stages and the staging buffer (its partials leave as atomics)
    constexpr size_t kMisc = (2 * STAGES + 4) * 8 + 16 + (2 * 2 * 128 + 2 * BN + 8 * 32 * kEpiStride) * 4 + 1024;
    constexpr size_t smem = CG2 ? (BPACK ? STAGES * (2 * BM * 128 + BN * 128) + 8192 + 8 * 8192 + 1024 : STAGES * (2 * BM * 128 + BN * 128) + kMisc)
                                : STAGES * (2 * BM * 128 + 2 * BN * 128) + kMisc;
    static_assert(smem <= 232448, "shared memory budget");
    auto kern = gemm_tc_kernel<PK, QK, BN, BPACK, MASK, CL, CG2>;
    alignas(64) CUtensorMap cmap;
    memset(&cmap, 0, sizeof(cmap));
    if (CG2 && a.C != nullptr) {
        const int rc = make_c_map(&cmap, a.C, a.M, a.N, a.ldc);
        if (rc != PS_OK) return rc;
    }
    // per DEVICE: a process may drive several devices through ps_set_device (function attributes and the SM count belong to one)
    constexpr int kMaxDevices = 64;
    static bool configured[kMaxDevices] = {};
    static int sm_count[kMaxDevices] = {};
    int dev = 0;
    PS_CUDA_CHECK(cudaGetDevice(&dev));
    const int slot = dev >= 0 && dev < kMaxDevices ? dev : 0;
    if (!configured[slot] || slot != dev) {
        PS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        PS_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count[slot], cudaDevAttrMultiProcessorCount, dev));
        configured[slot] = true;
    }
    const int sms = sm_count[slot];
    void* pack = nullptr;
    if (BPACK) {
        const int64_t nkb = ps_ceil_div(a.K, BK);
        const size_t bytes = static_cast<size_t>(a.nt * nkb) * 2 * BN * 128;
        const int rc = pack_scratch(stream, bytes, &pack);
        if (rc != PS_OK) return rc;
        pack_b_kernel<BN><<<static_cast<unsigned>(a.nt * nkb), 256, 0, stream>>>(a.Q, a.ldq, q_kmajor, a.N, a.K, nkb, static_cast<uint8_t*>(pack));
        PS_LAUNCH_CHECK();
        a.bpack = static_cast<const uint8_t*>(pack);
    }
    // Persistent CTAs, but bounded: with >= 8 work items per SM the grid is g_tc_waves x SMs, so every CTA retires
    // after 1/g_tc_waves of its SM's share and a pending higher-priority CTA (the side-stream batch preparation of the
    // next training step) gets the SM within a fraction of the GEMM instead of after all of it.
    const int64_t total = a.mt * a.nt * a.zs;
    const int64_t usable = sms - g_tc_reserve > 1 ? sms - g_tc_reserve : 1;  // SMs left free for other streams
    int64_t grid64 = total < usable ? total : usable;
    if (g_tc_waves > 1 && total >= static_cast<int64_t>(sms) * 8 * g_tc_waves) grid64 = static_cast<int64_t>(sms) * g_tc_waves;
    if (CL > 1) {  // one cluster per tile pair: an even grid, at most one CTA per SM
        const int64_t pairs = ps_ceil_div(a.mt, CL) * a.nt * (BPACK ? 1 : a.zs);
        int64_t clusters = usable / CL;
        if (clusters > pairs) clusters = pairs;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(static_cast<unsigned>(clusters * CL));
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        PS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, a, cmap));
        return PS_OK;
    }
    const unsigned grid = static_cast<unsigned>(grid64);
    kern<<<grid, kThreads, smem, stream>>>(a, cmap);
    PS_LAUNCH_CHECK();
    return PS_OK;
}

}  // namespace

extern "C" int ps_gemm_tc_reserve_sms(int n) { const int old = g_tc_reserve; if (n >= 0 && n <= 64) g_tc_reserve = n; return old; }
extern "C" int ps_gemm_tc_waves(int waves) { const int old = g_tc_waves; if (waves >= 1 && waves <= 64) g_tc_waves = waves; return old; }
extern "C" int ps_gemm_tc_prefetch(int on) {
    int old = 0;
    if (cudaMemcpyFromSymbol(&old, g_tc_prefetch, sizeof(int)) != cudaSuccess) return PS_ERR_CUDA;
    const int v = on != 0;
    if (cudaMemcpyToSymbol(g_tc_prefetch, &v, sizeof(int)) != cudaSuccess) return PS_ERR_CUDA;
    return old;
}
static unsigned long long* g_tc_dbg = nullptr;
static int g_tc_exp = 0;
extern "C" int ps_gemm_tc_experiment(int bits) { const int old = g_tc_exp; g_tc_exp = bits; return old; }
// development: buf = device array of 8 counters per CTA (148 x 8), filled by the next tensor-core GEMM launches:
// [0] MMA thread total cycles, [1] its wait for operands (full), [2] its wait for a free TMEM buffer, [3] weight-stream wait
// for a free stage, [4] producer wait for a free stage, [5] accumulate-warp wait for a finished chunk, [6] epilogue cycles
extern "C" int ps_gemm_tc_trace(unsigned long long* buf) { g_tc_dbg = buf; return PS_OK; }
static bool g_tc_pack = true;
static int g_tc_cluster = 3;  // 0 = single CTAs, 1 = tile pairs with the weight image multicast, 2 = tile pairs on one cta_group::2 MMA, 3 = 2 + the weight-gradient GEMMs on pairs as well
extern "C" int ps_gemm_tc_cluster(int mode) { const int old = g_tc_cluster; if (mode >= 0 && mode <= 3) g_tc_cluster = mode; return old; }
extern "C" int ps_gemm_tc_pack(int on) { const int old = g_tc_pack; if (on == 0 || on == 1) g_tc_pack = on != 0; return old; }

// Returns PS_ERR_UNSUPPORTED (without setting an error) when the shape is outside what this
// path covers; the dispatcher then uses the CUDA-core kernel.
int ps_gemm_tc_launch(const float* P, int64_t ldp, int p_kmajor, const int32_t* p_rows,
                      const float* Q, int64_t ldq, int q_kmajor, const int32_t* q_rows,
                      float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                      const float* bias, int act, int l2norm, float* norm_out, int accumulate, int splits,
                      uint32_t* mask, int64_t ldm, float* p_colsum, cudaStream_t stream) {
    if (M <= 0 || N < 64 || K < 32) return PS_ERR_UNSUPPORTED;
    if (p_colsum != nullptr && (p_kmajor || q_kmajor || !accumulate)) return PS_ERR_UNSUPPORTED;  // weight-gradient form only                 // tiny problems: not worth a 128-wide tile
    if (mask != nullptr && (N % 32 != 0 || (act != 1 && act != 2) || l2norm)) return PS_ERR_UNSUPPORTED;
    if ((ldp | ldq | ldc | N) % 4 != 0) return PS_ERR_UNSUPPORTED;
    if ((p_kmajor || q_kmajor) && K % 4 != 0) return PS_ERR_UNSUPPORTED;
    if (!p_kmajor && M % 4 != 0) return PS_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(C)) % 16 != 0) return PS_ERR_UNSUPPORTED;
    if (l2norm && N > 256) return PS_ERR_UNSUPPORTED;
    if (q_kmajor && q_rows) return PS_ERR_UNSUPPORTED;  // row pointers of Q are not precomputed
    if (accumulate && (bias || act || l2norm)) return PS_ERR_UNSUPPORTED;
    if (act == 2 && (bias || l2norm)) return PS_ERR_UNSUPPORTED;
    if (splits > 1 && !accumulate) return PS_ERR_UNSUPPORTED;
    // The tensor core's TMEM accumulator truncates on every add (measured: ~2e-8 relative per MMA, biased
    // towards zero), so a long accumulation chain drifts: K = 10 000 in one CTA gave 7e-5.  Keep chains at
    // <= kMaxChainK (error ~1.5e-5): accumulate-mode calls get more split-K CTAs (partials are combined with
    // round-to-nearest fp32 atomics), plain calls with a longer K go to the CUDA-core kernel.
    constexpr int64_t kMaxChainK = 2048;
    if (!accumulate && K > 4 * kMaxChainK) return PS_ERR_UNSUPPORTED;
    if (accumulate && ps_ceil_div(K, splits < 1 ? 1 : splits) > kMaxChainK) splits = static_cast<int>(ps_ceil_div(K, kMaxChainK));
    const int BN = (N > 128) ? 256 : 128;
    TcArgs a{P, ldp, p_rows, Q, ldq, q_rows, C, ldc, M, N, K, bias, norm_out, act, l2norm, accumulate, 0, 0, 0, 0, nullptr, mask, ldm, g_tc_dbg, g_tc_exp,
             nullptr, nullptr, nullptr, nullptr, 0, p_colsum};
    const int64_t num_kb = ps_ceil_div(K, BK);
    if (splits < 1) splits = 1;
    a.kb_per_split = static_cast<int>(ps_ceil_div(num_kb, splits));
    a.zs = ps_ceil_div(num_kb, a.kb_per_split);
    a.mt = ps_ceil_div(M, BM);
    a.nt = ps_ceil_div(N, BN);
    // weight operand (no gather, no split-K) against a tall K-major activation: pre-packed hi/lo images + bulk copies
    const bool packable = p_kmajor && q_rows == nullptr && !accumulate && a.zs == 1;
    // tile PAIRS (2-CTA clusters sharing the weight stream by multicast) once there are enough tiles to fill the SMs with pairs
    const bool pairs = g_tc_cluster != 0 && a.mt * a.nt >= 2 * 148;
    const bool cg2 = pairs && g_tc_cluster >= 2 && !(mask == nullptr && act == 2);  // the cta_group::2 kernels store through TMA: no read-modify-write form
#define PS_TC_PACKED(MASKV)                                                                                                              \
    do {                                                                                                                                 \
        if (cg2) return BN == 256 ? launch_tc<true, true, 256, true, MASKV, 2, true>(a, q_kmajor, stream)                                \
                                  : launch_tc<true, true, 128, true, MASKV, 2, true>(a, q_kmajor, stream);                               \
        if (pairs) return BN == 256 ? launch_tc<true, true, 256, true, MASKV, 2>(a, q_kmajor, stream)                                    \
                                    : launch_tc<true, true, 128, true, MASKV, 2>(a, q_kmajor, stream);                                   \
        return BN == 256 ? launch_tc<true, true, 256, true, MASKV>(a, q_kmajor, stream) : launch_tc<true, true, 128, true, MASKV>(a, q_kmajor, stream); \
    } while (0)
    if (mask != nullptr) {  // the sign-mask epilogue lives in the packed-weight kernels only
        if (!packable) return PS_ERR_UNSUPPORTED;
        if (act == 1) PS_TC_PACKED(1);
        PS_TC_PACKED(2);
    }
    if (g_tc_pack && packable && M >= 1024) PS_TC_PACKED(0);
#undef PS_TC_PACKED
    if (g_tc_cluster == 3 && !p_kmajor && !q_kmajor && accumulate && BN == 256 && a.mt >= 2 && a.mt * a.nt * a.zs >= 148)
        return launch_tc<false, false, 256, false, 0, 2, true>(a, q_kmajor, stream);  // weight gradients on tile pairs (cta_group::2)
#define PS_TC_CASE(pk, qk)                                                         \
    if (static_cast<bool>(p_kmajor) == pk && static_cast<bool>(q_kmajor) == qk)     \
        return BN == 256 ? launch_tc<pk, qk, 256, false>(a, q_kmajor, stream) : launch_tc<pk, qk, 128, false>(a, q_kmajor, stream);
    PS_TC_CASE(true, true)
    PS_TC_CASE(true, false)
    PS_TC_CASE(false, true)
    PS_TC_CASE(false, false)
#undef PS_TC_CASE
    return PS_ERR_UNSUPPORTED;
}

// kNN candidate filter on the packed-weight kernels: for every (i, j) with sum_r P[i,r] Q[j,r] >= thr[j], append (value, i)
// to list j.  P [M, K] (the embedding table) streams through the producers, Q [N, K] (the query tile) is the packed
// "weight" operand; the similarity tile is never written.
extern "C" int ps_gemm_filter(const float* P, int64_t ldp, const float* Q, int64_t ldq, int64_t M, int64_t N, int64_t K,
                              const float* thr, int32_t* cnt, float* cand_val, int32_t* cand_row, int cap, ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(P && Q && thr && cnt && cand_val && cand_row && cap > 0, "null pointer");
    if (M < 1024 || N < 64 || K < 32 || K > 8192 || (ldp | ldq | N | K) % 4 != 0 ||
        (reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(Q)) % 16 != 0)
        return ps_fail(PS_ERR_UNSUPPORTED, "ps_gemm_filter: shape outside the packed tensor-core path");
    const int BN = (N > 128) ? 256 : 128;
    TcArgs a{P, ldp, nullptr, Q, ldq, nullptr, nullptr, 0, M, N, K, nullptr, nullptr, 0, 0, 0, 0, 0, 0, 0, nullptr, nullptr, 0, g_tc_dbg, 0,
             thr, cnt, cand_val, cand_row, cap, nullptr};
    const int64_t num_kb = ps_ceil_div(K, BK);
    a.kb_per_split = static_cast<int>(num_kb);
    a.zs = 1;
    a.mt = ps_ceil_div(M, BM);
    a.nt = ps_ceil_div(N, BN);
    const bool pairs = g_tc_cluster != 0 && a.mt * a.nt >= 2 * 148;
    // the filter epilogue stores nothing, so the TMA store boxes of the cta_group::2 kernels buy nothing here: multicast pairs (measured 2.98 vs 3.16 ms)
    if (pairs) return BN == 256 ? launch_tc<true, true, 256, true, 0, 2>(a, 1, stream) : launch_tc<true, true, 128, true, 0, 2>(a, 1, stream);
    return BN == 256 ? launch_tc<true, true, 256, true>(a, 1, stream) : launch_tc<true, true, 128, true>(a, 1, stream);
}
