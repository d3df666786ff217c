// tcgen05 (5th-gen tensor core) path of ps_gemm for sm_100a:
//     C[i,j] (+)= act( sum_r P(i,r) Q(j,r) + bias[j] )      fp32 in, fp32 out
// computed as an error-compensated 3xTF32 product (hi/lo split of both operands,
// a_lo*b_hi + a_hi*b_lo + a_hi*b_hi accumulated in fp32 in TMEM), which keeps the
// reference's fp32 parity (<= 1e-4) that plain TF32 cannot.
//
// Structure of one CTA (one 128 x BN output tile, BN = 128 or 256, optional split-K):
//   warps 8-11 producers: LDG.128 the operand rows (row gather folded in, K4), split each
//              fp32 into hi/lo, STS.128 into the UMMA canonical shared-memory layout
//              (SWIZZLE_128B K-major or SWIZZLE_128B_BASE32B MN-major), fence.proxy.async,
//              arrive on the stage's mbarrier;
//   warp 12    lane 0 issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8), 12 per
//              32-wide k-block, into one of two TMEM accumulators; tcgen05.commit frees the
//              smem stage and, every 2 k-blocks, publishes the accumulator;
//   warps 0-7  tcgen05.ld the finished accumulator (32x32b, one output row x BN/2 columns per
//              thread) and add it to fp32 registers -- the tensor core's TMEM adder truncates
//              (measured ~2e-8 relative per MMA, biased), so the truncating chain is kept to
//              24 MMAs and the long sum is round-to-nearest; then the epilogue from registers:
//              bias + leaky_relu + row L2-normalise and store, or red.global.add for split-K
//              weight gradients.  Registers are rebalanced with setmaxnreg (184 / 88 / 40).
// Replaces nn.Linear / AddmmBackward of ConvLayer and the head (pinsage_model.py:201,
// 208-210, 259).
#include "common.cuh"
#include "../../include/pinsage_b200.h"

namespace {

constexpr int BM = 128;                 // UMMA M
constexpr int BK = 32;                  // floats per k-block = one 128-byte swizzle row
constexpr int kProducerThreads = 128;
constexpr int kThreads = 512;
constexpr uint32_t kHiMask = 0xFFFFE000u;  // keep sign, exponent and the 10 tf32 mantissa bits

struct TcArgs {
    const float* P; int64_t ldp; const int32_t* p_rows;
    const float* Q; int64_t ldq; const int32_t* q_rows;
    float* C; int64_t ldc;
    int64_t M, N, K;
    const float* bias; float* norm_out;
    int act, l2norm, accumulate;
    int kb_per_split;  // k-blocks (of BK) per blockIdx.z
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
        if (clock64() - t0 > 4000000000ll) { printf("ps_gemm_tc: mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x); __trap(); }
    }
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
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

// Fill one [ROWS x 32] operand tile (hi and lo copies) for the k-block starting at k0.
//   KMAJOR : element(i, r) = X[row(i)*ld + r]; SWIZZLE_128B: smem row i = 128 B, 16-B chunk c stored at c ^ (i % 8),
//            8-row groups 1024 B apart (SBO)
//   MNMAJOR: element(i, r) = X[row(r)*ld + i]; SWIZZLE_128B_BASE32B: atoms of 4 k-rows x 32 i (512 B), 32-B chunk c
//            of k-row r stored at c ^ (r % 4); atoms of consecutive i-chunks 512 B apart (LBO), 4-row k-groups
//            (ROWS/32)*512 B apart (SBO)
// Loads are issued in batches of 8 x LDG.128 per thread before the first use (memory-level parallelism).
template <bool KMAJOR, int ROWS>
__device__ __forceinline__ void fill_tile(uint8_t* hi, uint8_t* lo, const float* __restrict__ X, int64_t ld,
                                          const int32_t* __restrict__ rows, int64_t i0, int64_t ext_i,
                                          int64_t k0, int64_t k_end, int t) {
    constexpr int BATCH = 8;
    if (KMAJOR) {
        const int c = t & 7;
        const int64_t k = k0 + c * 4;
#pragma unroll
        for (int p0 = 0; p0 < ROWS / 16; p0 += BATCH) {
            float4 v[BATCH];
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                const int64_t i = i0 + (p0 + b) * 16 + (t >> 3);
                v[b] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < ext_i && k < k_end) {
                    const int64_t row = rows ? static_cast<int64_t>(__ldg(rows + i)) : i;
                    v[b] = ps_ldg4(X + row * ld + k);
                }
            }
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                const int r = (p0 + b) * 16 + (t >> 3);
                split_store(hi, lo, (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4), v[b]);
            }
        }
    } else {
        constexpr int C4 = ROWS / 4;               // float4 per k-row
        constexpr int RPP = kProducerThreads / C4;  // k-rows per pass
        const int c4 = t % C4;
        const int64_t i = i0 + c4 * 4;
#pragma unroll
        for (int p0 = 0; p0 < BK / RPP; p0 += BATCH) {
            float4 v[BATCH];
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                const int64_t k = k0 + (p0 + b) * RPP + t / C4;
                v[b] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < ext_i && k < k_end) {
                    const int64_t row = rows ? static_cast<int64_t>(__ldg(rows + k)) : k;
                    v[b] = ps_ldg4(X + row * ld + i);
                }
            }
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                const int kr = (p0 + b) * RPP + t / C4;
                split_store(hi, lo, (kr >> 2) * (ROWS / 32) * 512 + (c4 >> 3) * 512 + (kr & 3) * 128 +
                                        ((((c4 & 7) >> 1) ^ (kr & 3)) << 5) + ((c4 & 1) << 4), v[b]);
            }
        }
    }
}

// Thread roles (16 warps): 0-7 accumulate + epilogue, 8-11 producers, 12 MMA issue (13-15 only give their
// registers away: setmaxnreg works on whole warpgroups).
template <bool PK, bool QK, int BN>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(TcArgs a) {
    constexpr int STAGES = BN == 256 ? 2 : 3;
    constexpr int HALF = BN / 2;            // columns owned by one epilogue thread
    constexpr int CHUNK_KB = 2;             // k-blocks accumulated in TMEM before the fp32 register drain
    constexpr uint32_t A_BYTES = BM * 128, B_BYTES = BN * 128;
    constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    float* ss_buf = reinterpret_cast<float*>(tmem_slot + 4);  // [2][128] partial sums of squares (l2norm)
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const uint32_t tfull0 = smem_u32(bars + 2 * STAGES), tempty0 = smem_u32(bars + 2 * STAGES + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = static_cast<int64_t>(blockIdx.x) * BM, n0 = static_cast<int64_t>(blockIdx.y) * BN;
    const int64_t num_kb_total = (a.K + BK - 1) / BK;
    const int64_t kb_begin = static_cast<int64_t>(blockIdx.z) * a.kb_per_split;
    const int64_t kb_end = min(num_kb_total, kb_begin + a.kb_per_split);
    const int num_kb = static_cast<int>(kb_end - kb_begin);
    if (num_kb <= 0) return;  // uniform per CTA
    const int n_chunks = (num_kb + CHUNK_KB - 1) / CHUNK_KB;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, kProducerThreads); mbar_init(empty0 + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, 8); }
        fence_barrier_init();
    }
    if (warp == 12) tmem_alloc<2 * BN>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 12) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        // ------------------------------------------------ MMA issue (one lane of warp 12)
        if (warp == 12 && lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((PK ? 0u : 1u) << 15) | ((QK ? 0u : 1u) << 16) |
                                       (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
            constexpr uint32_t A_LBO = PK ? 16 : 512, A_SBO = PK ? 1024 : (BM / 32) * 512, A_STEP = PK ? 32 : 2 * (BM / 32) * 512;
            constexpr uint32_t B_LBO = QK ? 16 : 512, B_SBO = QK ? 1024 : (BN / 32) * 512, B_STEP = QK ? 32 : 2 * (BN / 32) * 512;
            constexpr uint32_t A_LAY = PK ? 2u : 1u, B_LAY = QK ? 2u : 1u;
            int stage = 0; uint32_t phase = 0;
            int kb = 0;
            for (int c = 0; c < n_chunks; ++c) {
                const int buf = c & 1;
                mbar_wait(tempty0 + 8 * buf, ((c >> 1) & 1) ^ 1);  // the drain of this buffer's previous chunk is done
                tc_fence_after();
                const uint32_t tacc = tmem_base + buf * BN;
                for (int q = 0; q < CHUNK_KB && kb < num_kb; ++q, ++kb) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t a_hi = sa, a_lo = sa + A_BYTES, b_hi = sa + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                    for (int s = 0; s < BK / 8; ++s) {
                        const uint64_t dah = make_desc(a_hi + s * A_STEP, A_LBO, A_SBO, A_LAY), dal = make_desc(a_lo + s * A_STEP, A_LBO, A_SBO, A_LAY);
                        const uint64_t dbh = make_desc(b_hi + s * B_STEP, B_LBO, B_SBO, B_LAY), dbl = make_desc(b_lo + s * B_STEP, B_LBO, B_SBO, B_LAY);
                        umma_tf32(tacc, dal, dbh, idesc, (q | s) != 0);  // a chunk starts a fresh accumulator; small terms first
                        umma_tf32(tacc, dah, dbl, idesc, 1u);
                        umma_tf32(tacc, dah, dbh, idesc, 1u);
                    }
                    umma_commit(empty0 + 8 * stage);  // frees the smem stage once these MMAs have read it
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull0 + 8 * buf);  // publishes the chunk to the accumulate warps
            }
        }
    } else if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        // ------------------------------------------------ producers
        const int t = tid - 8 * 32;
        int stage = 0; uint32_t phase = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            uint8_t* st = smem + stage * STAGE_BYTES;
            const int64_t k0 = (kb_begin + kb) * BK;
            fill_tile<PK, BM>(st, st + A_BYTES, a.P, a.ldp, a.p_rows, m0, a.M, k0, a.K, t);
            fill_tile<QK, BN>(st + 2 * A_BYTES, st + 2 * A_BYTES + B_BYTES, a.Q, a.ldq, a.q_rows, n0, a.N, k0, a.K, t);
            fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core's async proxy
            mbar_arrive(full0 + 8 * stage);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
        // ------------------------------------------------ accumulate (fp32, round-to-nearest, in registers) + epilogue
        // Each chunk of CHUNK_KB k-blocks (24 MMAs) is summed by the tensor core in TMEM, whose adder truncates;
        // draining it into registers keeps the truncating chain short (fp32-level accuracy).
        const int quarter = warp & 3, half = warp >> 2;
        const int64_t row = m0 + quarter * 32 + lane;
        const uint32_t tlane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + half * HALF;
        float acc[HALF];
#pragma unroll
        for (int j = 0; j < HALF; ++j) acc[j] = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
            const int buf = c & 1;
            mbar_wait(tfull0 + 8 * buf, (c >> 1) & 1);
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
            if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
        }
        // epilogue on this thread's [row, n0 + half*HALF .. +HALF)
        const int64_t col0 = n0 + half * HALF;
        const bool row_ok = row < a.M;
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
            const int64_t col = col0 + j;
            if (col < a.N) {
                float x = acc[j] + (a.bias ? __ldg(a.bias + col) : 0.f);
                if (a.act == 1) x = ps_leaky(x);
                acc[j] = x;
                ss = fmaf(x, x, ss);
            }
        }
        float inv_norm = 1.f;
        if (a.l2norm) {  // the row is split over two threads (column halves): combine through shared memory
            ss_buf[half * 128 + quarter * 32 + lane] = ss;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float nrm = sqrtf(ss_buf[quarter * 32 + lane] + ss_buf[128 + quarter * 32 + lane]);
            inv_norm = 1.f / nrm;
            if (half == 0 && row_ok && a.norm_out) a.norm_out[row] = nrm;
        }
        if (row_ok) {
            float* dst = a.C + row * a.ldc + col0;
#pragma unroll
            for (int j = 0; j < HALF; j += 4) {
                if (col0 + j < a.N) {  // N % 4 == 0: a float4 is all-or-nothing
                    if (a.accumulate) {
                        atomicAdd(dst + j + 0, acc[j + 0]); atomicAdd(dst + j + 1, acc[j + 1]);
                        atomicAdd(dst + j + 2, acc[j + 2]); atomicAdd(dst + j + 3, acc[j + 3]);
                    } else {
                        *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j] * inv_norm, acc[j + 1] * inv_norm,
                                                                          acc[j + 2] * inv_norm, acc[j + 3] * inv_norm);
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 12) { tc_fence_after(); tmem_dealloc<2 * BN>(tmem_base); }
}

template <bool PK, bool QK, int BN>
int launch_tc(const TcArgs& a, dim3 grid, cudaStream_t stream) {
    constexpr int STAGES = BN == 256 ? 2 : 3;
    constexpr size_t smem = STAGES * (2 * BM * 128 + 2 * BN * 128) + (2 * STAGES + 4) * 8 + 16 + 2 * 128 * 4 + 1024;
    auto kern = gemm_tc_kernel<PK, QK, BN>;
    static bool configured = false;
    if (!configured) {
        PS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        configured = true;
    }
    kern<<<grid, kThreads, smem, stream>>>(a);
    PS_LAUNCH_CHECK();
    return PS_OK;
}

}  // namespace

// Returns PS_ERR_UNSUPPORTED (without setting an error) when the shape is outside what this
// path covers; the dispatcher then uses the CUDA-core kernel.
int ps_gemm_tc_launch(const float* P, int64_t ldp, int p_kmajor, const int32_t* p_rows,
                      const float* Q, int64_t ldq, int q_kmajor, const int32_t* q_rows,
                      float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                      const float* bias, int act, int l2norm, float* norm_out, int accumulate, int splits,
                      cudaStream_t stream) {
    if (M <= 0 || N < 64 || K < 32) return PS_ERR_UNSUPPORTED;                 // tiny problems: not worth a 128-wide tile
    if ((ldp | ldq | ldc | N) % 4 != 0) return PS_ERR_UNSUPPORTED;
    if ((p_kmajor || q_kmajor) && K % 4 != 0) return PS_ERR_UNSUPPORTED;
    if (!p_kmajor && M % 4 != 0) return PS_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(C)) % 16 != 0) return PS_ERR_UNSUPPORTED;
    if (l2norm && N > 256) return PS_ERR_UNSUPPORTED;
    if (accumulate && (bias || act || l2norm)) return PS_ERR_UNSUPPORTED;
    if (splits > 1 && !accumulate) return PS_ERR_UNSUPPORTED;
    // The tensor core's TMEM accumulator truncates on every add (measured: ~2e-8 relative per MMA, biased
    // towards zero), so a long accumulation chain drifts: K = 10 000 in one CTA gave 7e-5.  Keep chains at
    // <= kMaxChainK (error ~1.5e-5): accumulate-mode calls get more split-K CTAs (partials are combined with
    // round-to-nearest fp32 atomics), plain calls with a longer K go to the CUDA-core kernel.
    constexpr int64_t kMaxChainK = 2048;
    if (!accumulate && K > 4 * kMaxChainK) return PS_ERR_UNSUPPORTED;
    if (accumulate && ps_ceil_div(K, splits < 1 ? 1 : splits) > kMaxChainK) splits = static_cast<int>(ps_ceil_div(K, kMaxChainK));
    const int BN = (N > 128) ? 256 : 128;
    TcArgs a{P, ldp, p_rows, Q, ldq, q_rows, C, ldc, M, N, K, bias, norm_out, act, l2norm, accumulate, 0};
    const int64_t num_kb = ps_ceil_div(K, BK);
    if (splits < 1) splits = 1;
    a.kb_per_split = static_cast<int>(ps_ceil_div(num_kb, splits));
    const int64_t zs = ps_ceil_div(num_kb, a.kb_per_split);
    dim3 grid(static_cast<unsigned>(ps_ceil_div(M, BM)), static_cast<unsigned>(ps_ceil_div(N, BN)), static_cast<unsigned>(zs));
    if (grid.y > 65535u || grid.z > 65535u) return PS_ERR_UNSUPPORTED;
#define PS_TC_CASE(pk, qk)                                                         \
    if (static_cast<bool>(p_kmajor) == pk && static_cast<bool>(q_kmajor) == qk)     \
        return BN == 256 ? launch_tc<pk, qk, 256>(a, grid, stream) : launch_tc<pk, qk, 128>(a, grid, stream);
    PS_TC_CASE(true, true)
    PS_TC_CASE(true, false)
    PS_TC_CASE(false, true)
    PS_TC_CASE(false, false)
#undef PS_TC_CASE
    return PS_ERR_UNSUPPORTED;
}
