// Generic fp32 contraction  C[i,j] (+)= act(sum_r P(i,r) Q(j,r) + bias[j])  on the CUDA
// cores: the exact-fp32 path used for the backward weight gradients (contraction over the
// row dimension, operands "MN-major") and as the parity anchor / fallback shape coverage
// for the tcgen05 kernel in gemm_tc.cu.  128x128x16 tiles, 256 threads, 8x8 outputs per
// thread, register-prefetch double buffering, row gather folded into the operand loads.
#include "common.cuh"
#include "../../include/pinsage_b200.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, LDS_ = BM + 4, THREADS = 256;

struct GemmArgs {
    const float* P; int64_t ldp; const int32_t* p_rows;
    const float* Q; int64_t ldq; const int32_t* q_rows;
    float* C; int64_t ldc;
    int64_t M, N, K;
    const float* bias; float* norm_out;
    int act, l2norm, accumulate;
    int64_t k_per_split;
};

// Fetch this thread's share (2 x float4) of a [128 x 16] operand tile into registers.
// KMAJOR : element(i, r) = X[row(i)*ld + r]  -> thread covers rows {t/4, t/4+64}, r4 = (t%4)*4
// MNMAJOR: element(i, r) = X[row(r)*ld + i]  -> thread covers r {t/32, t/32+8},   i4 = (t%32)*4
template <bool KMAJOR>
__device__ __forceinline__ void fetch_tile(const float* __restrict__ X, int64_t ld, const int32_t* __restrict__ rows,
                                           int64_t i0, int64_t ext_i, int64_t r0, int64_t r_end, int tid, float4 (&v)[2]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        v[h] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (KMAJOR) {
            const int64_t i = i0 + (tid >> 2) + h * 64;
            const int64_t r = r0 + (tid & 3) * 4;
            if (i < ext_i && r < r_end) {
                const int64_t row = rows ? static_cast<int64_t>(__ldg(rows + i)) : i;
                v[h] = ps_ldg4(X + row * ld + r);
            }
        } else {
            const int64_t r = r0 + (tid >> 5) + h * 8;
            const int64_t i = i0 + (tid & 31) * 4;
            if (i < ext_i && r < r_end) {
                const int64_t row = rows ? static_cast<int64_t>(__ldg(rows + r)) : r;
                v[h] = ps_ldg4(X + row * ld + i);
            }
        }
    }
}

template <bool KMAJOR>
__device__ __forceinline__ void stash_tile(float (*S)[LDS_], int tid, const float4 (&v)[2]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (KMAJOR) {
            const int i = (tid >> 2) + h * 64, r = (tid & 3) * 4;
            S[r + 0][i] = v[h].x; S[r + 1][i] = v[h].y; S[r + 2][i] = v[h].z; S[r + 3][i] = v[h].w;
        } else {
            const int r = (tid >> 5) + h * 8, i = (tid & 31) * 4;
            *reinterpret_cast<float4*>(&S[r][i]) = v[h];
        }
    }
}

template <bool PK, bool QK>
__global__ void __launch_bounds__(THREADS, 2) gemm_simt_kernel(GemmArgs a) {
    __shared__ __align__(16) float As[2][BK][LDS_];
    __shared__ __align__(16) float Bs[2][BK][LDS_];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t bm = static_cast<int64_t>(blockIdx.x) * BM, bn = static_cast<int64_t>(blockIdx.y) * BN;
    const int64_t k_begin = static_cast<int64_t>(blockIdx.z) * a.k_per_split;
    const int64_t k_end = min(a.K, k_begin + a.k_per_split);
    if (k_begin >= k_end) return;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 pa[2], pb[2];
    fetch_tile<PK>(a.P, a.ldp, a.p_rows, bm, a.M, k_begin, k_end, tid, pa);
    fetch_tile<QK>(a.Q, a.ldq, a.q_rows, bn, a.N, k_begin, k_end, tid, pb);
    stash_tile<PK>(As[0], tid, pa);
    stash_tile<QK>(Bs[0], tid, pb);
    __syncthreads();

    int buf = 0;
    for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
        const bool more = k0 + BK < k_end;
        if (more) {
            fetch_tile<PK>(a.P, a.ldp, a.p_rows, bm, a.M, k0 + BK, k_end, tid, pa);
            fetch_tile<QK>(a.Q, a.ldq, a.q_rows, bn, a.N, k0 + BK, k_end, tid, pb);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            stash_tile<PK>(As[buf ^ 1], tid, pa);
            stash_tile<QK>(Bs[buf ^ 1], tid, pb);
        }
        __syncthreads();
        buf ^= 1;
    }

    // ---- epilogue ----
    float bcol[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int64_t col = bn + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
        bcol[j] = (a.bias != nullptr && col < a.N) ? __ldg(a.bias + col) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t row = bm + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = acc[i][j] + bcol[j];
            if (a.act == 1) v = ps_leaky(v);
            acc[i][j] = v;
            const int64_t col = bn + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (col < a.N) ss = fmaf(v, v, ss);
        }
        if (a.l2norm) {  // the 16 threads sharing `ty` hold the whole row (N <= 128)
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            const float nrm = sqrtf(ss);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = acc[i][j] / nrm;
            if (tx == 0 && row < a.M && a.norm_out != nullptr) a.norm_out[row] = nrm;
        }
        if (row < a.M) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t col = bn + h * 64 + tx * 4;
                if (col < a.N) {  // N % 4 == 0 is required, so the float4 is all-or-nothing
                    float* dst = a.C + row * a.ldc + col;
                    if (a.accumulate) {
                        atomicAdd(dst + 0, acc[i][h * 4 + 0]); atomicAdd(dst + 1, acc[i][h * 4 + 1]);
                        atomicAdd(dst + 2, acc[i][h * 4 + 2]); atomicAdd(dst + 3, acc[i][h * 4 + 3]);
                    } else if (a.act == 2) {  // C holds leaky_relu outputs y: C = acc * leaky'(y)
                        const float4 y = *reinterpret_cast<const float4*>(dst);
                        *reinterpret_cast<float4*>(dst) = make_float4(acc[i][h * 4 + 0] * ps_leaky_grad_from_out(y.x), acc[i][h * 4 + 1] * ps_leaky_grad_from_out(y.y),
                                                                      acc[i][h * 4 + 2] * ps_leaky_grad_from_out(y.z), acc[i][h * 4 + 3] * ps_leaky_grad_from_out(y.w));
                    } else {
                        *reinterpret_cast<float4*>(dst) = make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
                    }
                }
            }
        }
    }
}

}  // namespace

int ps_gemm_simt_launch(const float* P, int64_t ldp, int p_kmajor, const int32_t* p_rows,
                        const float* Q, int64_t ldq, int q_kmajor, const int32_t* q_rows,
                        float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                        const float* bias, int act, int l2norm, float* norm_out, int accumulate, int splits,
                        cudaStream_t stream) {
    PS_REQUIRE(P && Q && C, "null operand");
    PS_REQUIRE(M >= 0 && N > 0 && K > 0, "bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    if (M == 0) return PS_OK;
    PS_REQUIRE(ldp % 4 == 0 && ldq % 4 == 0 && ldc % 4 == 0 && N % 4 == 0, "leading dimensions and N must be multiples of 4");
    PS_REQUIRE((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(C)) % 16 == 0, "operands must be 16-byte aligned");
    if (p_kmajor || q_kmajor) PS_REQUIRE(K % 4 == 0, "K must be a multiple of 4 for K-major operands");
    if (!p_kmajor) PS_REQUIRE(M % 4 == 0, "M must be a multiple of 4 for an MN-major P");
    PS_REQUIRE(!l2norm || N <= BN, "l2norm epilogue needs N <= 128");
    PS_REQUIRE(splits >= 1, "splits must be >= 1");
    PS_REQUIRE(!accumulate || (!bias && !act && !l2norm), "accumulate excludes bias/act/l2norm");
    PS_REQUIRE(act >= 0 && act <= 2 && (act != 2 || (!bias && !l2norm)), "act must be 0, 1 or 2 (2 excludes bias/l2norm)");
    PS_REQUIRE(splits == 1 || accumulate, "split-K needs accumulate");
    GemmArgs a{P, ldp, p_rows, Q, ldq, q_rows, C, ldc, M, N, K, bias, norm_out, act, l2norm, accumulate, 0};
    int64_t kps = ps_ceil_div(ps_ceil_div(K, splits), BK) * BK;
    a.k_per_split = kps;
    const int64_t zs = ps_ceil_div(K, kps);
    dim3 grid(static_cast<unsigned>(ps_ceil_div(M, BM)), static_cast<unsigned>(ps_ceil_div(N, BN)), static_cast<unsigned>(zs));
    PS_REQUIRE(grid.y <= 65535u && grid.z <= 65535u, "grid too large");
    if (p_kmajor && q_kmajor) gemm_simt_kernel<true, true><<<grid, THREADS, 0, stream>>>(a);
    else if (p_kmajor && !q_kmajor) gemm_simt_kernel<true, false><<<grid, THREADS, 0, stream>>>(a);
    else if (!p_kmajor && q_kmajor) gemm_simt_kernel<false, true><<<grid, THREADS, 0, stream>>>(a);
    else gemm_simt_kernel<false, false><<<grid, THREADS, 0, stream>>>(a);
    PS_LAUNCH_CHECK();
    return PS_OK;
}
