// K4 + K6 + K11: neighbour gather / importance-weighted aggregation fused with the
// concat, its backward as a segmented gather, and the small row-wise backward kernels of
// ConvLayer (reference pinsage_model.py:189-212).  All HBM-bound: one warp per row,
// 128-bit coalesced loads of whole feature rows, several rows in flight per lane.
#include "common.cuh"
#include "../../include/pinsage_b200.h"

namespace {

constexpr int kWarps = 8;  // warps per CTA for the row-per-warp kernels

// cat[i, :din] = hin[self_rows[i], :din];  cat[i, din:] = sum_t w[i,t] z[nbz[i,t], :] / sum_t w[i,t]
// CH = number of float4 chunks per lane covering dh (dh <= CH * 128)
template <int CH>
__global__ void __launch_bounds__(kWarps * 32)
aggregate_fwd_kernel(const float* __restrict__ hin, int64_t ld_hin, const int32_t* __restrict__ self_rows, int din,
                     const float* __restrict__ z, int64_t ldz, const int32_t* __restrict__ nbz,
                     const float* __restrict__ nbw, int T, int dh, int64_t n,
                     float* __restrict__ cat, int64_t ldcat, float* __restrict__ inv_wsum) {
    const int lane = threadIdx.x & 31;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
    if (i >= n) return;
    const int32_t* nb = nbz + i * T;
    const float* w = nbw + i * T;
    float* out = cat + i * ldcat;

    // self row copy (the concat)
    const float* self = hin + static_cast<int64_t>(__ldg(self_rows + i)) * ld_hin;
    for (int c = lane * 4; c < din; c += 128) *reinterpret_cast<float4*>(out + c) = ps_ldg4(self + c);

    float4 acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    float wsum = 0.f;
    constexpr int U = 4;  // neighbour rows in flight per lane
    int t = 0;
    for (; t + U <= T; t += U) {
        int32_t r[U]; float wt[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { r[u] = __ldg(nb + t + u); wt[u] = __ldg(w + t + u); }
        float4 v[U][CH];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const int col = (c * 32 + lane) * 4;
                v[u][c] = col < dh ? ps_ldg4(z + static_cast<int64_t>(r[u]) * ldz + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            wsum += wt[u];
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                acc[c].x = fmaf(wt[u], v[u][c].x, acc[c].x); acc[c].y = fmaf(wt[u], v[u][c].y, acc[c].y);
                acc[c].z = fmaf(wt[u], v[u][c].z, acc[c].z); acc[c].w = fmaf(wt[u], v[u][c].w, acc[c].w);
            }
        }
    }
    for (; t < T; ++t) {
        const int32_t r = __ldg(nb + t);
        const float wt = __ldg(w + t);
        wsum += wt;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int col = (c * 32 + lane) * 4;
            if (col < dh) {
                const float4 v = ps_ldg4(z + static_cast<int64_t>(r) * ldz + col);
                acc[c].x = fmaf(wt, v.x, acc[c].x); acc[c].y = fmaf(wt, v.y, acc[c].y);
                acc[c].z = fmaf(wt, v.z, acc[c].z); acc[c].w = fmaf(wt, v.w, acc[c].w);
            }
        }
    }
    const float inv = 1.f / wsum;  // 0-weight rows give inf/NaN exactly like the reference's 0/0 (:202)
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int col = (c * 32 + lane) * 4;
        if (col < dh)
            *reinterpret_cast<float4*>(out + din + col) = make_float4(acc[c].x * inv, acc[c].y * inv, acc[c].z * inv, acc[c].w * inv);
    }
    if (lane == 0) inv_wsum[i] = inv;
}

// Backward of the aggregation as a load-balanced segmented gather.  Z-row u receives
//   z[u,:] = leaky'(z[u,:]) * sum_{q in seg(u)} w[q] * inv_wsum[q/T] * dcat[q/T, col_off:col_off+dh]
// Popular nodes own segments of 10^4+ pairs while most rows own a handful, so the work unit is a CHUNK of at
// most `chunk_pairs` consecutive pairs of one segment (one warp each, chunk_off[u] = first chunk of row u).
// A row with a single chunk is finished in place; a row with several chunks writes raw partial sums to a
// scratch buffer that aggregate_bwd_reduce_kernel then folds (no atomics: the result is deterministic).
template <int CH>
__global__ void __launch_bounds__(kWarps * 32)
aggregate_bwd_kernel(const float* __restrict__ dcat, int64_t ldcat, int col_off, int dh,
                     const int32_t* __restrict__ seg_off, const int32_t* __restrict__ chunk_off, int chunk_pairs,
                     const int32_t* __restrict__ pair_q, const float* __restrict__ nbw,
                     const float* __restrict__ inv_wsum, int T,
                     float* __restrict__ z, int64_t ldz, int64_t n_zrows, float* __restrict__ partial,
                     const int32_t* __restrict__ chunk_row, int apply_leaky) {
    const int lane = threadIdx.x & 31;
    const int64_t chunk = static_cast<int64_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
    if (chunk >= __ldg(chunk_off + n_zrows)) return;
    // row that owns this chunk: the last u with chunk_off[u] <= chunk (precomputed by the caller, else searched)
    int64_t lo = 0, hi = n_zrows;
    if (chunk_row != nullptr) {
        lo = __ldg(chunk_row + chunk);
    } else {
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(chunk_off + mid) <= chunk) lo = mid; else hi = mid;
        }
    }
    const int64_t u = lo;
    const int first = __ldg(chunk_off + u), n_chunks = __ldg(chunk_off + u + 1) - first;
    const int beg = __ldg(seg_off + u) + static_cast<int>(chunk - first) * chunk_pairs;
    const int end = min(__ldg(seg_off + u + 1), beg + chunk_pairs);
    float4 acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    // The per-pair scalars (pair -> target row needs an integer division by T, two dependent 4-byte gathers give the
    // coefficient) are computed ONE PAIR PER LANE, 32 pairs at a time, and handed round by shuffle: computed by every lane for
    // every pair they were 45 of the kernel's 55 warp-instructions per pair (ncu), for one 128-bit load and four FMAs of
    // payload at dh = 128.  The order of the sum is unchanged (pairs ascending): results are bit-identical.
    for (int p0 = beg; p0 < end; p0 += 32) {
        const int np = min(32, end - p0);
        int my_row = 0;
        float my_coef = 0.f;
        if (lane < np) {
            const int q = __ldg(pair_q + p0 + lane);
            my_row = q / T;
            my_coef = __ldg(nbw + q) * __ldg(inv_wsum + my_row);
        }
#pragma unroll 4
        for (int k = 0; k < np; ++k) {
            const int64_t row = __shfl_sync(0xffffffffu, my_row, k);
            const float coef = __shfl_sync(0xffffffffu, my_coef, k);
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const int col = (c * 32 + lane) * 4;
                if (col < dh) {
                    const float4 v = ps_ldg4(dcat + row * ldcat + col_off + col);
                    acc[c].x = fmaf(coef, v.x, acc[c].x); acc[c].y = fmaf(coef, v.y, acc[c].y);
                    acc[c].z = fmaf(coef, v.z, acc[c].z); acc[c].w = fmaf(coef, v.w, acc[c].w);
                }
            }
        }
    }
    // z rows and partials stream through (read once / written once): evict-first keeps L2 for the dcat rows,
    // which every z-row of a target's neighbourhood re-reads
    if (n_chunks == 1) {
        float* zr = z + u * ldz;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int col = (c * 32 + lane) * 4;
            if (col < dh) {
                if (apply_leaky) {
                    const float4 y = __ldcs(reinterpret_cast<const float4*>(zr + col));
                    __stcs(reinterpret_cast<float4*>(zr + col),
                           make_float4(acc[c].x * ps_leaky_grad_from_out(y.x), acc[c].y * ps_leaky_grad_from_out(y.y),
                                       acc[c].z * ps_leaky_grad_from_out(y.z), acc[c].w * ps_leaky_grad_from_out(y.w)));
                } else {
                    *reinterpret_cast<float4*>(zr + col) = acc[c];
                }
            }
        }
    } else {
        float* pr = partial + chunk * dh;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int col = (c * 32 + lane) * 4;
            if (col < dh) *reinterpret_cast<float4*>(pr + col) = acc[c];
        }
    }
}

// Second pass: rows with no pair get a zero gradient, rows split over several chunks get the sum of their
// partials (in chunk order) times leaky'(z).  Rows with exactly one chunk were finished by the first pass.
template <int CH>
__global__ void __launch_bounds__(kWarps * 32)
aggregate_bwd_reduce_kernel(const int32_t* __restrict__ chunk_off, const float* __restrict__ partial, int dh,
                            float* __restrict__ z, int64_t ldz, int64_t n_zrows, int apply_leaky) {
    const int lane = threadIdx.x & 31;
    const int64_t u = static_cast<int64_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
    if (u >= n_zrows) return;
    const int first = __ldg(chunk_off + u), n_chunks = __ldg(chunk_off + u + 1) - first;
    if (n_chunks == 1) return;
    float4 acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < n_chunks; ++k) {
        const float* pr = partial + static_cast<int64_t>(first + k) * dh;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int col = (c * 32 + lane) * 4;
            if (col < dh) {
                const float4 v = ps_ldg4(pr + col);
                acc[c].x += v.x; acc[c].y += v.y; acc[c].z += v.z; acc[c].w += v.w;
            }
        }
    }
    float* zr = z + u * ldz;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int col = (c * 32 + lane) * 4;
        if (col < dh) {
            if (apply_leaky) {
                const float4 y = *reinterpret_cast<const float4*>(zr + col);
                *reinterpret_cast<float4*>(zr + col) =
                    make_float4(acc[c].x * ps_leaky_grad_from_out(y.x), acc[c].y * ps_leaky_grad_from_out(y.y),
                                acc[c].z * ps_leaky_grad_from_out(y.z), acc[c].w * ps_leaky_grad_from_out(y.w));
            } else {
                *reinterpret_cast<float4*>(zr + col) = acc[c];
            }
        }
    }
}

// dpre = leaky'(h) * (dh - h (h.dh)) / norm, one warp per row
__global__ void __launch_bounds__(kWarps * 32)
norm_leaky_bwd_kernel(const float* __restrict__ h, int64_t ldh, const float* __restrict__ norm,
                      const float* __restrict__ dh, int64_t lddh, float* __restrict__ dpre, int64_t ldp,
                      int64_t n, int d) {
    const int lane = threadIdx.x & 31;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
    if (i >= n) return;
    const float* hr = h + i * ldh;
    const float* gr = dh + i * lddh;
    float dot = 0.f;
    for (int c = lane * 4; c < d; c += 128) {
        const float4 a = ps_ldg4(hr + c), g = ps_ldg4(gr + c);
        dot += a.x * g.x + a.y * g.y + a.z * g.z + a.w * g.w;
    }
    dot = ps_warp_sum(dot);
    const float inv = 1.f / __ldg(norm + i);
    for (int c = lane * 4; c < d; c += 128) {
        const float4 a = ps_ldg4(hr + c), g = ps_ldg4(gr + c);
        float4 o;
        o.x = ps_leaky_grad_from_out(a.x) * (g.x - a.x * dot) * inv;
        o.y = ps_leaky_grad_from_out(a.y) * (g.y - a.y * dot) * inv;
        o.z = ps_leaky_grad_from_out(a.z) * (g.z - a.z * dot) * inv;
        o.w = ps_leaky_grad_from_out(a.w) * (g.w - a.w * dot) * inv;
        *reinterpret_cast<float4*>(dpre + i * ldp + c) = o;
    }
}

__global__ void leaky_bwd_kernel(const float* __restrict__ y, float* __restrict__ dy, int64_t n4) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(y) + i);
        float4 g = reinterpret_cast<float4*>(dy)[i];
        g.x *= ps_leaky_grad_from_out(a.x); g.y *= ps_leaky_grad_from_out(a.y);
        g.z *= ps_leaky_grad_from_out(a.z); g.w *= ps_leaky_grad_from_out(a.w);
        reinterpret_cast<float4*>(dy)[i] = g;
    }
}

// out[j] += sum_i x[i, j]; block = 32 column-lanes x 8 row-groups, grid.x tiles rows, grid.y tiles columns
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, int64_t ld, int64_t n, int d, int64_t rows_per_block, float* __restrict__ out) {
    __shared__ float part[8][33];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int col = blockIdx.y * 32 + lx;
    const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
    const int64_t r1 = min(n, r0 + rows_per_block);
    float s = 0.f;
    if (col < d)
        for (int64_t r = r0 + ly; r < r1; r += 8) s += __ldg(x + r * ld + col);
    part[ly][lx] = s;
    __syncthreads();
    if (ly == 0 && col < d) {
#pragma unroll
        for (int k = 1; k < 8; ++k) s += part[k][lx];
        atomicAdd(out + col, s);
    }
}

__global__ void __launch_bounds__(kWarps * 32)
scatter_add_rows_kernel(const float* __restrict__ src, int64_t lds, const int32_t* __restrict__ rows,
                        float* __restrict__ dst, int64_t ldd, int64_t n, int d) {
    const int lane = threadIdx.x & 31;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
    if (i >= n) return;
    const float* s = src + i * lds;
    float* t = dst + static_cast<int64_t>(__ldg(rows + i)) * ldd;
    for (int c = lane * 4; c < d; c += 128) {
        const float4 a = ps_ldg4(s + c);
        float4 b = *reinterpret_cast<float4*>(t + c);
        b.x += a.x; b.y += a.y; b.z += a.z; b.w += a.w;
        *reinterpret_cast<float4*>(t + c) = b;
    }
}

// x[i,:] /= ||x[i,:]||, norm_out[i] = ||x[i,:]||  (used when out_dim > 128, where the GEMM epilogue cannot see a whole row)
__global__ void __launch_bounds__(kWarps * 32)
l2norm_rows_kernel(float* __restrict__ x, int64_t ld, int64_t n, int d, float* __restrict__ norm_out) {
    const int lane = threadIdx.x & 31;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
    if (i >= n) return;
    float* r = x + i * ld;
    float ss = 0.f;
    for (int c = lane * 4; c < d; c += 128) {
        const float4 a = *reinterpret_cast<const float4*>(r + c);
        ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    const float nrm = sqrtf(ps_warp_sum(ss));
    for (int c = lane * 4; c < d; c += 128) {
        float4 a = *reinterpret_cast<float4*>(r + c);
        a.x /= nrm; a.y /= nrm; a.z /= nrm; a.w /= nrm;
        *reinterpret_cast<float4*>(r + c) = a;
    }
    if (lane == 0 && norm_out != nullptr) norm_out[i] = nrm;
}

}  // namespace

#define PS_DISPATCH_CH(dh, CALL)                                            \
    do {                                                                    \
        if ((dh) <= 128) { constexpr int CH = 1; CALL; }                    \
        else if ((dh) <= 256) { constexpr int CH = 2; CALL; }               \
        else if ((dh) <= 512) { constexpr int CH = 4; CALL; }               \
        else { constexpr int CH = 8; CALL; }                                \
    } while (0)

extern "C" int ps_aggregate_fwd(const float* hin, int64_t ld_hin, const int32_t* self_rows, int din,
                                const float* z, int64_t ldz, const int32_t* nbz, const float* nbw, int T, int dh,
                                int64_t n, float* cat, int64_t ldcat, float* inv_wsum, ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(hin && self_rows && z && nbz && nbw && cat && inv_wsum, "null pointer");
    PS_REQUIRE(din > 0 && dh > 0 && T > 0 && n >= 0, "bad shape");
    PS_REQUIRE(din % 4 == 0 && dh % 4 == 0 && ld_hin % 4 == 0 && ldz % 4 == 0 && ldcat % 4 == 0, "dims and leading dimensions must be multiples of 4");
    PS_REQUIRE(dh <= 1024, "hidden dim > 1024 not supported");
    PS_REQUIRE(ldcat >= din + dh, "cat row too short");
    if (n == 0) return PS_OK;
    const unsigned blocks = static_cast<unsigned>(ps_ceil_div(n, kWarps));
    PS_DISPATCH_CH(dh, (aggregate_fwd_kernel<CH><<<blocks, kWarps * 32, 0, stream>>>(
                           hin, ld_hin, self_rows, din, z, ldz, nbz, nbw, T, dh, n, cat, ldcat, inv_wsum)));
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_aggregate_bwd(const float* dcat, int64_t ldcat, int col_off, int dh, const int32_t* seg_off,
                                const int32_t* chunk_off, int chunk_pairs, int64_t max_chunks,
                                const int32_t* pair_q, const float* nbw, const float* inv_wsum, int T,
                                float* z, int64_t ldz, int64_t n_zrows, float* partial_ws, const int32_t* chunk_row,
                                int apply_leaky, ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(dcat && seg_off && chunk_off && pair_q && nbw && inv_wsum && z && partial_ws, "null pointer");
    PS_REQUIRE(dh > 0 && dh % 4 == 0 && col_off % 4 == 0 && ldcat % 4 == 0 && ldz % 4 == 0 && T > 0, "bad shape");
    PS_REQUIRE(dh <= 1024, "hidden dim > 1024 not supported");
    PS_REQUIRE(chunk_pairs > 0 && max_chunks >= 0, "bad chunking");
    if (n_zrows == 0) return PS_OK;
    if (max_chunks > 0) {
        const unsigned blocks = static_cast<unsigned>(ps_ceil_div(max_chunks, kWarps));
        PS_DISPATCH_CH(dh, (aggregate_bwd_kernel<CH><<<blocks, kWarps * 32, 0, stream>>>(
                               dcat, ldcat, col_off, dh, seg_off, chunk_off, chunk_pairs, pair_q, nbw, inv_wsum, T, z, ldz,
                               n_zrows, partial_ws, chunk_row, apply_leaky)));
        PS_LAUNCH_CHECK();
    }
    const unsigned rblocks = static_cast<unsigned>(ps_ceil_div(n_zrows, kWarps));
    PS_DISPATCH_CH(dh, (aggregate_bwd_reduce_kernel<CH><<<rblocks, kWarps * 32, 0, stream>>>(chunk_off, partial_ws, dh, z, ldz, n_zrows, apply_leaky)));
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_norm_leaky_bwd(const float* h, int64_t ldh, const float* norm, const float* dh, int64_t lddh,
                                 float* dpre, int64_t ldp, int64_t n, int d, ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(h && norm && dh && dpre, "null pointer");
    PS_REQUIRE(d > 0 && d % 4 == 0 && ldh % 4 == 0 && lddh % 4 == 0 && ldp % 4 == 0, "bad shape");
    if (n == 0) return PS_OK;
    norm_leaky_bwd_kernel<<<static_cast<unsigned>(ps_ceil_div(n, kWarps)), kWarps * 32, 0, stream>>>(h, ldh, norm, dh, lddh, dpre, ldp, n, d);
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_leaky_bwd(const float* y, float* dy, int64_t n_elems, ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(y && dy, "null pointer");
    PS_REQUIRE(n_elems % 4 == 0, "element count must be a multiple of 4");
    if (n_elems == 0) return PS_OK;
    const int64_t n4 = n_elems / 4;
    int64_t blocks = ps_ceil_div(n4, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    leaky_bwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(y, dy, n4);
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_colsum(const float* x, int64_t ld, int64_t n, int d, float* out, ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(x && out, "null pointer");
    PS_REQUIRE(d > 0, "bad shape");
    if (n == 0) return PS_OK;
    const int col_tiles = static_cast<int>(ps_ceil_div(d, 32));
    int64_t row_blocks = ps_ceil_div(148 * 8, col_tiles);
    if (row_blocks > ps_ceil_div(n, 64)) row_blocks = ps_ceil_div(n, 64);
    if (row_blocks < 1) row_blocks = 1;
    const int64_t rpb = ps_ceil_div(n, row_blocks);
    dim3 grid(static_cast<unsigned>(ps_ceil_div(n, rpb)), static_cast<unsigned>(col_tiles));
    colsum_kernel<<<grid, 256, 0, stream>>>(x, ld, n, d, rpb, out);
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_scatter_add_rows(const float* src, int64_t lds, const int32_t* rows, float* dst, int64_t ldd,
                                   int64_t n, int d, ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(src && rows && dst, "null pointer");
    PS_REQUIRE(d > 0 && d % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "bad shape");
    if (n == 0) return PS_OK;
    scatter_add_rows_kernel<<<static_cast<unsigned>(ps_ceil_div(n, kWarps)), kWarps * 32, 0, stream>>>(src, lds, rows, dst, ldd, n, d);
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_l2norm_rows(float* x, int64_t ld, int64_t n, int d, float* norm_out, ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(x != nullptr, "null pointer");
    PS_REQUIRE(d > 0 && d % 4 == 0 && ld % 4 == 0, "bad shape");
    if (n == 0) return PS_OK;
    l2norm_rows_kernel<<<static_cast<unsigned>(ps_ceil_div(n, kWarps)), kWarps * 32, 0, stream>>>(x, ld, n, d, norm_out);
    PS_LAUNCH_CHECK();
    return PS_OK;
}
