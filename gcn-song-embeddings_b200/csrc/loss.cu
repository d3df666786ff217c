// K10: max-margin loss forward + backward with the (q, pos, neg) row gather and the
// gradient scatter-add fused; K13: Adam on a flat buffer.
// Reference: max_margin_loss (pinsage_training.py:31-41), train_batch (:184-191).
#include "common.cuh"
#include "../../include/pinsage_b200.h"

namespace {

constexpr int kWarps = 4;
constexpr int kMaxChunks = 8;  // d <= 1024

__global__ void count_triples_kernel(const int32_t* __restrict__ triples, int64_t B, int64_t U, int32_t* __restrict__ counts) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= B * 3) return;
    const int col = static_cast<int>(i % 3);
    atomicAdd(counts + col * U + __ldg(triples + i), 1);
}

// One warp per (q, pos, neg) triple.
__global__ void __launch_bounds__(kWarps * 32)
margin_loss_kernel(const float* __restrict__ emb, int64_t ld, const int32_t* __restrict__ triples, int64_t B, int d,
                   float margin, float grad_scale, const int32_t* __restrict__ dup_counts, int64_t U,
                   float* __restrict__ loss_out, float* __restrict__ demb, int64_t ldd) {
    const int lane = threadIdx.x & 31;
    const int64_t b = static_cast<int64_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
    if (b >= B) return;
    const int32_t iq = __ldg(triples + b * 3 + 0), ip = __ldg(triples + b * 3 + 1), in_ = __ldg(triples + b * 3 + 2);
    const float* q = emb + static_cast<int64_t>(iq) * ld;
    const float* p = emb + static_cast<int64_t>(ip) * ld;
    const float* ng = emb + static_cast<int64_t>(in_) * ld;
    float4 vq[kMaxChunks], vp[kMaxChunks], vn[kMaxChunks];
    float qq = 0.f, pp = 0.f, nn = 0.f, qp = 0.f, qn = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
        const int col = (c * 32 + lane) * 4;
        if (col < d) {
            vq[c] = ps_ldg4(q + col); vp[c] = ps_ldg4(p + col); vn[c] = ps_ldg4(ng + col);
            qq += vq[c].x * vq[c].x + vq[c].y * vq[c].y + vq[c].z * vq[c].z + vq[c].w * vq[c].w;
            pp += vp[c].x * vp[c].x + vp[c].y * vp[c].y + vp[c].z * vp[c].z + vp[c].w * vp[c].w;
            nn += vn[c].x * vn[c].x + vn[c].y * vn[c].y + vn[c].z * vn[c].z + vn[c].w * vn[c].w;
            qp += vq[c].x * vp[c].x + vq[c].y * vp[c].y + vq[c].z * vp[c].z + vq[c].w * vp[c].w;
            qn += vq[c].x * vn[c].x + vq[c].y * vn[c].y + vq[c].z * vn[c].z + vq[c].w * vn[c].w;
        }
    }
    qq = ps_warp_sum(qq); pp = ps_warp_sum(pp); nn = ps_warp_sum(nn); qp = ps_warp_sum(qp); qn = ps_warp_sum(qn);
    // F.normalize: x / max(||x||, 1e-12)
    const float nq = sqrtf(qq), np_ = sqrtf(pp), nn_ = sqrtf(nn);
    const float rq = 1.f / fmaxf(nq, 1e-12f), rp = 1.f / fmaxf(np_, 1e-12f), rn = 1.f / fmaxf(nn_, 1e-12f);
    const float cp = qp * rq * rp, cn = qn * rq * rn;  // q^.p^, q^.n^
    const float dsum = cn - cp + margin;
    const bool active = dsum >= 0.f;  // torch.max over the stacked pair sends a tie to the first argument
    if (lane == 0 && active) atomicAdd(loss_out, dsum / static_cast<float>(B));
    if (demb == nullptr || !active) return;
    const float g = grad_scale / static_cast<float>(B);
    float kq = 1.f, kp = 1.f, kn = 1.f;
    if (dup_counts != nullptr) {
        kq = static_cast<float>(__ldg(dup_counts + 0 * U + iq));
        kp = static_cast<float>(__ldg(dup_counts + 1 * U + ip));
        kn = static_cast<float>(__ldg(dup_counts + 2 * U + in_));
    }
    // clamp-aware normalize backward: if ||x|| <= eps the normalisation is x/eps (no projection term)
    const bool cq = nq > 1e-12f, cpp = np_ > 1e-12f, cnn = nn_ > 1e-12f;
    float* dq = demb + static_cast<int64_t>(iq) * ldd;
    float* dp = demb + static_cast<int64_t>(ip) * ldd;
    float* dn = demb + static_cast<int64_t>(in_) * ldd;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
        const int col = (c * 32 + lane) * 4;
        if (col < d) {
            const float q4[4] = {vq[c].x, vq[c].y, vq[c].z, vq[c].w};
            const float p4[4] = {vp[c].x, vp[c].y, vp[c].z, vp[c].w};
            const float n4[4] = {vn[c].x, vn[c].y, vn[c].z, vn[c].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float qh = q4[e] * rq, ph = p4[e] * rp, nh = n4[e] * rn;
                // d(cn - cp)/dq^ = n^ - p^ ; project out q^ unless clamped
                const float gq = (nh - ph) - (cq ? qh * (cn - cp) : 0.f);
                const float gp = -(qh - (cpp ? ph * cp : 0.f));
                const float gn = (qh - (cnn ? nh * cn : 0.f));
                atomicAdd(dq + col + e, g * kq * gq * rq);
                atomicAdd(dp + col + e, g * kp * gp * rp);
                atomicAdd(dn + col + e, g * kn * gn * rn);
            }
        }
    }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float gscale) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float gi = g[i] * gscale;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        // torch.optim.Adam: denom = sqrt(v)/sqrt(bc2) + eps; p -= lr/bc1 * m/denom
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] -= (lr / bc1) * (mi / denom);
    }
}

// Training diagnostics of PinSage.train_batch (pinsage_training.py:200-212), which the reference computes every step
// with ~25 framework ops: (1) the cosine triplet loss of the RAW node features of the batch,
// mean_i max(d(a,p) - d(a,n) + margin, 0), d = 1 - cos, on F.normalize'd rows; (2) the batch variance of the query
// embeddings, sum_ij (h_ij - mean_j)^2 / (B - 1).  One warp per triple / one CTA per 32 columns.
__global__ void __launch_bounds__(kWarps * 32)
feat_triplet_kernel(const float* __restrict__ feats, int64_t ld, int d, const int64_t* __restrict__ batch, int64_t B,
                    float margin, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
    if (i >= B) return;
    const float* a = feats + batch[3 * i] * ld;
    const float* p = feats + batch[3 * i + 1] * ld;
    const float* n = feats + batch[3 * i + 2] * ld;
    float aa = 0.f, pp = 0.f, nn = 0.f, ap = 0.f, an = 0.f;
    for (int c = lane * 4; c < d; c += 128) {
        const float4 x = ps_ldg4(a + c), y = ps_ldg4(p + c), z = ps_ldg4(n + c);
        aa += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        pp += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
        nn += z.x * z.x + z.y * z.y + z.z * z.z + z.w * z.w;
        ap += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
        an += x.x * z.x + x.y * z.y + x.z * z.z + x.w * z.w;
    }
    aa = ps_warp_sum(aa); pp = ps_warp_sum(pp); nn = ps_warp_sum(nn); ap = ps_warp_sum(ap); an = ps_warp_sum(an);
    if (lane == 0) {
        // F.normalize (eps 1e-12), then cosine_similarity of the unit rows (eps 1e-8 on each norm)
        const float na = fmaxf(sqrtf(aa), 1e-12f), np_ = fmaxf(sqrtf(pp), 1e-12f), nn_ = fmaxf(sqrtf(nn), 1e-12f);
        const float ua = fmaxf(sqrtf(aa) / na, 1e-8f), up = fmaxf(sqrtf(pp) / np_, 1e-8f), un = fmaxf(sqrtf(nn) / nn_, 1e-8f);
        const float cap = ap / (na * np_) / (ua * up), can = an / (na * nn_) / (ua * un);
        atomicAdd(out, fmaxf((1.f - cap) - (1.f - can) + margin, 0.f) / static_cast<float>(B));
    }
}

__global__ void __launch_bounds__(256)
batch_variance_kernel(const float* __restrict__ emb, int64_t ld, int d, const int32_t* __restrict__ triples, int64_t B,
                      float* __restrict__ out) {
    __shared__ float part[8][33];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + lx;
    float s = 0.f;
    if (col < d)
        for (int64_t r = ly; r < B; r += 8) s += __ldg(emb + static_cast<int64_t>(__ldg(triples + 3 * r)) * ld + col);
    part[ly][lx] = s;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) mean += part[k][lx];
    mean /= static_cast<float>(B);
    __syncthreads();
    float q = 0.f;
    if (col < d)
        for (int64_t r = ly; r < B; r += 8) {
            const float x = __ldg(emb + static_cast<int64_t>(__ldg(triples + 3 * r)) * ld + col) - mean;
            q = fmaf(x, x, q);
        }
    part[ly][lx] = q;
    __syncthreads();
    if (ly == 0) {
#pragma unroll
        for (int k = 1; k < 8; ++k) q += part[k][lx];
        q = ps_warp_sum(q);
        if (lx == 0) atomicAdd(out, q / static_cast<float>(B - 1));
    }
}

}  // namespace

extern "C" int ps_train_diagnostics(const float* feats, int64_t ld_feats, int d_feat, const int64_t* batch, int64_t B,
                                    const float* emb, int64_t ld_emb, int d_emb, const int32_t* triples, float feat_margin,
                                    float* out2, ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(feats && batch && emb && triples && out2, "null pointer");
    PS_REQUIRE(B > 0 && d_feat > 0 && d_feat % 4 == 0 && ld_feats % 4 == 0 && d_emb > 0, "bad shape");
    PS_CUDA_CHECK(cudaMemsetAsync(out2, 0, 2 * sizeof(float), stream));
    feat_triplet_kernel<<<static_cast<unsigned>(ps_ceil_div(B, kWarps)), kWarps * 32, 0, stream>>>(feats, ld_feats, d_feat, batch, B, feat_margin, out2);
    PS_LAUNCH_CHECK();
    batch_variance_kernel<<<static_cast<unsigned>(ps_ceil_div(d_emb, 32)), 256, 0, stream>>>(emb, ld_emb, d_emb, triples, B, out2 + 1);
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_count_triples(const int32_t* triples, int64_t B, int64_t U, int32_t* dup_counts, ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(triples && dup_counts, "null pointer");
    PS_REQUIRE(B >= 0 && U > 0, "bad shape");
    PS_CUDA_CHECK(cudaMemsetAsync(dup_counts, 0, sizeof(int32_t) * 3 * U, stream));
    if (B == 0) return PS_OK;
    count_triples_kernel<<<static_cast<unsigned>(ps_ceil_div(B * 3, 256)), 256, 0, stream>>>(triples, B, U, dup_counts);
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_margin_loss_fwd_bwd(const float* emb, int64_t ld, const int32_t* triples, int64_t B, int d,
                                      float margin, float grad_scale, const int32_t* dup_counts, int64_t U,
                                      float* loss_out, float* demb, int64_t ldd, ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(emb && triples && loss_out, "null pointer");
    PS_REQUIRE(B > 0 && d > 0 && d % 4 == 0 && d <= 128 * kMaxChunks && ld % 4 == 0, "bad shape (d must be a multiple of 4, <= 1024)");
    margin_loss_kernel<<<static_cast<unsigned>(ps_ceil_div(B, kWarps)), kWarps * 32, 0, stream>>>(
        emb, ld, triples, B, d, margin, grad_scale, dup_counts, U, loss_out, demb, ldd);
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                            float lr, float beta1, float beta2, float eps, int64_t step, float grad_scale,
                            ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_REQUIRE(param && grad && exp_avg && exp_avg_sq, "null pointer");
    PS_REQUIRE(n >= 0 && step >= 1, "bad arguments");
    if (n == 0) return PS_OK;
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
    int64_t blocks = ps_ceil_div(n, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    adam_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                               static_cast<float>(bc1), static_cast<float>(sqrt(bc2)), grad_scale);
    PS_LAUNCH_CHECK();
    return PS_OK;
}
