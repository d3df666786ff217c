// The fused training step as ONE host call: forward of the shared frontier, max-margin loss, backward into the
// flat gradient buffer, per-step diagnostics -- the ~45 kernel launches that ps_engine.Engine.train_step otherwise
// issues one ctypes call at a time (reference: PinSage.train_batch's three forwards + loss + backward,
// pinsage_training.py:184-190; PinSageModel.forward, pinsage_model.py:246-265; ConvLayer.forward, :189-212).
// Same kernels, same arguments, same order as the Python composition (ps_engine.py: Engine.forward / backward), so
// the results are identical; what changes is the host cost of a step (2.6 ms of interpreter time at the reference's
// default sizes, where the device work is 0.5 ms) and that the host thread holds no interpreter lock while launching.
// Host code only: every launch goes through the public ABI of this library.
#include "common.cuh"
#include "../../include/pinsage_b200.h"
#include <mutex>
#include <string>
#include <vector>

namespace {

// ---- optional per-call timing (bench.py's roofline leg): CUDA events on the launching stream around every tagged call
struct ProfRec { std::string tag; cudaEvent_t e0, e1; double flops, bytes; };
std::mutex g_prof_mutex;
bool g_prof_on = false;
std::vector<ProfRec> g_prof;

struct Timed {
    cudaStream_t s; bool on; ProfRec r;
    Timed(cudaStream_t s_, const char* tag, int layer, double flops, double bytes) : s(s_), on(g_prof_on) {
        if (!on) return;
        r.tag = tag;
        if (layer >= 0) r.tag += "_l" + std::to_string(layer);
        r.flops = flops; r.bytes = bytes;
        cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, s);
    }
    ~Timed() {
        if (!on) return;
        cudaEventRecord(r.e1, s);
        std::lock_guard<std::mutex> lock(g_prof_mutex);
        g_prof.push_back(r);
    }
};

inline int64_t align256(int64_t b) { return (b + 255) & ~static_cast<int64_t>(255); }

struct Arena {
    char* base; int64_t cap, off;
    template <typename T> T* take(int64_t count) {
        const int64_t bytes = align256(count * static_cast<int64_t>(sizeof(T)));
        char* p = base ? base + off : nullptr;
        off += bytes;
        return reinterpret_cast<T*>(p);
    }
};

int splits_for(int64_t M, int64_t N, int64_t K) {  // ps_engine._splits_for
    const int64_t tiles = ps_ceil_div(M, 128) * ps_ceil_div(N, 128);
    const int64_t want = ps_ceil_div(148 * 4, tiles);
    const int64_t kmax = ps_ceil_div(K, 256);
    const int64_t s = want < kmax ? want : kmax;
    return static_cast<int>(s < 1 ? 1 : s);
}

struct LayerBufs {
    float *z, *cat, *inv_wsum, *h, *norm, *d_pre, *s_buf, *agg_ws, *d_h_in, *d_self;
    uint32_t* zmask;
    int64_t max_chunks;
};

struct Bufs {
    LayerBufs L[PS_MAX_LAYERS];
    float *a1, *out, *d_out, *d_a1, *d_h_top;
};

// One carve of the workspace, used both to size it (base == nullptr) and to lay it out.
void carve(const ps_step_args* a, Arena& ar, Bufs& b) {
    const int dh = a->hidden_dim, dout = a->out_dim;
    for (int l = 0; l < a->n_layers; ++l) {
        const ps_layer_plan& lp = a->layers[l];
        const int din = l == 0 ? a->in_dim : dout;
        LayerBufs& lb = b.L[l];
        lb.z = ar.take<float>(lp.nz * dh);
        const bool mask = ps_gemm_mask_supported(lp.nz, dh, din) && ps_gemm_mask_supported(lp.nz, dh, dout);
        lb.zmask = mask ? ar.take<uint32_t>(lp.nz * (dh / 32)) : nullptr;
        lb.cat = ar.take<float>(lp.n * (din + dh));
        lb.inv_wsum = ar.take<float>(lp.n);
        lb.h = ar.take<float>(lp.n * dout);
        lb.norm = ar.take<float>(lp.n);
        lb.d_pre = ar.take<float>(lp.n * dout);
        lb.s_buf = ar.take<float>(lp.nz * dout);
        lb.max_chunks = lp.n * a->T / PS_AGG_BWD_CHUNK + lp.nz;
        lb.agg_ws = ar.take<float>((lb.max_chunks > 1 ? lb.max_chunks : 1) * dout);
        lb.d_h_in = l > 0 ? ar.take<float>(lp.nz * din) : nullptr;
        lb.d_self = l > 0 ? ar.take<float>(lp.n * din) : nullptr;
    }
    const int64_t n_top = a->layers[a->n_layers - 1].n;
    b.a1 = ar.take<float>(n_top * dout);
    b.out = ar.take<float>(n_top * dout);
    b.d_out = ar.take<float>(n_top * dout);
    b.d_a1 = ar.take<float>(n_top * dout);
    b.d_h_top = ar.take<float>(n_top * dout);
}

int check_args(const ps_step_args* a) {
    PS_REQUIRE(a != nullptr, "null pointer");
    PS_REQUIRE(a->n_layers >= 1 && a->n_layers <= PS_MAX_LAYERS, "n_layers must be in [1, %d]", PS_MAX_LAYERS);
    PS_REQUIRE(a->T > 0 && a->in_dim > 0 && a->hidden_dim > 0 && a->out_dim > 0, "bad dims");
    PS_REQUIRE(a->in_dim % 4 == 0 && a->hidden_dim % 4 == 0 && a->out_dim % 4 == 0, "dims must be multiples of 4");
    for (int l = 0; l < a->n_layers; ++l) PS_REQUIRE(a->layers[l].n > 0 && a->layers[l].nz > 0, "empty layer plan");
    return PS_OK;
}

#define PS_TRY(expr)                 \
    do {                             \
        const int _rc = (expr);      \
        if (_rc != PS_OK) return _rc; \
    } while (0)

double gemm_bytes(int64_t M, int64_t N, int64_t K) { return 4.0 * (static_cast<double>(M) * K + static_cast<double>(N) * K + static_cast<double>(M) * N); }

}  // namespace

extern "C" int ps_profile_enable(int on) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_prof.clear();
    g_prof_on = on != 0;
    return PS_OK;
}

// "tag ms launches flops bytes\n" per tag, after a device synchronise; returns the number of bytes written (or < 0)
extern "C" int64_t ps_profile_dump(char* out, int64_t cap) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    if (cudaDeviceSynchronize() != cudaSuccess) return PS_ERR_CUDA;
    struct Acc { std::string tag; double ms = 0, flops = 0, bytes = 0; int64_t n = 0; };
    std::vector<Acc> accs;
    for (auto& r : g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) { (void)cudaGetLastError(); continue; }
        Acc* t = nullptr;
        for (auto& x : accs) if (x.tag == r.tag) { t = &x; break; }
        if (!t) { accs.push_back(Acc()); t = &accs.back(); t->tag = r.tag; }
        t->ms += ms; t->flops += r.flops; t->bytes += r.bytes; t->n += 1;
    }
    std::string s;
    char line[256];
    for (auto& x : accs) {
        snprintf(line, sizeof(line), "%s %.6f %lld %.6e %.6e\n", x.tag.c_str(), x.ms, static_cast<long long>(x.n), x.flops, x.bytes);
        s += line;
    }
    if (out == nullptr || cap <= static_cast<int64_t>(s.size())) return static_cast<int64_t>(s.size()) + 1;
    memcpy(out, s.c_str(), s.size() + 1);
    return static_cast<int64_t>(s.size());
}

extern "C" int64_t ps_train_step_workspace(const ps_step_args* a) {
    if (check_args(a) != PS_OK) return PS_ERR_INVALID;
    Arena ar{nullptr, 0, 0};
    Bufs b;
    carve(a, ar, b);
    return ar.off;
}

extern "C" int ps_train_step(const ps_step_args* a, ps_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PS_TRY(check_args(a));
    PS_REQUIRE(a->feats && a->triples && a->loss_out && a->flat_grad && a->workspace && a->G1w && a->G1b && a->G2w, "null pointer");
    Arena ar{static_cast<char*>(a->workspace), a->workspace_bytes, 0};
    Bufs b;
    carve(a, ar, b);
    PS_REQUIRE(ar.off <= a->workspace_bytes, "workspace too small: %lld bytes needed (ps_train_step_workspace)", static_cast<long long>(ar.off));
    const int L = a->n_layers, T = a->T, dh = a->hidden_dim, dout = a->out_dim;
    const int64_t n_top = a->layers[L - 1].n;

    // ---------------- forward (Engine.forward)
    const float* h_prev = a->feats;
    int64_t ld_prev = a->ld_feats;
    for (int l = 0; l < L; ++l) {
        const ps_layer_plan& lp = a->layers[l];
        const ps_layer_params& pr = a->params[l];
        const LayerBufs& lb = b.L[l];
        const int din = l == 0 ? a->in_dim : dout;
        {
            Timed t(stream, "gemm_q_fwd", l, 2.0 * lp.nz * dh * din, gemm_bytes(lp.nz, dh, din));
            PS_TRY(ps_gemm_ex(h_prev, ld_prev, 1, lp.zrows, pr.Qw, din, 1, nullptr, lb.z, dh, lp.nz, dh, din, pr.Qb, 1, 0, nullptr, 0, 1,
                              lb.zmask, dh / 32, stream));
        }
        {
            Timed t(stream, "aggregate_fwd", l, 2.0 * lp.n * T * dh,
                    static_cast<double>(lp.n) * (static_cast<double>(T) * dh * 4 + din * 4 + T * 8 + 4 + (din + dh) * 4 + 4));
            PS_TRY(ps_aggregate_fwd(h_prev, ld_prev, lp.self_rows, din, lb.z, dh, lp.nbz, lp.w, T, dh, lp.n, lb.cat, din + dh, lb.inv_wsum, stream));
        }
        {
            Timed t(stream, "gemm_w_fwd", l, 2.0 * lp.n * dout * (din + dh), gemm_bytes(lp.n, dout, din + dh));
            if (dout <= 128) {
                PS_TRY(ps_gemm_ex(lb.cat, din + dh, 1, nullptr, pr.Ww, din + dh, 1, nullptr, lb.h, dout, lp.n, dout, din + dh, pr.Wb, 1, 1, lb.norm,
                                  0, 1, nullptr, 0, stream));
            } else {
                PS_TRY(ps_gemm_ex(lb.cat, din + dh, 1, nullptr, pr.Ww, din + dh, 1, nullptr, lb.h, dout, lp.n, dout, din + dh, pr.Wb, 1, 0, nullptr,
                                  0, 1, nullptr, 0, stream));
                PS_TRY(ps_l2norm_rows(lb.h, dout, lp.n, dout, lb.norm, stream));
            }
        }
        h_prev = lb.h;
        ld_prev = dout;
    }
    {
        Timed t(stream, "gemm_head", -1, 4.0 * n_top * dout * dout, 2 * gemm_bytes(n_top, dout, dout));
        PS_TRY(ps_gemm_ex(h_prev, dout, 1, nullptr, a->G1w, dout, 1, nullptr, b.a1, dout, n_top, dout, dout, a->G1b, 1, 0, nullptr, 0, 1, nullptr, 0, stream));
        PS_TRY(ps_gemm_ex(b.a1, dout, 1, nullptr, a->G2w, dout, 1, nullptr, b.out, dout, n_top, dout, dout, nullptr, 0, 0, nullptr, 0, 1, nullptr, 0, stream));
    }

    // ---------------- loss + its gradient (Engine.train_step)
    PS_CUDA_CHECK(cudaMemsetAsync(a->loss_out, 0, sizeof(float), stream));
    PS_CUDA_CHECK(cudaMemsetAsync(b.d_out, 0, sizeof(float) * n_top * dout, stream));
    PS_TRY(ps_margin_loss_fwd_bwd(b.out, dout, a->triples, a->B, dout, a->margin, 1.0f, a->dup_counts, n_top, a->loss_out, b.d_out, dout, stream));
    if (a->diag_out != nullptr && a->batch != nullptr)
        PS_TRY(ps_train_diagnostics(a->feats, a->ld_feats, a->in_dim, a->batch, a->B, b.out, dout, dout, a->triples, a->feat_margin, a->diag_out, stream));
    PS_CUDA_CHECK(cudaMemsetAsync(a->flat_grad, 0, sizeof(float) * a->n_params, stream));

    // ---------------- backward (Engine.backward)
    {
        Timed t(stream, "gemm_head_bwd", -1, 8.0 * n_top * dout * dout, 4 * gemm_bytes(n_top, dout, dout));
        PS_TRY(ps_gemm_wgrad(b.d_out, dout, b.a1, dout, nullptr, a->gG2w, dout, dout, dout, n_top, splits_for(dout, dout, n_top), nullptr, stream));
        PS_TRY(ps_gemm_ex(b.d_out, dout, 1, nullptr, a->G2w, dout, 0, nullptr, b.d_a1, dout, n_top, dout, dout, nullptr, 0, 0, nullptr, 0, 1, nullptr, 0, stream));
        PS_TRY(ps_leaky_bwd(b.a1, b.d_a1, n_top * dout, stream));
        PS_TRY(ps_gemm_wgrad(b.d_a1, dout, b.L[L - 1].h, dout, nullptr, a->gG1w, dout, dout, dout, n_top, splits_for(dout, dout, n_top), a->gG1b, stream));
        PS_TRY(ps_gemm_ex(b.d_a1, dout, 1, nullptr, a->G1w, dout, 0, nullptr, b.d_h_top, dout, n_top, dout, dout, nullptr, 0, 0, nullptr, 0, 1, nullptr, 0, stream));
    }
    const float* d_h = b.d_h_top;
    for (int l = L - 1; l >= 0; --l) {
        const ps_layer_plan& lp = a->layers[l];
        const ps_layer_params& pr = a->params[l];
        const LayerBufs& lb = b.L[l];
        const int din = l == 0 ? a->in_dim : dout;
        const float* h_in = l == 0 ? a->feats : b.L[l - 1].h;
        const int64_t ld_in = l == 0 ? a->ld_feats : dout;
        PS_TRY(ps_norm_leaky_bwd(lb.h, dout, lb.norm, d_h, dout, lb.d_pre, dout, lp.n, dout, stream));
        {
            Timed t(stream, "gemm_w_wgrad", l, 2.0 * dout * (din + dh) * lp.n, gemm_bytes(dout, din + dh, lp.n));
            PS_TRY(ps_gemm_wgrad(lb.d_pre, dout, lb.cat, din + dh, nullptr, pr.gWw, din + dh, dout, din + dh, lp.n, splits_for(dout, din + dh, lp.n), pr.gWb, stream));
        }
        if (l == 0 && L >= 2 && a->upper_grads_event != nullptr)  // every gradient but layer 0's Q.weight / Q.bias is final from here on
            PS_CUDA_CHECK(cudaEventRecord(static_cast<cudaEvent_t>(a->upper_grads_event), stream));
        {
            const double pairs = static_cast<double>(lp.n) * T;
            Timed t(stream, "aggregate_bwd", l, 2.0 * pairs * dout, pairs * (dout * 4 + 12) + static_cast<double>(lp.nz) * (dout * 4 + 8));
            PS_TRY(ps_aggregate_bwd(lb.d_pre, dout, 0, dout, lp.seg_off, lp.chunk_off, PS_AGG_BWD_CHUNK, lb.max_chunks, lp.pair_q, lp.w, lb.inv_wsum, T,
                                    lb.s_buf, dout, lp.nz, lb.agg_ws, lp.chunk_row, 0, stream));
        }
        {
            Timed t(stream, "gemm_agg_dgrad", l, 2.0 * lp.nz * dh * dout, gemm_bytes(lp.nz, dh, dout));
            PS_TRY(ps_gemm_ex(lb.s_buf, dout, 1, nullptr, pr.Ww + din, din + dh, 0, nullptr, lb.z, dh, lp.nz, dh, dout, nullptr, 2, 0, nullptr, 0, 1,
                              lb.zmask, dh / 32, stream));
        }
        {
            Timed t(stream, "gemm_q_wgrad", l, 2.0 * dh * din * lp.nz, gemm_bytes(dh, din, lp.nz));
            PS_TRY(ps_gemm_wgrad(lb.z, dh, h_in, ld_in, lp.zrows, pr.gQw, din, dh, din, lp.nz, splits_for(dh, din, lp.nz), pr.gQb, stream));
        }
        if (l > 0) {
            {
                Timed t(stream, "gemm_q_dgrad", l, 2.0 * lp.nz * din * dh, gemm_bytes(lp.nz, din, dh));
                PS_TRY(ps_gemm_ex(lb.z, dh, 1, nullptr, pr.Qw, din, 0, nullptr, lb.d_h_in, din, lp.nz, din, dh, nullptr, 0, 0, nullptr, 0, 1, nullptr, 0, stream));
            }
            {
                Timed t(stream, "gemm_w_dgrad", l, 2.0 * lp.n * din * dout, gemm_bytes(lp.n, din, dout));
                PS_TRY(ps_gemm_ex(lb.d_pre, dout, 1, nullptr, pr.Ww, din + dh, 0, nullptr, lb.d_self, din, lp.n, din, dout, nullptr, 0, 0, nullptr, 0, 1, nullptr, 0, stream));
            }
            PS_TRY(ps_scatter_add_rows(lb.d_self, din, lp.self_rows, lb.d_h_in, din, lp.n, din, stream));
            d_h = lb.d_h_in;
        }
    }
    if (a->emb_out != nullptr) *a->emb_out = b.out;
    return PS_OK;
}
