/*
 * pinsage_b200.h -- C ABI of libpinsage_b200.so, the B200 (sm_100a) engine behind the
 * PinSage hot path of MatejBevec/gcn-song-embeddings.
 *
 * The reference is pure Python and has no FFI layer of its own, so
 * each entry point below cites the reference function (file:line under /root/reference)
 * whose arithmetic it replaces.  INTEGRATION.md shows the ctypes stubs a maintainer of
 * the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative PS_ERR_* code; the message is
 *     available from ps_last_error() (thread-local);
 *   - all pointers are DEVICE pointers owned by the caller unless stated otherwise; the
 *     library never allocates device memory behind the caller's back, except inside the
 *     opaque ps_graph_t: it keeps the caller's CSR pointers and owns a 4-byte copy of the
 *     row offsets (when the graph has fewer than 2^32 entries) that halves the bytes per hop;
 *     ps_gemm keeps one small scratch buffer per (device, stream) for packed weight images;
 *   - every launch is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - row-major matrices with an explicit leading dimension in ELEMENTS;
 *   - no global state; one host thread per device.
 */
#ifndef PINSAGE_B200_H
#define PINSAGE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PS_ABI_VERSION 1

typedef struct ps_graph ps_graph_t;
typedef void* ps_stream_t; /* cudaStream_t */

int ps_version(void);
const char* ps_last_error(void);
/* Select the CUDA device for the calling host thread (call once per process/thread). */
int ps_set_device(int device);

/* ---- graph: adjacency lookup that replaces DGL (pinsage_model.py:41,44,93;
 *      spotify_graph.py:48-63).  Nodes [0, n_tracks) are tracks, [n_tracks,
 *      n_tracks+n_cols) collections; CSR over all of them, multi-edges kept.
 *      Fails with PS_ERR_GRAPH if any node has no successors (the reference's
 *      torch.randint(0) raises there, pinsage_model.py:42). Synchronises `stream`. ---- */
int ps_graph_create(const int64_t* indptr, const int32_t* indices, int64_t n_tracks, int64_t n_cols,
                    int64_t n_entries, ps_graph_t** out, ps_stream_t stream);
int ps_graph_destroy(ps_graph_t* g);
/* The walker reads 4-byte row offsets (a compact copy owned by the handle) when the CSR has < 2^32 entries and the
 * caller's 8-byte indptr otherwise (BASELINE.json configs[3]: 2 x 10^9 entries).  on = 0 forces the 8-byte path on a
 * small graph (parity tests of that path), on = 1 restores the default.  Returns the previous setting (or < 0). */
int ps_graph_use_indptr32(ps_graph_t* g, int on);

/* ---- K1+K2: restart random walks fused with the visit-count top-T reduction.
 *      Replaces do_random_walks + sample_neighborhood + sample_neighborhood_topt
 *      (pinsage_model.py:32-53, 88-107).  One warp per source; draws keyed by
 *      Philox4x32-10(counter=(step, source, 0, 0), key=seed).  fixed_len > 0 = restart
 *      deterministically every fixed_len steps (BASELINE.json config 5) instead of the
 *      alpha draw.  Any output pointer may be NULL.  Outputs are ordered by (count desc,
 *      node id asc); slots beyond the number of distinct visited nodes carry weight 0 and
 *      the source id.  weights = count / n_hops (IEEE double divide). ---- */
int ps_walk_topt(const ps_graph_t* g, const int64_t* sources, int64_t n, int n_hops, double alpha,
                 int fixed_len, int T, uint64_t seed,
                 int64_t* out_nodes_i64, double* out_w_f64, int32_t* out_nodes_i32, float* out_w_f32,
                 int32_t* out_trace, ps_stream_t stream);
/* Two kernels implement ps_walk_topt / ps_trace_topt with identical results: the sort-based one (trace in shared memory,
 * sorted in registers; n_hops <= 512 and T <= 256) and the hash-table one (everything else).  mode = 1 forces the
 * hash-table kernel (parity tests of that path), 0 restores the default; returns the previous mode. */
int ps_walk_algo(int mode);
/* K2 alone on a caller-supplied trace [n, n_hops] (bit-exact parity hook against
 * sample_neighborhood_topt fed the same trace, pinsage_model.py:96-99,107). */
int ps_trace_topt(const int64_t* trace, const int64_t* sources, int64_t n, int n_hops, int T,
                  int64_t* out_nodes_i64, double* out_w_f64, int32_t* out_nodes_i32, float* out_w_f32,
                  ps_stream_t stream);

/* ---- K5/K7/K9 and their backward: the dense contraction
 *        C[i, j] (+)= act( sum_{r<K} P(i, r) * Q(j, r) + bias[j] )
 *      P(i, r) = p_kmajor ? P[prow(i)*ldp + r] : P[prow(r)*ldp + i]   (same for Q),
 *      prow(x) = p_rows ? p_rows[x] : x  (row gather folded into the operand load, K4).
 *      act: 0 none, 1 leaky_relu(0.01), 2 = C holds leaky_relu outputs y on entry and receives
 *      sum * leaky'(y) (backward through the activation fused into the store; no bias / l2norm).
 *      l2norm: divide each output row by its L2 norm
 *      (requires N <= 128; norm_out[i] receives the norm, may be NULL).
 *      accumulate: atomically add into C instead of storing (bias/act/l2norm must be off);
 *      splits > 1 partitions K over CTAs (needs accumulate).
 *      Replaces nn.Linear + leaky_relu + row normalise in ConvLayer.forward and the head
 *      (pinsage_model.py:201,208-210,259) and autograd's AddmmBackward. ---- */
int ps_gemm(const float* P, int64_t ldp, int p_kmajor, const int32_t* p_rows,
            const float* Q, int64_t ldq, int q_kmajor, const int32_t* q_rows,
            float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
            const float* bias, int act, int l2norm, float* norm_out, int accumulate, int splits,
            ps_stream_t stream);

/* ps_gemm with a sign mask of the activation (one bit per output element, row-major, ld_mask 32-bit words per row,
 * bit (j % 32) of word j / 32 = column j):  act = 1 additionally WRITES mask = (output > 0);  act = 2 READS it and
 * stores sum * leaky'(mask) without touching the old contents of C.  The forward Q GEMM records the mask and the
 * aggregation backward consumes it, which saves re-reading the dh-wide activations.  Tensor-core path only:
 * ps_gemm_mask_supported(M, N, K) tells whether a shape qualifies (N % 32 == 0 among others); mask = NULL is ps_gemm. */
int ps_gemm_mask_supported(int64_t M, int64_t N, int64_t K);
int ps_gemm_ex(const float* P, int64_t ldp, int p_kmajor, const int32_t* p_rows,
               const float* Q, int64_t ldq, int q_kmajor, const int32_t* q_rows,
               float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
               const float* bias, int act, int l2norm, float* norm_out, int accumulate, int splits,
               uint32_t* mask, int64_t ld_mask, ps_stream_t stream);

/* Select the ps_gemm implementation: 0 (default) = tcgen05 tensor cores with the 3xTF32
 * error-compensated split wherever the shape allows (CUDA cores otherwise), 1 = CUDA-core
 * fp32 only.  Returns the previous mode (any other argument just queries). */
int ps_gemm_backend(int mode);
/* Tuning knob of the tensor-core path: 1 (default) = weight operands (no gather, no split-K, M >= 1024) are
 * pre-split into hi/lo tile images once per call and streamed with bulk copies; 0 = every operand goes through
 * the producer warps.  Results are identical either way.  Returns the previous setting. */
int ps_gemm_tc_pack(int on);
/* Tuning knob: large tensor-core GEMMs launch waves x SMs persistent CTAs (default 1 = one CTA per SM; measured best on cfg3), so an SM is handed to a
 * pending higher-priority stream after 1/waves of the GEMM.  1 = one CTA per SM.  Returns the previous value. */
int ps_gemm_tc_waves(int waves);
/* Tuning knob: the persistent tensor-core GEMMs leave n SMs free (default 0) so that kernels of other streams (the
 * batch preparation of the next training step) never wait for a GEMM to retire.  Returns the previous value. */
int ps_gemm_tc_reserve_sms(int n);
/* 1 (default) = the tensor-core producers prefetch their next operand rows into L2 (a tile / 6 k-blocks ahead), 0 = off
 * (A/B measurements).  Returns the previous setting. */
int ps_gemm_tc_prefetch(int on);
/* How tall packed-weight GEMMs use thread-block clusters: 0 = one CTA per tile; 1 = 2-CTA clusters on tile pairs that share
 * the weight stream by TMA multicast (cta_group::1 MMAs); 2 = tile pairs on ONE 256-row tcgen05.mma.cta_group::2 per step,
 * each CTA staging half of the weight image and the output tile leaving through TMA tensor-map stores (default); 3 = 2, and the
 * weight-gradient GEMMs (activation x activation, split-K) run on pairs as well.  Returns the previous mode. */
int ps_gemm_tc_cluster(int mode);
/* Development: device array of 8 uint64 counters per CTA that the following tensor-core GEMM launches fill with the
 * cycles their warp roles spent waiting (NULL = off): [0] MMA-issue total, [1] its wait for operands, [2] its wait for a
 * free accumulator, [3] weight-stream wait for a free stage, [4] producer wait for a free stage, [5] accumulate-warp
 * wait for a finished chunk, [6] epilogue. */
int ps_gemm_tc_trace(unsigned long long* buf);
/* Development: ablation switches for the tensor-core kernels (results are WRONG while set): bit 0 = the activation producers skip
 * their global loads, bit 1 = the weight stream skips its bulk copies, bit 2 = the epilogue skips its stores.  Returns the
 * previous value; 0 restores normal operation. */
int ps_gemm_tc_experiment(int bits);

/* ---- K4+K6: neighbour gather + importance-weighted mean fused with the concat
 *      (pinsage_model.py:195-197,202,208):
 *        cat[i, 0:din]       = hin[self_rows[i], 0:din]
 *        cat[i, din:din+dh]  = sum_t nbw[i,t] * z[nbz[i,t], 0:dh] / sum_t nbw[i,t]
 *      inv_wsum[i] = 1 / sum_t nbw[i,t] is saved for the backward. ---- */
int ps_aggregate_fwd(const float* hin, int64_t ld_hin, const int32_t* self_rows, int din,
                     const float* z, int64_t ldz, const int32_t* nbz, const float* nbw, int T, int dh,
                     int64_t n, float* cat, int64_t ldcat, float* inv_wsum, ps_stream_t stream);
/* ---- K11: backward of the aggregation as a load-balanced segmented gather (no atomics,
 *      deterministic).  For z-row u with incoming (target, slot) pairs
 *      pair_q[seg_off[u] .. seg_off[u+1]) (q = i*T + t):
 *        z[u, :] = leaky'(z[u, :]) * sum_q nbw[q] * inv_wsum[q / T] * dcat[q / T, col_off : col_off+dh]
 *      (z holds leaky_relu outputs on entry and d(pre-activation) on exit; a row without
 *      pairs gets zeros).  Segments are cut into chunks of at most chunk_pairs pairs, one
 *      warp each: chunk_off[u] = sum_{v<u} ceil(len(v) / chunk_pairs), int32 [n_zrows+1];
 *      max_chunks >= chunk_off[n_zrows] sizes the launch (no host sync needed) and
 *      partial_ws holds max_chunks * dh floats of scratch for rows that span several chunks.
 *      chunk_row (optional, int32 [max_chunks]): the z-row that owns every chunk
 *      (last u with chunk_off[u] <= chunk); NULL = searched per chunk.
 *      apply_leaky = 0 turns the kernel into the plain segmented sum  z[u, :] = sum_q ...  (z is output only):
 *      the engine uses it on the do-wide d(pre-activation) rows and applies the W block + leaky' afterwards
 *      with ps_gemm(act = 2), which moves 4x fewer bytes than gathering the dh-wide dcat rows. ---- */
int ps_aggregate_bwd(const float* dcat, int64_t ldcat, int col_off, int dh,
                     const int32_t* seg_off, const int32_t* chunk_off, int chunk_pairs, int64_t max_chunks,
                     const int32_t* pair_q, const float* nbw, const float* inv_wsum, int T,
                     float* z, int64_t ldz, int64_t n_zrows, float* partial_ws, const int32_t* chunk_row,
                     int apply_leaky, ps_stream_t stream);
/* backward of  h = y / ||y||,  y = leaky_relu(pre)  (pinsage_model.py:209-210):
 *   dpre = leaky'(h) * (dh - h * (h . dh)) / norm */
int ps_norm_leaky_bwd(const float* h, int64_t ldh, const float* norm, const float* dh, int64_t lddh,
                      float* dpre, int64_t ldp, int64_t n, int d, ps_stream_t stream);
/* x[i,:] /= ||x[i,:]||, norm_out[i] = the norm (row normalise of pinsage_model.py:210 when out_dim > 128). */
int ps_l2norm_rows(float* x, int64_t ld, int64_t n, int d, float* norm_out, ps_stream_t stream);
/* dy[i] *= leaky'(y[i]) elementwise (y = leaky_relu output). */
int ps_leaky_bwd(const float* y, float* dy, int64_t n_elems, ps_stream_t stream);
/* out[j] += sum_i x[i*ld + j]  (bias gradients). */
int ps_colsum(const float* x, int64_t ld, int64_t n, int d, float* out, ps_stream_t stream);
/* dst[rows[i], 0:d] += src[i, 0:d]; rows must be unique (self-row gradient). */
int ps_scatter_add_rows(const float* src, int64_t lds, const int32_t* rows, float* dst, int64_t ldd,
                        int64_t n, int d, ps_stream_t stream);

/* ---- K10: max-margin loss forward + backward with the (q, pos, neg) gather and the
 *      gradient scatter-add fused (pinsage_training.py:31-41, 184-190).
 *      emb [U, d] holds one embedding per DISTINCT batch node; triples [B, 3] index rows
 *      of emb.  loss_out[0] += mean_b max(q^.n^ - q^.p^ + margin, 0)  (x^ = F.normalize,
 *      eps 1e-12).  demb [U, d] += d loss / d emb, where the contribution of a node that
 *      occurs k times in one column is additionally multiplied by k when dup_counts
 *      ([3, U] occurrence counts per column, from ps_count_triples) is non-NULL: that is
 *      the reference's duplicate-nodeset gradient factor (pinsage_model.py:260,265;
 *      SURVEY.md section 0 item 8).  grad_scale multiplies every gradient. ---- */
int ps_count_triples(const int32_t* triples, int64_t B, int64_t U, int32_t* dup_counts, ps_stream_t stream);
int ps_margin_loss_fwd_bwd(const float* emb, int64_t ld, const int32_t* triples, int64_t B, int d,
                           float margin, float grad_scale, const int32_t* dup_counts, int64_t U,
                           float* loss_out, float* demb, int64_t ldd, ps_stream_t stream);

/* ---- K3: frontier plans (relevant_nodes_per_layer[_precomp], pinsage_model.py:142-168) and the backward's
 *      (target, slot) -> z-row transpose, on the device.
 *      ps_plan_layer: nb int32 [n, T] = neighbour ids of the layer's targets cur int64 [n] (ids in [0, n_ids)).
 *        Builds the sorted distinct set of nb (united with cur when with_self) into uniq_i64 / uniq_i32 (either may
 *        be NULL; capacity min(n_ids, n*T + n)), nbz[q] = position of nb[q] in it, self_rows[i] = position of
 *        cur[i] (with_self only), and writes the set's size to count_out (device int32; the caller reads it).
 *      ps_plan_transpose: pairs q in [0, n_pairs) sorted by nbz[q] (stable) into pair_q; seg_off int32 [nz+1];
 *        chunk_off int32 [nz+1] and chunk_row int32 [max_chunks] as ps_aggregate_bwd expects them
 *        (max_chunks = n_pairs / chunk_pairs + nz).  Scratch is owned by the library, per (device, stream). ---- */
int ps_plan_layer(const int32_t* nb, int64_t n, int T, const int64_t* cur, int with_self, int64_t n_ids,
                  int64_t* uniq_i64, int32_t* uniq_i32, int32_t* nbz, int32_t* self_rows, int32_t* count_out,
                  ps_stream_t stream);
int ps_plan_transpose(const int32_t* nbz, int64_t n_pairs, int64_t nz, int chunk_pairs,
                      int32_t* pair_q, int32_t* seg_off, int32_t* chunk_off, int32_t* chunk_row, int64_t max_chunks,
                      ps_stream_t stream);

/* ---- K12: one training batch drawn on the device in one launch (pinsage_training.py:53-77,
 *      easy negatives): out_batch int64 [B, 3] = (q, pos) of B distinct uniformly random rows of
 *      positives [P, 2], and one negative per row: distinct uniformly random positions of all_ids
 *      (NULL = identity) that are not a node of those pairs.  Same law as the reference's
 *      randperm(P)[:B] / mask + randperm; draws are Philox4x32-10(counter = (candidate, phase,
 *      step), key = seed), reproducible on the CPU (oracle.sample_batch_philox).
 *      Limits: B <= 2600, 16*B <= P < 2^32-1, 16*B <= n_items < 2^31.  short_flag (may be NULL)
 *      is set to 1 if the candidate stream ran out (not reachable within the limits). ---- */
int64_t ps_sample_batch_workspace(int B); /* bytes of device scratch ps_sample_batch needs for batch size B */
int ps_sample_batch(const int64_t* positives, int64_t P, const int64_t* all_ids, int64_t n_items, int B,
                    uint64_t seed, uint64_t step, int64_t* out_batch, void* workspace, int64_t workspace_bytes,
                    int* short_flag, ps_stream_t stream);

/* ---- K14: per-row top-k of a row-major float matrix (the selection half of the cosine kNN search,
 *      baselines.py:91-103: cosine_sim.topk(k+1, dim=1); the similarity tile is a ps_gemm).
 *      out_val / out_idx [n_rows, k]: the k largest values of every row in descending order, ties by ascending
 *      column; NaN sorts as the largest value (as torch.topk).  1 <= k <= min(8192, n_cols), n_cols < 2^32. ---- */
int ps_topk_rows(const float* x, int64_t ld, int64_t n_rows, int64_t n_cols, int k,
                 float* out_val, int64_t* out_idx, ps_stream_t stream);
/* The fused form of the kNN search (baselines.py:91-103 without the [queries, N] similarity
 * tile): ps_gemm_filter runs the similarity GEMM sim[i, j] = sum_r P[i,r] Q[j,r] (P = the embedding table [M, K],
 * Q = a tile of query rows [N, K], both K-major, tcgen05 packed-weight path) and, instead of storing sim, appends
 * (sim, i) to query j's candidate list when sim >= thr[j]: slot = cnt[j]++ (cnt zeroed by the caller), written when
 * slot < cap to cand_val / cand_row [j * cap + slot].  The caller derives thr[j] from a sample of the table so that
 * a list ends up with a small multiple of k entries, checks cnt afterwards (cnt[j] < k or > cap: widen and redo),
 * and finishes with ps_topk_rows_mapped: the exact top-k of every list, ties by ascending row id like ps_topk_rows.
 * Returns PS_ERR_UNSUPPORTED for shapes outside the packed path (M < 1024, N < 64, K < 32, K % 4). */
int ps_gemm_filter(const float* P, int64_t ldp, const float* Q, int64_t ldq, int64_t M, int64_t N, int64_t K,
                   const float* thr, int32_t* cnt, float* cand_val, int32_t* cand_row, int cap, ps_stream_t stream);
int ps_topk_rows_mapped(const float* x, const int32_t* col_ids, const int32_t* row_counts, int64_t ld,
                        int64_t n_rows, int k, float* out_val, int64_t* out_idx, ps_stream_t stream);


/* ---- weight + bias gradient of a Linear layer in one call (AddmmBackward of nn.Linear, pinsage_model.py:201,208,259):
 *      C[i,j] += sum_r P[r*ldp + i] * Q[rows(r)*ldq + j]   (dW += dY^T X, optional row gather on X)
 *      p_colsum[i] += sum_r P[r*ldp + i]                   (db += colsum(dY); NULL = skip)
 *      splits = split-K factor (partials combined with fp32 atomics). ---- */
int ps_gemm_wgrad(const float* P, int64_t ldp, const float* Q, int64_t ldq, const int32_t* q_rows,
                  float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int splits, float* p_colsum,
                  ps_stream_t stream);

/* ---- per-step training diagnostics of PinSage.train_batch (pinsage_training.py:200-212): out2[0] = cosine triplet
 *      loss (margin feat_margin, mean) of the F.normalize'd raw feature rows of the batch's (q, pos, neg) ids
 *      (batch int64 [B,3]); out2[1] = batch variance of the query embeddings emb[triples[:,0]] (batch_variance,
 *      pinsage_training.py:99-103: sum of squared deviations from the column means / (B - 1)). ---- */
int ps_train_diagnostics(const float* feats, int64_t ld_feats, int d_feat, const int64_t* batch, int64_t B,
                         const float* emb, int64_t ld_emb, int d_emb, const int32_t* triples, float feat_margin,
                         float* out2, ps_stream_t stream);

/* ---- the fused training step as ONE host call (PinSage.train_batch's three forwards + max_margin_loss + backward,
 *      pinsage_training.py:184-190, on the shared frontier of the batch): forward of every layer of the plan, head,
 *      loss, backward into flat_grad (zeroed by the call), optional diagnostics.  The frontier plan (ps_plan_layer /
 *      ps_plan_transpose outputs per layer) and the parameters are passed by pointer; activations live in a caller-owned
 *      workspace of ps_train_step_workspace(args) bytes.  No host synchronisation; ~45 launches on `stream`. ---- */
#define PS_MAX_LAYERS 8
#define PS_AGG_BWD_CHUNK 64
typedef struct ps_layer_plan {
    int64_t n, nz;              /* targets of the layer; rows of its input that get the Q transform */
    const int32_t* self_rows;   /* [n]    row of each target in the layer input */
    const int32_t* nbz;         /* [n,T]  row of each neighbour among the transformed rows */
    const float* w;             /* [n,T]  importance weights */
    const int32_t* zrows;       /* [nz]   gather index into the feature table (layer 0) or NULL */
    const int32_t* seg_off;     /* [nz+1] backward transpose (ps_plan_transpose) */
    const int32_t* pair_q;      /* [n*T] */
    const int32_t* chunk_off;   /* [nz+1] */
    const int32_t* chunk_row;   /* [n*T/64 + nz] */
} ps_layer_plan;
typedef struct ps_layer_params {
    const float *Qw, *Qb, *Ww, *Wb; /* ConvLayer.Q / .W (pinsage_model.py:181-187): [dh,din], [dh], [do,din+dh], [do] */
    float *gQw, *gQb, *gWw, *gWb;   /* their gradients (views of flat_grad) */
} ps_layer_params;
typedef struct ps_step_args {
    int32_t n_layers, T, in_dim, hidden_dim, out_dim, reserved;
    const float* feats; int64_t ld_feats;
    ps_layer_plan layers[PS_MAX_LAYERS];
    ps_layer_params params[PS_MAX_LAYERS];
    const float *G1w, *G1b, *G2w; float *gG1w, *gG1b, *gG2w;
    const int32_t* triples; int64_t B;   /* [B,3] rows of the top layer's output per (q, pos, neg) */
    const int32_t* dup_counts;           /* [3, n_top] per-column occurrence counts (reference's duplicate factor) or NULL */
    float margin, feat_margin;
    float* flat_grad; int64_t n_params;
    void* workspace; int64_t workspace_bytes;
    float* loss_out;                     /* [1] */
    const int64_t* batch; float* diag_out; /* optional diagnostics: batch int64 [B,3] node ids, diag_out [2] (ps_train_diagnostics) */
    float** emb_out;                     /* optional HOST location that receives the device pointer of the [n_top, out_dim] embeddings */
    void* upper_grads_event;             /* optional cudaEvent_t recorded on `stream` once every gradient except layer 0's Q.weight / Q.bias is
                                            final (after layer 0's W weight gradient: the aggregation backward and the two largest GEMMs of the
                                            step still follow): a data-parallel caller starts the allreduce of that part of flat_grad behind it */
} ps_step_args;
int64_t ps_train_step_workspace(const ps_step_args* args);
int ps_train_step(const ps_step_args* args, ps_stream_t stream);
/* Per-call device timing of the launches inside ps_train_step (bench.py's roofline leg): enable, run steps, dump
 * "tag ms launches flops bytes" lines (synchronises the device).  ps_profile_dump returns the bytes written, or the
 * size needed when `out` is NULL / too small. */
int ps_profile_enable(int on);
int64_t ps_profile_dump(char* out, int64_t cap);

/* ---- the index-only preparation of a training batch in ONE host call (what PinSageModel.forward does before the
 *      first layer: relevant_nodes_per_layer_precomp, pinsage_model.py:156-168, 246-252, on the distinct nodes of the
 *      batch): top = sorted distinct ids of batch [B,3], triples / per-column counts, and per layer the neighbourhood
 *      lookup in the [n_ids, Tp] table, the next frontier, positions and the backward transpose (ps_plan_layer /
 *      ps_plan_transpose).  Outputs are laid out in the caller's device arena; `out` (HOST) receives sizes and byte
 *      offsets (-1 = absent).  Synchronises `stream` n_layers + 1 times.  PS_ERR_NOSPACE (-5): arena too small,
 *      out->bytes_needed says how much the part reached needs -- grow and call again.  PS_ERR_RANGE (-6): a batch id is
 *      outside [0, n_ids) (the reference raises IndexError). ---- */
typedef struct ps_plan_desc_layer {
    int64_t n, nz;
    int64_t off_nodes;      /* int64 [n]   node id of every target (sorted, distinct) */
    int64_t off_self_rows;  /* int32 [n] */
    int64_t off_nbz;        /* int32 [n,T] */
    int64_t off_w;          /* float [n,T] */
    int64_t off_zrows;      /* int32 [nz]  (layer 0 only) */
    int64_t off_seg_off, off_pair_q, off_chunk_off, off_chunk_row; /* backward transpose (need_backward) */
} ps_plan_desc_layer;
typedef struct ps_plan_desc {
    int64_t U;                                   /* distinct nodes of the batch */
    int64_t off_top, off_triples, off_counts;    /* int64 [U], int32 [B,3], int32 [3,U] */
    ps_plan_desc_layer layers[PS_MAX_LAYERS];
    int64_t bytes_used, bytes_needed;
} ps_plan_desc;
int ps_prepare_plan(const int64_t* batch, int64_t B, const int32_t* table_nodes, const float* table_w, int64_t n_ids, int Tp,
                    int T, int n_layers, int need_backward, void* arena, int64_t arena_bytes, ps_plan_desc* out, ps_stream_t stream);
/* The same preparation with ONLINE neighbourhoods (relevant_nodes_per_layer, pinsage_model.py:142-154): instead of a
 * table lookup, ps_walk_topt runs on every layer's targets (n_hops steps each, restart probability alpha, Philox key
 * `seed`; a node that is a target of several layers gets the same neighbourhood in each, since draws are keyed by
 * (seed, source, step)).  Node ids of the batch must lie in [0, n_items). */
int ps_prepare_plan_online(const int64_t* batch, int64_t B, const ps_graph_t* graph, int64_t n_items, int n_hops,
                           double alpha, uint64_t seed, int T, int n_layers, int need_backward, void* arena,
                           int64_t arena_bytes, ps_plan_desc* out, ps_stream_t stream);

/* ---- K13: Adam step on a flat fp32 parameter buffer (torch.optim.Adam defaults:
 *      betas, eps, no weight decay, no amsgrad; pinsage_training.py:147,191).
 *      step is the 1-based step count used for bias correction. ---- */
int ps_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                 float lr, float beta1, float beta2, float eps, int64_t step, float grad_scale,
                 ps_stream_t stream);

/* ---- ingest: the device side of SpotifyGraph.to_dgl_graph (spotify_graph.py:41-85).
 *      ps_csr_build: directed edge list as listed in graph.json (both directions present, duplicates kept,
 *      spotify_graph.py:48-63) -> CSR over n_nodes rows: indptr int64 [n_nodes + 1], indices int32 [n_edges].
 *      One stable radix sort by source keeps the listed order inside a row (DGL's successor order).  Endpoints
 *      outside [0, n_nodes) -> PS_ERR_RANGE (the reference's index_map raises KeyError on unknown ids).
 *      Synchronises `stream`; temporaries come from the stream-ordered allocator (ingest runs once).
 *      ps_standardize: x[:, j] = (x[:, j] - mean_j) / (std_j + eps) in place, std unbiased (N - 1)
 *      (spotify_graph.py:77-79; eps = 1e-12 there); column statistics accumulated in fp64;
 *      mean_out / std_out (float32 [d], may be NULL) receive mean_j and std_j + eps. ---- */
int ps_csr_build(const int64_t* src, const int64_t* dst, int64_t n_edges, int64_t n_nodes,
                 int64_t* indptr, int32_t* indices, ps_stream_t stream);
int ps_standardize(float* x, int64_t ld, int64_t n, int d, double eps, float* mean_out, float* std_out,
                   ps_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PINSAGE_B200_H */
