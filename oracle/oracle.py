"""CPU oracle for the PinSage hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the algorithm of the reference
(MatejBevec/gcn-song-embeddings) for the path SURVEY.md section 8 names.  It is the
checker for the CUDA product path and the `cpu_baseline` / `--impl reference` arm of
bench.py.  Nothing under `gcn-song-embeddings_b200/` may import it: only `tests/`,
`__graft_entry__.smoke()` and `bench.py` do.

Parity pinning: the reference ships no tests and no golden vectors,
so this oracle is pinned against outputs of the reference ITSELF, generated in the dev
container by `oracle/make_golden.py` (which imports the unmodified reference through
`oracle/refshim`) and committed under `tests/golden/`.  `tests/test_oracle_golden.py`
checks every function below against those fixtures.

Two kinds of code live here:
  * integer / index work in numpy (walk traces, visit-count top-T, frontier
    construction, hit-rate / MRR) -- compared bit-exactly;
  * the floating-point model (ConvLayer, PinSageModel, max-margin loss, train step) in
    CPU torch fp32/fp64 with autograd -- compared within 1e-4 relative.

All `file:line` citations are into /root/reference.
"""
from __future__ import annotations

import numpy as np
import torch

LEAKY_SLOPE = 0.01  # torch.nn.functional.leaky_relu default, pinsage_model.py:201,209,259

# ----------------------------------------------------------------------------------------
# Philox4x32-10 counter-based RNG (Salmon et al., "Parallel random numbers: as easy as
# 1, 2, 3", SC'11).  The reference draws from torch's global mt19937 stream
# (pinsage_model.py:42,45,50), which cannot be parallelised; the product walker keys
# every draw by (seed, source node, step) instead.  The oracle restates the SAME keyed
# draws so that GPU and CPU traces are bit-identical.
# ----------------------------------------------------------------------------------------
_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds.  All arguments broadcastable uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.asarray(k0, dtype=np.uint32)
    k1 = np.asarray(k1, dtype=np.uint32)
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            if r != 9:
                k0 = (k0 + _W0).astype(np.uint32)
                k1 = (k1 + _W1).astype(np.uint32)
    return c0, c1, c2, c3


def restart_threshold(alpha: float) -> int:
    """Integer threshold t such that `x < t` (x uniform uint32) has probability ~alpha.
    The reference tests `torch.rand(()) < alpha` (pinsage_model.py:50)."""
    return int(float(alpha) * 4294967296.0)


def _mulhi32(x, n):
    """floor(x * n / 2^32): uniform index in [0, n) from a uniform uint32 x."""
    return ((x.astype(np.uint64) * n.astype(np.uint64)) >> np.uint64(32)).astype(np.int64)


def do_random_walks_philox(indptr, indices, nodeset, n_hops, alpha, seed, fixed_len=0):
    """Restart random walk traces; restates `do_random_walks` (pinsage_model.py:32-53).

    Per source i and step j (exactly n_hops steps): item -> uniform successor
    (a collection) -> uniform successor (an item); record it in trace[i, j]; then restart
    to the source with probability alpha (alpha IS the restart probability, :50-51).
    The three draws of step j come from Philox(counter=(j, source, 0, 0),
    key=(seed_lo, seed_hi)): x0 picks the collection, x1 the item, x2 the restart.
    `fixed_len` > 0 selects the deterministic-restart variant of BASELINE.json config 5:
    restart after every `fixed_len` steps instead of the alpha draw.
    Vectorised over sources (steps stay sequential, as in the reference).
    """
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices)
    src = np.asarray(nodeset, dtype=np.int64)
    n = src.shape[0]
    trace = np.zeros((n, n_hops), dtype=np.int64)
    k0 = np.uint32(seed & 0xFFFFFFFF)
    k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
    thr = np.uint64(restart_threshold(alpha))
    item = src.copy()
    src32 = src.astype(np.uint32)
    zero = np.zeros(n, dtype=np.uint32)
    for j in range(n_hops):
        x0, x1, x2, _ = philox4x32_10(np.full(n, j, dtype=np.uint32), src32, zero, zero, k0, k1)
        beg = indptr[item]
        deg = indptr[item + 1] - beg
        if np.any(deg <= 0):
            raise RuntimeError("random walk reached a node with no successors")
        col = indices[beg + _mulhi32(x0, deg)].astype(np.int64)
        beg = indptr[col]
        deg = indptr[col + 1] - beg
        if np.any(deg <= 0):
            raise RuntimeError("random walk reached a node with no successors")
        item = indices[beg + _mulhi32(x1, deg)].astype(np.int64)
        trace[:, j] = item
        if fixed_len > 0:
            restart = np.full(n, (j + 1) % fixed_len == 0)
        else:
            restart = x2.astype(np.uint64) < thr
        item = np.where(restart, src, item)
    return trace


def topt_from_trace(trace, nodeset, T):
    """Visit counts -> probabilities -> zero the self entry -> top-T.

    Restates `sample_neighborhood` + `sample_neighborhood_topt`
    (pinsage_model.py:88-107): prob = count / n_hops in IEEE float64, the source's own
    entry is zeroed AFTER normalising (:98-99), then the T largest are kept (:107).
    `torch.topk` breaks ties in an implementation-defined order and, when fewer than T
    distinct nodes were visited, returns zero-weight fillers with arbitrary indices
    (SURVEY.md section 0 item 9).  The engine's canonical order, restated here, is
    (count descending, node id ascending); zero-weight slots carry weight 0.0 and the
    source's own id (they add no frontier node and cannot index out of bounds).
    Returns (weights float64 [n, T], nodes int64 [n, T]).
    """
    trace = np.asarray(trace, dtype=np.int64)
    nodeset = np.asarray(nodeset, dtype=np.int64)
    n, n_hops = trace.shape
    weights = np.zeros((n, T), dtype=np.float64)
    nodes = np.repeat(nodeset[:, None], T, axis=1)
    for i in range(n):
        ids, counts = np.unique(trace[i], return_counts=True)
        keep = ids != nodeset[i]
        ids, counts = ids[keep], counts[keep]
        order = np.lexsort((ids, -counts))[:T]
        k = order.shape[0]
        weights[i, :k] = counts[order].astype(np.float64) / np.float64(n_hops)
        nodes[i, :k] = ids[order]
    return weights, nodes


def check_topt_against_reference(ref_w, ref_nodes, our_w, our_nodes, trace, nodeset):
    """The parity rule of SURVEY.md section 8a row A3, as an assertion helper.

    * the weight vectors are bit-equal;
    * for every weight strictly greater than the T-th (boundary) weight the index SETS
      are equal;
    * indices at the boundary weight are drawn from the nodes that have exactly that
      count in the trace;
    * zero-weight slots are ignored on the reference side (arbitrary fillers) and must be
      (0.0, source id) on ours.
    Returns None or raises AssertionError.
    """
    ref_w = np.asarray(ref_w); our_w = np.asarray(our_w)
    ref_nodes = np.asarray(ref_nodes); our_nodes = np.asarray(our_nodes)
    trace = np.asarray(trace); nodeset = np.asarray(nodeset)
    n, T = ref_w.shape
    n_hops = trace.shape[1]
    assert our_w.shape == ref_w.shape and our_nodes.shape == ref_nodes.shape
    assert np.array_equal(ref_w.view(np.int64), our_w.view(np.int64)), "weights not bit-equal"
    for i in range(n):
        ids, counts = np.unique(trace[i], return_counts=True)
        cnt = dict(zip(ids.tolist(), counts.tolist()))
        cnt.pop(int(nodeset[i]), None)
        boundary = ref_w[i, T - 1]
        above = ref_w[i] > boundary
        assert set(ref_nodes[i, above].tolist()) == set(our_nodes[i, above].tolist()), f"row {i}: set above boundary differs"
        for t in range(T):
            w = our_w[i, t]
            if w == 0.0:
                assert our_nodes[i, t] == nodeset[i], f"row {i}: zero slot must carry the source id"
            else:
                c = cnt.get(int(our_nodes[i, t]))
                assert c is not None and np.float64(c) / np.float64(n_hops) == w, f"row {i} slot {t}: weight/count mismatch"
        nz = our_w[i] > 0
        assert len(set(our_nodes[i, nz].tolist())) == int(nz.sum()), f"row {i}: duplicate neighbours"


def relevant_nodes_per_layer_precomp(nodeset, n_layers, T, nbhds):
    """T-hop computation graph from precomputed neighbourhoods; restates
    `relevant_nodes_per_layer_precomp` (pinsage_model.py:156-168).  Top-layer nodeset
    keeps duplicates and order; lower layers are sorted-unique (:166)."""
    all_w, all_nb = nbhds
    all_w = np.asarray(all_w); all_nb = np.asarray(all_nb)
    cur = np.asarray(nodeset, dtype=np.int64)
    S = []
    for _ in range(n_layers):
        w, nb = all_w[cur, :T], all_nb[cur, :T]
        S.insert(0, (cur, w, nb))
        cur = np.unique(np.concatenate([nb.reshape(-1), cur]))
    return S


# ----------------------------------------------------------------------------------------
# Floating-point model (CPU torch).  Parameters are a flat dict with the reference's
# state-dict keys: conv_layers.{i}.Q.weight/bias, conv_layers.{i}.W.weight/bias,
# G1.weight/bias, G2.weight  (pinsage_model.py:181-187,234-244).
# ----------------------------------------------------------------------------------------

def make_params(n_layers, dims, rng: np.random.RandomState, dtype=torch.float32):
    """Deterministic parameter set with the reference's shapes and init law
    (xavier-uniform weights, biases 0.3; pinsage_model.py:181-187,239-244), drawn from a
    numpy RandomState so fixtures are reproducible on any platform."""
    din, dh, do = dims
    in_dims = [din] + [do] * (n_layers - 1)

    def xavier(o, i):
        a = np.sqrt(6.0 / (i + o))
        return torch.tensor(rng.uniform(-a, a, size=(o, i)), dtype=dtype)

    p = {}
    for l in range(n_layers):
        p[f"conv_layers.{l}.Q.weight"] = xavier(dh, in_dims[l])
        p[f"conv_layers.{l}.Q.bias"] = torch.full((dh,), 0.3, dtype=dtype)
        p[f"conv_layers.{l}.W.weight"] = xavier(do, in_dims[l] + dh)
        p[f"conv_layers.{l}.W.bias"] = torch.full((do,), 0.3, dtype=dtype)
    p["G1.weight"] = xavier(do, do)
    p["G1.bias"] = torch.full((do,), 0.3, dtype=dtype)
    p["G2.weight"] = xavier(do, do)
    return p


def _leaky_with_signs(pre, row_nodes, forced):
    """leaky_relu(pre) where, for the few entries listed in `forced` = [(node id, column, positive?)], the branch
    of the activation is GIVEN instead of decided by the sign of `pre`.  leaky_relu's derivative jumps by 100x at
    0, so an entry whose pre-activation lies within fp32 rounding of 0 gets its branch from the summation order of
    whoever computes it (MKL, a CUDA-core loop, the tensor core) -- no two fp32 implementations agree on those.
    The parity tests pass the product's decisions for exactly the entries a float64 evaluation proves to be
    rounding-ambiguous, so that both sides differentiate the same function.  Values change by < 1e-6 absolute.
    `pre` is [rows, D]; `row_nodes` [rows] the node id each row belongs to."""
    act = torch.nn.functional.leaky_relu(pre, LEAKY_SLOPE)
    if not forced:
        return act
    rr, cc, slope = [], [], []
    for node, col, positive in forced:
        hit = (row_nodes == int(node)).nonzero().flatten()
        rr.append(hit); cc.append(torch.full_like(hit, int(col)))
        slope.append(torch.full((hit.numel(),), 1.0 if positive else LEAKY_SLOPE, dtype=pre.dtype))
    rr, cc, slope = torch.cat(rr), torch.cat(cc), torch.cat(slope)
    if rr.numel() == 0:
        return act
    return act.index_put((rr, cc), pre[rr, cc] * slope)


def conv_layer_forward(h, nodeset, nb_nodes, nb_weights, Qw, Qb, Ww, Wb, forced=None):
    """One PinSage convolution; restates `ConvLayer.forward` (pinsage_model.py:189-212).
    leaky_relu(Q .) on the T gathered neighbour rows, importance-weighted mean in float64
    (the weights are f64, :202), concat with the self row, `.float()`, leaky_relu(W .),
    row L2-normalise without eps (:210).  `forced` = {"Q": [...], "W": [...]} (optional, tests only): activation
    branches given for rounding-ambiguous entries, see _leaky_with_signs."""
    din = Qw.shape[1]
    n, T = nb_nodes.shape
    forced = forced or {}
    self_h = h[nodeset, :din]
    nb_flat = nb_nodes.reshape(-1)
    pre = torch.nn.functional.linear(h[nb_flat, :din], Qw, Qb)
    nb_h = _leaky_with_signs(pre, nb_flat, forced.get("Q")).reshape(n, T, -1)
    agg = (nb_weights[:, :, None] * nb_h).sum(1) / nb_weights.sum(1, keepdim=True)
    cat = torch.cat([self_h, agg], 1).float()
    new_h = _leaky_with_signs(torch.nn.functional.linear(cat, Ww, Wb), nodeset, forced.get("W"))
    return new_h / new_h.norm(dim=1, keepdim=True)


def _put(h, nodeset, x):
    """`put_embeddings` (pinsage_model.py:24-30): detached clone of the WHOLE table, rows
    of `nodeset` overwritten with x zero-padded to the table width."""
    new_h = h.clone().detach()
    pad = torch.zeros(x.shape[0], new_h.shape[1] - x.shape[1], dtype=x.dtype)
    new_h[nodeset, :] = torch.cat([x, pad], 1)
    return new_h


def model_forward(params, features, nodeset, nbhds, T, n_layers, forced=None):
    """`PinSageModel.forward` (pinsage_model.py:246-265) including the full-table
    put/get round trips, so that autograd reproduces the duplicate-node gradient factor
    of the final put/get pair (:260,265)."""
    all_w, all_nb = nbhds
    nodeset = torch.as_tensor(nodeset, dtype=torch.int64)
    S = relevant_nodes_per_layer_precomp(nodeset.numpy(), n_layers, T, (all_w.numpy(), all_nb.numpy()))
    h = features
    new = None
    for l, (ns, w, nb) in enumerate(S):
        ns_t = torch.from_numpy(ns)
        new = conv_layer_forward(h, ns_t, torch.from_numpy(nb), torch.from_numpy(w),
                                 params[f"conv_layers.{l}.Q.weight"], params[f"conv_layers.{l}.Q.bias"],
                                 params[f"conv_layers.{l}.W.weight"], params[f"conv_layers.{l}.W.bias"],
                                 forced=(forced or {}).get(l))
        h = _put(h, ns_t, new)
    new = torch.nn.functional.linear(
        torch.nn.functional.leaky_relu(torch.nn.functional.linear(new, params["G1.weight"], params["G1.bias"]), LEAKY_SLOPE),
        params["G2.weight"])
    h = _put(h, nodeset, new)
    return h[nodeset, :params["G2.weight"].shape[0]]


def max_margin_loss(h_q, h_pos, h_neg, margin):
    """`max_margin_loss` (pinsage_training.py:31-41): F.normalize (eps 1e-12) each input,
    per-row dots, mean(max(q.n - q.p + margin, 0)); max taken over a stacked pair so a
    tie sends the subgradient to the first argument."""
    norm = torch.nn.functional.normalize
    h_q, h_pos, h_neg = norm(h_q, dim=1), norm(h_pos, dim=1), norm(h_neg, dim=1)
    d = (h_q * h_neg).sum(1) - (h_q * h_pos).sum(1) + margin
    return torch.stack([d, torch.zeros_like(d)], 1).max(1).values.mean()


def train_batch_grads(params, features, batch, nbhds, T, n_layers, margin, forced=None):
    """Loss and parameter gradients of one (q, pos, neg) batch; restates the first half
    of `PinSage.train_batch` (pinsage_training.py:184-190): three separate forwards, the
    max-margin loss, backward.  `forced` = {layer: {"Q": [(node, col, positive?)], "W": [...]}}: see
    _leaky_with_signs (tests at sizes where rounding-ambiguous activations exist)."""
    p = {k: v.clone().detach().requires_grad_(True) for k, v in params.items()}
    batch = torch.as_tensor(batch, dtype=torch.int64)
    hq = model_forward(p, features, batch[:, 0], nbhds, T, n_layers, forced)
    hp = model_forward(p, features, batch[:, 1], nbhds, T, n_layers, forced)
    hn = model_forward(p, features, batch[:, 2], nbhds, T, n_layers, forced)
    loss = max_margin_loss(hq, hp, hn, margin)
    loss.backward()
    return loss.detach(), {k: v.grad.detach() for k, v in p.items()}, (hq.detach(), hp.detach(), hn.detach())


class OracleTrainer:
    """CPU restatement of the `PinSage` trainer's step (pinsage_training.py:142-148,
    181-191): Adam(lr) on the parameters, ExponentialLR(decay) per epoch.  Used as the
    CPU baseline ("port") by bench.py and as the step-level parity checker."""

    def __init__(self, params, features, nbhds, T, n_layers, margin=1e-5, lr=1e-4):
        self.params = {k: v.clone().detach().requires_grad_(True) for k, v in params.items()}
        self.features, self.nbhds, self.T, self.n_layers, self.margin = features, nbhds, T, n_layers, margin
        self.optimizer = torch.optim.Adam(list(self.params.values()), lr=lr)

    def train_batch(self, batch, forced=None):
        batch = torch.as_tensor(batch, dtype=torch.int64)
        hq = model_forward(self.params, self.features, batch[:, 0], self.nbhds, self.T, self.n_layers, forced)
        hp = model_forward(self.params, self.features, batch[:, 1], self.nbhds, self.T, self.n_layers, forced)
        hn = model_forward(self.params, self.features, batch[:, 2], self.nbhds, self.T, self.n_layers, forced)
        loss = max_margin_loss(hq, hp, hn, self.margin)
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return float(loss.detach())


# ----------------------------------------------------------------------------------------
# Batch construction and evaluation metrics
# ----------------------------------------------------------------------------------------

def check_batch_properties(batch, positives, n_items):
    """Properties the reference's `sample_batch` guarantees with easy negatives
    (pinsage_training.py:53-77): the B pairs are distinct rows of `positives`, and the
    negatives are distinct ids that appear in none of the batch's pairs."""
    batch = np.asarray(batch); positives = np.asarray(positives)
    pos_set = {tuple(r) for r in positives.tolist()}
    assert all(tuple(r) in pos_set for r in batch[:, :2].tolist())
    neg = batch[:, 2]
    assert neg.min() >= 0 and neg.max() < n_items
    assert len(set(neg.tolist())) == len(neg)
    assert not (set(neg.tolist()) & set(batch[:, :2].reshape(-1).tolist()))


def sample_batch_philox(positives, n_items, B, seed, step):
    """CPU restatement of ps_sample_batch (csrc/sampler.cu), which draws what the reference's
    sample_positives_with_rep + sample_easy_negatives draw (pinsage_training.py:53-77): B distinct uniformly
    random rows of `positives`, then B distinct uniformly random ids that are no node of those pairs.
    Candidate c of phase ph = floor(r64 * range / 2^64), r64 = x0 << 32 | x1 of
    Philox(counter=(c, ph, step_lo, step_hi), key=seed); first occurrences in stream order are kept."""
    positives = np.asarray(positives, dtype=np.int64)
    k0, k1 = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    s0, s1 = np.uint32(step & 0xFFFFFFFF), np.uint32((step >> 32) & 0xFFFFFFFF)

    def stream(phase, rng, forbidden):
        seen, out, c = set(forbidden), [], 0
        while len(out) < B:
            cs = np.arange(c, c + 4 * B + 64, dtype=np.uint32)
            x0, x1, _, _ = philox4x32_10(cs, np.full_like(cs, phase), np.full_like(cs, s0), np.full_like(cs, s1), k0, k1)
            for a, b in zip(x0.tolist(), x1.tolist()):
                v = (((a << 32) | b) * rng) >> 64
                if v not in seen:
                    seen.add(v); out.append(v)
                    if len(out) == B:
                        break
            c += 4 * B + 64
        return np.array(out, dtype=np.int64)

    rows = stream(0, positives.shape[0], ())
    pairs = positives[rows]
    neg = stream(1, n_items, set(pairs.reshape(-1).tolist()))
    return np.concatenate([pairs, neg[:, None]], axis=1)


def hit_rate(knn_mat, test_positives, K):
    """`hit_rate` (eval.py:227-238): fraction of test pairs (q, pos) with pos among the
    first K neighbours of q."""
    knn = np.asarray(knn_mat)[:, :K]
    tp = np.asarray(test_positives)
    hits = (knn[tp[:, 0]] == tp[:, 1:2]).any(1)
    return float(hits.sum()) / tp.shape[0]


def mrr(knn_mat, test_positives, K, scaling=1):
    """`mrr` (eval.py:240-250): mean of scaling / rank, rank 1-based, rank = K when the
    positive is absent from the first K neighbours."""
    knn = np.asarray(knn_mat)[:, :K]
    tp = np.asarray(test_positives)
    eq = knn[tp[:, 0]] == tp[:, 1:2]
    rank = np.where(eq.any(1), eq.argmax(1) + 1, K).astype(np.float64)
    return float((1.0 / (rank / scaling)).sum() / tp.shape[0])


def knn_from_emb(emb, q, k):
    """`cosine_sim_ab` + `knn_from_emb` (baselines.py:69-77,91-103):
    sim = q.E^T / (|q||e| + 1e-16); top-(k+1) per query, column 0 dropped (assumed self)."""
    emb = torch.as_tensor(emb, dtype=torch.float32)
    qe = emb[torch.as_tensor(q, dtype=torch.int64)]
    dots = qe @ emb.T
    lens = qe.norm(dim=1)[:, None] @ emb.norm(dim=1)[None, :] + 1e-16
    w, nidx = (dots / lens).topk(k + 1, dim=1, largest=True)
    return w[:, 1:], nidx[:, 1:]
