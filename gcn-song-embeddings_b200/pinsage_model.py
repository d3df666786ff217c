"""B200 drop-in for the reference's `pinsage_model` module (same public names, argument
meaning and return formats; bodies call libpinsage_b200.so -- there is no CPU path).

Reference: /root/reference/pinsage_model.py.  What changed underneath:
  * the walker draws from a counter-based Philox stream keyed by (seed, source, step)
    instead of torch's global mt19937, so traces are reproducible and parallel; the
    visit law (restart probability alpha, self entry zeroed after normalising, top-T) is
    the reference's (pinsage_model.py:32-53, 88-107);
  * sample_neighborhood_topt never materialises the dense [n, N+C] histogram;
  * ConvLayer / PinSageModel keep the reference's parameters and state-dict keys but run
    on compact per-layer buffers (see ps_engine.py).
Ties in the top-T are broken by (count desc, node id asc); slots beyond the number of
distinct visited nodes carry weight 0 and the source's own id (the reference returns
arbitrary zero-weight fillers there).
"""
from __future__ import annotations

import os
import time

import torch
import torch.nn as nn

import ps_native
from ps_engine import Engine, NeighborTable, PinSageFunction
from ps_graph import PSGraph, as_psgraph

DEF_T_PRECOMP = 100
DEF_HOPS = 500
DEF_ALPHA = 0.85

# Philox key of the walker.  Deterministic by default; call seed_walker() to change it.
_WALK_SEED = [0x5EED5EED]


def seed_walker(seed: int):
    """Set the Philox key used by the walker entry points of this module."""
    _WALK_SEED[0] = int(seed)


def _next_seed():
    s = _WALK_SEED[0]
    _WALK_SEED[0] = (s * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
    return s


def _n_items_of(g, n_items=None):
    if isinstance(g, PSGraph):
        return g.n_tracks
    if n_items is None:
        raise ValueError("n_items is required for a non-PSGraph graph")
    return n_items


def get_embeddings(h, nodeset, d):
    """pinsage_model.py:21-22."""
    return h[nodeset, :d]


def put_embeddings(h, nodeset, nodeset_new_h):
    """pinsage_model.py:24-30, kept for API compatibility.  The engine itself never calls
    it (no full-table clone on the hot path)."""
    new_h = h.clone().detach()
    new_h[nodeset, : nodeset_new_h.shape[1]] = nodeset_new_h.detach()
    new_h[nodeset, nodeset_new_h.shape[1]:] = 0
    return new_h


def do_random_walks(g, nodeset, n_hops, alpha, n_items=None, seed=None):
    """Restart random walk traces, int64 [len(nodeset), n_hops] (pinsage_model.py:32-53)."""
    pg = as_psgraph(g, _n_items_of(g, n_items))
    nodeset = torch.as_tensor(nodeset)
    out = ps_native.walk_topt(pg.device(), nodeset, n_hops, alpha, 1, _next_seed() if seed is None else seed,
                              want_i64=False, want_trace=True)
    trace = out["trace"].to(torch.int64)
    return trace if nodeset.is_cuda else trace.cpu()


def sample_neighborhood(g, n_items, nodeset, n_hops, alpha, seed=None):
    """Dense normalised visit counts float64 [len(nodeset), number_of_nodes] with the self
    entry zeroed (pinsage_model.py:88-101).  Compatibility entry point: it materialises the
    dense row the reference builds; the hot path uses sample_neighborhood_topt."""
    pg = as_psgraph(g, n_items)
    nodeset = torch.as_tensor(nodeset)
    src = nodeset.to("cuda", torch.int64)
    trace = do_random_walks(pg, src, n_hops, alpha, seed=seed)
    n = src.numel()
    counts = torch.zeros((n, pg.number_of_nodes()), dtype=torch.float64, device="cuda")
    counts.scatter_add_(1, trace, torch.ones_like(trace, dtype=torch.float64))
    prob = counts / counts.sum(1, keepdim=True)
    prob[torch.arange(n, device="cuda"), src] = 0
    return prob if nodeset.is_cuda else prob.cpu()


def sample_neighborhood_topt(g, n_items, nodeset, n_hops, alpha, T, seed=None):
    """(weights float64 [n, T], nodes int64 [n, T]): the T-sized PPR neighbourhoods of the
    nodes in nodeset (pinsage_model.py:103-107), walker and top-T fused on the device."""
    pg = as_psgraph(g, n_items)
    nodeset = torch.as_tensor(nodeset)
    out = ps_native.walk_topt(pg.device(), nodeset, n_hops, alpha, T, _next_seed() if seed is None else seed)
    w, nb = out["weights"], out["nodes"]
    return (w, nb) if nodeset.is_cuda else (w.cpu(), nb.cpu())


def precompute_neighborhoods_topt(g, n_items, n_hops, alpha, T, path, seed=None):
    """T-sized PPR neighbourhoods of ALL items; loads `path` when it holds tables of the
    right shape, else computes and saves them.  Returns (weights [N,T] f64, nodes [N,T]
    i64) like the reference (pinsage_model.py:109-132; both of its branches return
    weights first)."""
    if path is not None and os.path.isfile(path):
        weights, nodes = torch.load(path)
        if weights.shape[0] == n_items and weights.shape[1] == T:
            return (weights, nodes)
    pg = as_psgraph(g, n_items)
    t0 = time.time()
    out = ps_native.walk_topt(pg.device(), torch.arange(n_items, device="cuda"), n_hops, alpha, T,
                              _next_seed() if seed is None else seed, want_i32=True)
    weights, nodes = out["weights"].cpu(), out["nodes"].cpu()
    print(f"{n_items}/{n_items} done.\n{time.time() - t0}s elapsed.")
    if path is not None:
        torch.save((weights, nodes), path)  # plain tensors, the reference's file format (before anything is attached)
    # keep the engine-native device copy (int32 / float32) attached, so the trainer does not upload it again
    table = NeighborTable.__new__(NeighborTable)
    table.nodes, table.w, table.n, table.Tp, table.scratch = out["nodes_i32"], out["weights_f32"], n_items, T, {}
    weights._ps_table = table
    return (weights, nodes)


def sample_hard_negatives(g, n_items, visit_prob, hn_per_query, min_rank, max_rank):
    """(NOT USED by the reference either; pinsage_model.py:135-140.)"""
    rng = visit_prob.topk(max_rank, 1)[1][:, min_rank:]
    sample = torch.randint(0, rng.shape[1], (hn_per_query,), device=rng.device)
    return rng[:, sample]


def relevant_nodes_per_layer(g, n_items, nodeset, n_layers, n_hops, alpha, T):
    """Online variant: the walker runs per layer (pinsage_model.py:142-154)."""
    S = []
    cur = torch.as_tensor(nodeset)
    for _ in range(n_layers):
        nb_weights, nb_nodes = sample_neighborhood_topt(g, n_items, cur, n_hops, alpha, T)
        S.insert(0, (cur, nb_weights, nb_nodes))
        cur = torch.cat([nb_nodes.flatten(), cur]).unique()
    return S


def relevant_nodes_per_layer_precomp(nodeset, n_layers, T, nbhds):
    """List of (nodeset_l, weights_l, neighbours_l), bottom layer first
    (pinsage_model.py:156-168).  Index-only work on whatever device the inputs live on."""
    all_nb_weights, all_nb_nodes = nbhds
    S = []
    cur = torch.as_tensor(nodeset)
    for _ in range(n_layers):
        S.insert(0, (cur, all_nb_weights[cur, :T], all_nb_nodes[cur, :T]))
        cur = torch.cat([all_nb_nodes[cur, :T].flatten(), cur]).unique()
    return S


class ConvLayer(nn.Module):
    """A single PinSage convolution (pinsage_model.py:171-212).  Parameters, init law and
    state-dict keys (Q.weight, Q.bias, W.weight, W.bias) are the reference's."""

    def __init__(self, in_dim, out_dim, hidden_dim):
        super().__init__()
        self.in_dim, self.out_dim, self.hidden_dim = in_dim, out_dim, hidden_dim
        self.Q = nn.Linear(in_dim, hidden_dim)
        torch.nn.init.xavier_uniform_(self.Q.weight)
        self.Q.bias.data.fill_(0.3)
        self.W = nn.Linear(in_dim + hidden_dim, out_dim)
        torch.nn.init.xavier_uniform_(self.W.weight)
        self.W.bias.data.fill_(0.3)

    def forward(self, h, nodeset, nb_nodes, nb_weights):
        """Stand-alone layer call with the reference's signature: h is the full [N, >=in_dim]
        table, nodeset the targets, nb_nodes / nb_weights their [n, T] neighbourhoods.
        Inference-style (no autograd); training goes through PinSageModel."""
        dev_in = h.is_cuda
        hd = h.to("cuda", torch.float32).contiguous()
        ns = torch.as_tensor(nodeset).to("cuda", torch.int64)
        nb = torch.as_tensor(nb_nodes).to("cuda", torch.int64)
        w = torch.as_tensor(nb_weights).to("cuda", torch.float32).contiguous()
        n, T = nb.shape
        zr, inv = torch.unique(nb.reshape(-1), return_inverse=True)
        nz = zr.numel()
        z = torch.empty((nz, self.hidden_dim), dtype=torch.float32, device="cuda")
        ps_native.gemm(hd, self.Q.weight, z, nz, self.hidden_dim, self.in_dim, p_rows=zr.to(torch.int32),
                       bias=self.Q.bias, act=1)
        cat = torch.empty((n, self.in_dim + self.hidden_dim), dtype=torch.float32, device="cuda")
        inv_wsum = torch.empty((n,), dtype=torch.float32, device="cuda")
        ps_native.aggregate_fwd(hd, ns.to(torch.int32), self.in_dim, z, inv.view(n, T).to(torch.int32).contiguous(),
                                w, self.hidden_dim, cat, inv_wsum)
        out = torch.empty((n, self.out_dim), dtype=torch.float32, device="cuda")
        norm = torch.empty((n,), dtype=torch.float32, device="cuda")
        if self.out_dim <= 128:
            ps_native.gemm(cat, self.W.weight, out, n, self.out_dim, self.in_dim + self.hidden_dim,
                           bias=self.W.bias, act=1, l2norm=True, norm_out=norm)
        else:
            ps_native.gemm(cat, self.W.weight, out, n, self.out_dim, self.in_dim + self.hidden_dim, bias=self.W.bias, act=1)
            ps_native.l2norm_rows(out, norm)
        return out if dev_in else out.cpu()


class PinSageModel(nn.Module):
    """A PinSage model; forward() is one feed-forward step (pinsage_model.py:215-265).
    Parameters live on the current CUDA device.  `reference_compat` keeps the reference's
    duplicate-nodeset gradient factor (default, for parity); set it False for the plain
    sum-of-row-gradients."""

    def __init__(self, g, n_items, n_layers, dimensions, n_hops, alpha, T, nbhds):
        super().__init__()
        self.g, self.n_items, self.T, self.n_hops, self.alpha, self.nbhds = g, n_items, T, n_hops, alpha, nbhds
        self.n_layers = n_layers
        self.in_dim, self.hidden_dim, self.out_dim = dimensions[0], dimensions[1], dimensions[2]
        if self.in_dim < self.out_dim:
            raise ValueError("in_dim must be >= out_dim (the reference zero-pads new rows to the table width, pinsage_model.py:27)")
        self.in_dim_per_layer = [self.in_dim] + [self.out_dim for _ in range(n_layers - 1)]
        self.conv_layers = nn.ModuleList(
            ConvLayer(self.in_dim_per_layer[i], self.out_dim, self.hidden_dim) for i in range(n_layers))
        self.G1 = nn.Linear(self.out_dim, self.out_dim)
        torch.nn.init.xavier_uniform_(self.G1.weight)
        self.G1.bias.data.fill_(0.3)
        self.G2 = nn.Linear(self.out_dim, self.out_dim, bias=False)
        torch.nn.init.xavier_uniform_(self.G2.weight)
        self.reference_compat = True
        ps_native._ensure_device()  # fail loudly here, not at the first forward
        if nbhds is None:  # online sampling: the walker runs inside every forward (pinsage_model.py:249-250, commented out there)
            self.nbhds = self.online_neighbors()
        self.to("cuda")
        self._engine = Engine(self)

    @property
    def engine(self) -> Engine:
        return self._engine

    def online_neighbors(self, seed=None):
        """A neighbourhood provider that samples with the walker on demand (assign it to `self.nbhds`)."""
        from ps_engine import OnlineNeighbors
        pg = as_psgraph(self.g, self.n_items)
        return OnlineNeighbors(pg.device(), self.n_items, self.n_hops, self.alpha, _next_seed() if seed is None else seed)

    def forward(self, initial_h, nodeset):
        feats = self._engine.features(initial_h)
        ns = torch.as_tensor(nodeset).to("cuda", torch.int64)
        if ns.numel() and (int(ns.max()) >= self.n_items or int(ns.min()) < 0):
            raise IndexError("nodeset holds ids outside [0, n_items)")
        params = [p for _, p in self.named_parameters()]
        out = PinSageFunction.apply(self._engine, feats, ns, self.reference_compat, *params)
        return out if initial_h.is_cuda else out.cpu()
