"""Device-side PinSage engine: frontier plans, forward, backward and the fused train step,
composed from the kernels of libpinsage_b200.so.

What it replaces in the reference (all under /root/reference):
  * relevant_nodes_per_layer_precomp      pinsage_model.py:156-168   -> build_plan
  * ConvLayer.forward / PinSageModel.forward  :189-212, 246-265      -> Engine.forward
  * autograd through those + put/get_embeddings :21-30               -> Engine.backward
  * PinSage.train_batch's three forwards + loss  pinsage_training.py:184-190 -> Engine.train_step

Design differences (results identical, see tests/):
  * no full-table clones (put_embeddings): every layer writes a compact activation buffer
    indexed by frontier position;
  * Q is applied once per DISTINCT input row of a layer and the T-neighbour aggregation
    gathers the transformed rows (the reference transforms n*T gathered rows, repeats
    included, pinsage_model.py:193-201) -- per row the arithmetic is the same;
  * the q / pos / neg forwards of a batch share one frontier (a node's embedding does not
    depend on which column asked for it); the reference's duplicate-nodeset gradient
    factor is applied explicitly in the loss kernel;
  * the aggregation backward is a segmented gather over a per-step transpose of the
    (target, neighbour) relation instead of dense [N, D] index_add buffers.
"""
from __future__ import annotations

import os
import threading
import time
from dataclasses import dataclass, field
from typing import List, Optional

import torch

import ps_native as nat


# PS_PREP_TIMING=1: host-side timing of Engine.prepare (development aid; adds stream syncs)
_PREP_TIMING = {"sample": 0.0, "unique": 0.0, "plan_host": 0.0, "plan_drain": 0.0, "n": 0} if os.environ.get("PS_PREP_TIMING") else None


def _dev(t, dtype=None, device="cuda"):
    t = t.to(device=device, non_blocking=True)
    return t.to(dtype) if dtype is not None and t.dtype != dtype else t


class NeighborTable:
    """Precomputed neighbourhoods in engine-native layout: nodes int32 [N, Tp], weights
    float32 [N, Tp] in HBM.  Built from the reference-format tuple
    (weights float64 [N, Tp], nodes int64 [N, Tp]) of precompute_neighborhoods_topt
    (pinsage_model.py:119-132)."""

    def __init__(self, weights: torch.Tensor, nodes: torch.Tensor, device="cuda"):
        if weights.shape != nodes.shape or weights.dim() != 2:
            raise ValueError("nbhds must be (weights [N,T], nodes [N,T])")
        self.nodes = _dev(nodes, torch.int32, device).contiguous()
        self.w = _dev(weights, None, device).to(torch.float32).contiguous()
        self.n, self.Tp = self.nodes.shape
        self.scratch = {}
        self._sanitize()

    def _sanitize(self):
        """One-time range check of the neighbour ids (the plan builder indexes dense [N] maps with them).  A
        neighborhoods.pt written by the reference can hold zero-weight FILLER ids >= n_items: its topk runs over the
        dense [n, N+C] row (pinsage_model.py:93-107).  Those slots are remapped to the
        row's own id (weight 0: they contribute nothing and add no frontier node); an out-of-range id with a
        non-zero weight raises IndexError, as indexing the reference's feature table with it would."""
        if self.nodes.numel() == 0:
            return
        bad = (self.nodes < 0) | (self.nodes >= self.n)
        if bool(bad.any()):
            if bool((bad & (self.w != 0)).any()):
                raise IndexError("neighbourhood table holds node ids outside [0, n_items) with non-zero weight")
            own = torch.arange(self.n, dtype=torch.int32, device=self.nodes.device)[:, None].expand_as(self.nodes)
            self.nodes = torch.where(bad, own, self.nodes).contiguous()

    def lookup(self, cur: torch.Tensor, T: int):
        """(neighbours int32 [n, T], weights float32 [n, T]) of the nodes `cur` (int64 on the device)."""
        return self.nodes[cur, :T], self.w[cur, :T].contiguous()

    @classmethod
    def of(cls, nbhds) -> "NeighborTable":
        if isinstance(nbhds, (NeighborTable, OnlineNeighbors)):
            return nbhds
        weights, nodes = nbhds
        key = "_ps_table"
        cached = getattr(weights, key, None)
        if cached is None or cached.nodes.shape != nodes.shape:
            cached = cls(weights, nodes)
            try:
                setattr(weights, key, cached)
            except Exception:
                pass
        return cached


class OnlineNeighbors:
    """Neighbourhoods sampled on demand: the walker (ps_walk_topt) runs inside every forward on the nodes of each
    layer's frontier, like the reference's online `relevant_nodes_per_layer` (pinsage_model.py:142-154, the
    "sample" of "sample + fwd + bwd").  Same interface as NeighborTable.  One difference from the reference: the
    q / pos / neg forwards of a training batch share one frontier here, so a node gets ONE fresh neighbourhood per
    step instead of one per column / layer it appears in (the Philox key changes once per plan; within a plan the
    walks of a node do not depend on which layer asks, because draws are keyed by (seed, source, step))."""

    def __init__(self, graph_handle, n_items: int, n_hops: int, alpha: float, seed: int = 0x0417E5EED):
        self.graph, self.n, self.n_hops, self.alpha = graph_handle, int(n_items), int(n_hops), float(alpha)
        self.Tp = 1 << 30
        self.scratch = {}
        self._seed = int(seed)

    def new_plan(self):
        """Fresh walks for the next plan (called by build_plan)."""
        self._seed = (self._seed * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF

    def lookup(self, cur: torch.Tensor, T: int):
        with nat._Timed("walk_topt_online", 0.0, 28.0 * cur.numel() * self.n_hops):
            out = nat.walk_topt(self.graph, cur, self.n_hops, self.alpha, T, self._seed, want_i64=False, want_i32=True)
        return out["nodes_i32"], out["weights_f32"]

    def materialize(self, T: int) -> NeighborTable:
        """One walker pass over all items -> a fixed table (full-graph inference needs consistent neighbourhoods)."""
        nodes, w = self.lookup(torch.arange(self.n, device="cuda"), T)
        table = NeighborTable.__new__(NeighborTable)
        table.nodes, table.w, table.n, table.Tp, table.scratch = nodes, w, self.n, T, {}
        return table


@dataclass
class LayerPlan:
    n: int                       # targets of this layer
    nz: int                      # rows of the layer input that get the Q transform
    self_rows: torch.Tensor      # int32 [n]   row of each target in the layer input
    nbz: torch.Tensor            # int32 [n,T] row of each neighbour in Z
    w: torch.Tensor              # float32 [n,T]
    zrows: Optional[torch.Tensor]  # int32 [nz] gather index into the feature table (layer 0) or None
    seg_off: Optional[torch.Tensor] = None  # int32 [nz+1]  (backward)
    pair_q: Optional[torch.Tensor] = None   # int32 [n*T]   (backward)
    chunk_off: Optional[torch.Tensor] = None  # int32 [nz+1] first work chunk of every z-row (backward)
    chunk_row: Optional[torch.Tensor] = None  # int32 [max_chunks] z-row of every work chunk (backward)
    nodes: Optional[torch.Tensor] = None      # int64 [n] node id of every target (sorted, distinct)


@dataclass
class Plan:
    top: torch.Tensor            # int64 [n_top] sorted distinct nodes whose embeddings are produced
    layers: List[LayerPlan] = field(default_factory=list)
    arena: Optional[torch.Tensor] = None  # ps_prepare_plan: the one device buffer every tensor above is a view of


def _unique_inverse(ids: torch.Tensor, n_ids: int, scratch: dict):
    """Sorted distinct values of `ids` (all in [0, n_ids)) and the position of every element among them.
    Same result as torch.unique(ids, return_inverse=True) but through a dense flag/prefix-sum map over the id
    space instead of a radix sort of the (much longer) id list."""
    if ids.numel() < 65536 or n_ids > 64 * ids.numel():
        uniq, inv = torch.unique(ids, return_inverse=True)
        return uniq.to(torch.int64), inv
    key = ("flag", torch.cuda.current_stream().cuda_stream if ids.is_cuda else 0)  # one scratch map per stream
    flag = scratch.get(key)
    if flag is None or flag.numel() != n_ids or flag.device != ids.device:
        flag = scratch[key] = torch.empty(n_ids, dtype=torch.int32, device=ids.device)
    flag.zero_()
    idx = ids.to(torch.int64)
    flag[idx] = 1
    pos = torch.cumsum(flag, 0, dtype=torch.int32)
    uniq = flag.nonzero().squeeze(1)
    inv = (pos[idx] - 1).to(torch.int64)
    return uniq, inv


def prepare_native(batch: torch.Tensor, n_layers: int, T: int, table: NeighborTable, need_backward: bool = True):
    """(Plan, triples int32 [B,3], counts int32 [3,U]) of a batch int64 [B,3] on the device through ONE host call
    (ps_prepare_plan, csrc/plan.cu): what torch.unique + build_plan + count_triples compose from ~25 calls."""
    if T > table.Tp:
        raise ValueError(f"T={T} exceeds the precomputed neighbourhood width {table.Tp}")
    if isinstance(table, OnlineNeighbors):  # the walker runs on every layer's targets inside the same host call
        table.new_plan()
        with nat._Timed("prepare_plan_online"):
            arena, base, d = nat.prepare_plan(batch, None, None, T, n_layers, need_backward,
                                              online=(table.graph, table.n, table.n_hops, table.alpha, table._seed))
    else:
        arena, base, d = nat.prepare_plan(batch, table.nodes, table.w, T, n_layers, need_backward)
    B = batch.shape[0]

    def view(off, count, dtype, shape=None):
        if off < 0:
            return None
        t = arena[base + off: base + off + count * dtype.itemsize].view(dtype)
        return t.view(shape) if shape is not None else t

    U = int(d.U)
    top = view(d.off_top, U, torch.int64)
    plan = Plan(top=top, layers=[None] * n_layers, arena=arena)
    for l in range(n_layers):
        c = d.layers[l]
        n, nz = int(c.n), int(c.nz)
        max_chunks = max(1, n * T // nat.AGG_BWD_CHUNK + nz)
        plan.layers[l] = LayerPlan(
            n=n, nz=nz, self_rows=view(c.off_self_rows, n, torch.int32), nbz=view(c.off_nbz, n * T, torch.int32, (n, T)),
            w=view(c.off_w, n * T, torch.float32, (n, T)), zrows=view(c.off_zrows, nz, torch.int32),
            seg_off=view(c.off_seg_off, nz + 1, torch.int32), pair_q=view(c.off_pair_q, n * T, torch.int32),
            chunk_off=view(c.off_chunk_off, nz + 1, torch.int32), chunk_row=view(c.off_chunk_row, max_chunks, torch.int32),
            nodes=view(c.off_nodes, n, torch.int64))
    triples = view(d.off_triples, 3 * B, torch.int32, (B, 3))
    counts = view(d.off_counts, 3 * U, torch.int32, (3, U))
    return plan, triples, counts


def build_plan(top: torch.Tensor, n_layers: int, T: int, table: NeighborTable, need_backward: bool, check_ids: bool = True) -> Plan:
    """T-hop computation graph around the distinct nodes `top` (sorted int64, on device).
    Restates relevant_nodes_per_layer_precomp (pinsage_model.py:156-168): layer l-1's
    targets are unique(neighbours of layer l's targets + the targets themselves)."""
    if T > table.Tp:
        raise ValueError(f"T={T} exceeds the precomputed neighbourhood width {table.Tp}")
    if check_ids and top.numel():
        lo, hi = top[[0, -1]].tolist()  # top is sorted: one host read
        if hi >= table.n or lo < 0:
            raise IndexError("node id out of range")  # the reference raises IndexError on OOB ids too
    plan = Plan(top=top, layers=[None] * n_layers)
    if hasattr(table, "new_plan"):
        table.new_plan()
    cur = top
    for l in reversed(range(n_layers)):
        n = cur.numel()
        nb, w = table.lookup(cur, T)
        if nb.is_cuda and n > 0:
            # native plan builder (csrc/plan.cu): dense flag map + scan for the next frontier, one radix sort for the
            # backward transpose; a handful of launches and one host read per layer
            nb = nb.contiguous()
            uniq, nbz, self_rows, nz = nat.plan_layer(nb, cur, l > 0, table.n)
            if l > 0:
                nxt, zrows = uniq, None
            else:
                nxt, zrows, self_rows = None, uniq, cur.to(torch.int32)
            lp = LayerPlan(n=n, nz=nz, self_rows=self_rows, nbz=nbz, w=w, zrows=zrows, nodes=cur)
            if need_backward:
                lp.pair_q, lp.seg_off, lp.chunk_off, lp.chunk_row = nat.plan_transpose(nbz, nz)
            plan.layers[l] = lp
            cur = nxt
            continue
        if l > 0:
            allv = torch.cat([nb.reshape(-1).to(torch.int64), cur])
            nxt, inv = _unique_inverse(allv, table.n, table.scratch)
            nbz = inv[: n * T].view(n, T).to(torch.int32).contiguous()
            self_rows = inv[n * T:].to(torch.int32).contiguous()
            zrows, nz = None, nxt.numel()
        else:
            zr, inv = _unique_inverse(nb.reshape(-1), table.n, table.scratch)
            nbz = inv.view(n, T).to(torch.int32).contiguous()
            self_rows = cur.to(torch.int32).contiguous()
            zrows, nz, nxt = zr.to(torch.int32).contiguous(), zr.numel(), None
        lp = LayerPlan(n=n, nz=nz, self_rows=self_rows, nbz=nbz, w=w, zrows=zrows, nodes=cur)
        if need_backward:
            flat = nbz.reshape(-1)
            skeys, order = torch.sort(flat, stable=True)  # q ascends inside a segment: reproducible backward sums
            lp.pair_q = order.to(torch.int32).contiguous()
            # segment starts of the sorted keys (no bincount: it reads its maximum back to the host)
            lp.seg_off = torch.searchsorted(skeys, torch.arange(nz + 1, dtype=torch.int32, device=flat.device)).to(torch.int32)
            lp.chunk_off = nat.aggregate_bwd_chunks(lp.seg_off)
            lp.chunk_row = nat.aggregate_bwd_chunk_rows(lp.chunk_off, flat.numel(), nz)
        plan.layers[l] = lp
        cur = nxt
    return plan


@dataclass
class Prepared:
    """A training batch with everything index-only already built (Engine.prepare)."""
    batch: torch.Tensor      # int64 [B,3] on device
    triples: torch.Tensor    # int32 [B,3] rows of the distinct-node embedding matrix
    plan: Plan
    counts: torch.Tensor     # int32 [3,U] occurrences of every distinct node per column
    ready: object            # CUDA event recorded on the side stream

    def tensors(self):
        if self.plan.arena is not None:  # one storage behind every view
            return [self.batch, self.plan.arena]
        out = [self.batch, self.triples, self.counts, self.plan.top]
        for lp in self.plan.layers:
            out += [t for t in (lp.self_rows, lp.nbz, lp.w, lp.zrows, lp.seg_off, lp.pair_q, lp.chunk_off, lp.chunk_row, lp.nodes) if t is not None]
        return out


def _splits_for(M, N, K):
    tiles = -(-M // 128) * -(-N // 128)
    want = -(-148 * 4 // tiles)
    return max(1, min(want, -(-K // 256)))


class Engine:
    """Forward / backward over a PinSageModel's parameters (read in place, by pointer)."""

    def __init__(self, model):
        self.model = model
        self._feat_src = None
        self._feat_dev = None
        self._tls = threading.local()

    # ---- inputs ---------------------------------------------------------------------
    def features(self, features: torch.Tensor) -> torch.Tensor:
        """Feature table resident in HBM (uploaded once per distinct host tensor)."""
        if features.is_cuda and features.dtype == torch.float32 and features.is_contiguous():
            return features
        key = (features.data_ptr(), tuple(features.shape), features._version)
        if self._feat_src != key:
            self._feat_dev = _dev(features, torch.float32).contiguous()
            self._feat_src = key
        return self._feat_dev

    def _dims(self):
        m = self.model
        return m.in_dim_per_layer, m.hidden_dim, m.out_dim

    # ---- forward ---------------------------------------------------------------------
    def forward(self, feats: torch.Tensor, plan: Plan, keep: bool):
        """Embeddings [n_top, out_dim] of plan.top.  keep=True saves what backward needs."""
        m = self.model
        in_dims, dh, do = self._dims()
        if feats.shape[1] < in_dims[0] or in_dims[0] % 4 or dh % 4 or do % 4:
            raise ValueError("feature / hidden / output dims must be multiples of 4 and features at least in_dim wide")
        saved = []
        h_prev = feats
        for l, lp in enumerate(plan.layers):
            conv = m.conv_layers[l]
            din = in_dims[l]
            if not keep and do <= 128 and do < dh:
                h_prev = self._conv_projected(h_prev, din, lp.nz, lp.zrows, lp.n, lp.self_rows, lp.nbz, lp.w, conv, l)
                continue
            z = torch.empty((lp.nz, dh), dtype=torch.float32, device="cuda")
            # training: the GEMM also records sign(z) (1 bit per element) so the backward never re-reads z for leaky'
            zmask = None
            if keep and nat.gemm_mask_supported(lp.nz, dh, din) and nat.gemm_mask_supported(lp.nz, dh, do):
                zmask = torch.empty((lp.nz, dh // 32), dtype=torch.int32, device="cuda")
            nat.gemm(h_prev, conv.Q.weight, z, lp.nz, dh, din, p_rows=lp.zrows, bias=conv.Q.bias, act=1, mask=zmask, tag=f"gemm_q_fwd_l{l}")
            cat = torch.empty((lp.n, din + dh), dtype=torch.float32, device="cuda")
            inv_wsum = torch.empty((lp.n,), dtype=torch.float32, device="cuda")
            nat.aggregate_fwd(h_prev, lp.self_rows, din, z, lp.nbz, lp.w, dh, cat, inv_wsum, tag=f"aggregate_fwd_l{l}")
            h = torch.empty((lp.n, do), dtype=torch.float32, device="cuda")
            norm = torch.empty((lp.n,), dtype=torch.float32, device="cuda")
            if do <= 128:
                nat.gemm(cat, conv.W.weight, h, lp.n, do, din + dh, bias=conv.W.bias, act=1, l2norm=True, norm_out=norm, tag=f"gemm_w_fwd_l{l}")
            else:
                nat.gemm(cat, conv.W.weight, h, lp.n, do, din + dh, bias=conv.W.bias, act=1)
                nat.l2norm_rows(h, norm)
            if keep:
                saved.append((h_prev, z, cat, inv_wsum, h, norm, zmask))
            h_prev = h
        n_top = plan.layers[-1].n
        a1 = torch.empty((n_top, do), dtype=torch.float32, device="cuda")
        nat.gemm(h_prev, m.G1.weight, a1, n_top, do, do, bias=m.G1.bias, act=1)
        out = torch.empty((n_top, do), dtype=torch.float32, device="cuda")
        nat.gemm(a1, m.G2.weight, out, n_top, do, do)
        ctx = (plan, saved, a1) if keep else None
        return out, ctx

    # ---- inference form of a layer: aggregate AFTER the W projection ------------------------------------------
    def _projected_rows(self, h_in, din, nz, zrows, conv, l, row_block=1 << 21, out=None):
        """zp [nz, do] = leaky(Q x + b) . W[:, din:]^T for the nz needed input rows, in row blocks so that the dh-wide z
        (2 KB per row at dh = 512: 41 GB for 20 M rows) is never materialised as a whole."""
        _, dh, do = self._dims()
        w2 = conv.W.weight.detach()[:, din:]  # [do, dh] view, leading dimension din + dh
        zp = out if out is not None else torch.empty((nz, do), dtype=torch.float32, device="cuda")
        for b0 in range(0, nz, row_block):
            nb = min(row_block, nz - b0)
            z = torch.empty((nb, dh), dtype=torch.float32, device="cuda")
            rows = zrows[b0:b0 + nb] if zrows is not None else None
            src = h_in if zrows is not None else h_in[b0:b0 + nb]
            nat.gemm(src, conv.Q.weight, z, nb, dh, din, p_rows=rows, bias=conv.Q.bias, act=1, tag=f"gemm_q_fwd_l{l}")
            nat.gemm(z, w2, zp[b0:b0 + nb], nb, do, dh, tag=f"gemm_proj_fwd_l{l}")
            del z
        return zp

    def _projected_weight(self, conv, din):
        """[W[:, :din] | I]: the weight that turns cat' = [x_self | aggregated projections] into W x_cat of the reference
        layer (the identity block multiplies exactly under the hi/lo split: hi(1) = 1, lo(1) = 0)."""
        _, _, do = self._dims()
        w = conv.W.weight.detach()
        return torch.cat([w[:, :din], torch.eye(do, dtype=torch.float32, device=w.device)], dim=1).contiguous()

    def _conv_projected(self, h_in, din, nz, zrows, n, self_rows, nbz, w, conv, l):
        """One ConvLayer without a backward (pinsage_model.py:189-212), with the importance-weighted mean taken AFTER the
        neighbour part of the W projection: W [x | mean_t z_t] = W1 x + mean_t (W2 z_t).  The gathered rows are do wide
        instead of dh wide (4x fewer bytes at 512 / 128, and the aggregation is the HBM-bound part of inference); per
        row the result differs from the training-form layer by fp32 summation order only (~1e-7)."""
        _, dh, do = self._dims()
        zp = self._projected_rows(h_in, din, nz, zrows, conv, l)
        cat = torch.empty((n, din + do), dtype=torch.float32, device="cuda")
        inv_wsum = torch.empty((n,), dtype=torch.float32, device="cuda")
        nat.aggregate_fwd(h_in, self_rows, din, zp, nbz, w, do, cat, inv_wsum, tag=f"aggregate_fwd_l{l}")
        h = torch.empty((n, do), dtype=torch.float32, device="cuda")
        nat.gemm(cat, self._projected_weight(conv, din), h, n, do, din + do, bias=conv.W.bias, act=1, l2norm=True, tag=f"gemm_w_fwd_l{l}")
        return h

    # ---- backward --------------------------------------------------------------------
    def backward(self, ctx, d_out: torch.Tensor, grads: dict):
        """Accumulate parameter gradients of sum(out * d_out) into `grads` (name -> fp32
        tensor shaped like the parameter, caller-zeroed).  Consumes ctx (Z is overwritten)."""
        plan, saved, a1 = ctx
        m = self.model
        in_dims, dh, do = self._dims()
        n_top = plan.layers[-1].n
        h_top = saved[-1][4]
        d_out = d_out.contiguous()
        # head: out = G2(leaky(G1 h + b1))
        nat.gemm(d_out, a1, grads["G2.weight"], do, do, n_top, p_kmajor=False, q_kmajor=False,
                 accumulate=True, splits=_splits_for(do, do, n_top))
        d_a1 = torch.empty_like(a1)
        nat.gemm(d_out, m.G2.weight, d_a1, n_top, do, do, q_kmajor=False)
        nat.leaky_bwd(a1, d_a1)
        nat.gemm_wgrad(d_a1, h_top, grads["G1.weight"], do, do, n_top, splits=_splits_for(do, do, n_top), bias_grad=grads["G1.bias"])
        d_h = torch.empty((n_top, do), dtype=torch.float32, device="cuda")
        nat.gemm(d_a1, m.G1.weight, d_h, n_top, do, do, q_kmajor=False)

        for l in reversed(range(len(plan.layers))):
            lp = plan.layers[l]
            conv = m.conv_layers[l]
            din = in_dims[l]
            h_in, z, cat, inv_wsum, h, norm, zmask = saved[l]
            pre = f"conv_layers.{l}."
            d_pre = torch.empty((lp.n, do), dtype=torch.float32, device="cuda")
            nat.norm_leaky_bwd(h, norm, d_h, d_pre)
            # weight AND bias gradient in one call: the bias gradient (column sums of the upstream gradient) comes from
            # the operand tiles the GEMM's producers already hold
            nat.gemm_wgrad(d_pre, cat, grads[pre + "W.weight"], do, din + dh, lp.n, splits=_splits_for(do, din + dh, lp.n),
                           bias_grad=grads[pre + "W.bias"], tag=f"gemm_w_wgrad_l{l}")
            # Backward of the aggregation.  d_cat[:, din:] = d_pre . W[:, din:] is linear in d_pre, so the weighted
            # segment sum runs on the do-wide d_pre rows (4x fewer gathered bytes than the dh-wide d_cat rows, and
            # d_pre fits in L2) and ONE GEMM applies the W block per z-row, with leaky'(z) fused into its store:
            #   S[u]      = sum_{(t,s) -> u} w[t,s] / wsum[t] * d_pre[t]                    [nz, do]
            #   dZ_pre[u] = (S[u] . W[:, din:]) * leaky'(z[u])                              [nz, dh]   (in place in z)
            s_buf = torch.empty((lp.nz, do), dtype=torch.float32, device="cuda")
            nat.aggregate_bwd(d_pre, 0, do, lp.seg_off, lp.pair_q, lp.w, inv_wsum, lp.w.shape[1], s_buf, chunk_off=lp.chunk_off,
                              chunk_row=lp.chunk_row, apply_leaky=False, tag=f"aggregate_bwd_l{l}")
            nat.gemm(s_buf, conv.W.weight.detach()[:, din:], z, lp.nz, dh, do, q_kmajor=False, act=2, mask=zmask,
                     tag=f"gemm_agg_dgrad_l{l}")  # z := dZ_pre (leaky' from the recorded sign bits, else from z itself)
            nat.gemm_wgrad(z, h_in, grads[pre + "Q.weight"], dh, din, lp.nz, x_rows=lp.zrows, splits=_splits_for(dh, din, lp.nz),
                           bias_grad=grads[pre + "Q.bias"], tag=f"gemm_q_wgrad_l{l}")
            if l > 0:
                d_h = torch.empty((lp.nz, din), dtype=torch.float32, device="cuda")
                nat.gemm(z, conv.Q.weight, d_h, lp.nz, din, dh, q_kmajor=False, tag=f"gemm_q_dgrad_l{l}")
                d_self = torch.empty((lp.n, din), dtype=torch.float32, device="cuda")  # d_cat[:, :din]: the self-row gradient
                nat.gemm(d_pre, conv.W.weight.detach()[:, :din], d_self, lp.n, din, do, q_kmajor=False, tag=f"gemm_w_dgrad_l{l}")
                nat.scatter_add_rows(d_self, lp.self_rows, d_h, din)
        return grads

    def zero_grads(self, zero: bool = True):
        """Dict name -> zeroed gradient tensor.  All gradients are views into ONE flat fp32
        buffer (self.flat_grad) installed as the parameters' .grad, so the data-parallel
        exchange is a single allreduce and zeroing a single memset."""
        params = list(self.model.named_parameters())
        total = sum(p.numel() for _, p in params)
        flat = getattr(self, "flat_grad", None)
        if flat is None or flat.numel() != total or flat.device != params[0][1].device:
            flat = self.flat_grad = torch.zeros(total, dtype=torch.float32, device=params[0][1].device)
        elif zero:
            flat.zero_()
        grads, off = {}, 0
        for name, p in params:
            view = flat[off: off + p.numel()].view_as(p)
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                p.grad = view
            grads[name] = view
            off += p.numel()
        return grads

    # ---- fused training step -----------------------------------------------------------
    def prepare(self, batch, sampler=None) -> "Prepared":
        """Index-only preparation of a training batch on a high-priority side stream, so it overlaps the
        previous step's kernels: the distinct nodes of the batch, the layer plans, the backward transposes and
        the duplicate counts.  None of it depends on the model weights.  `batch` is an int64 [B,3] tensor (host
        tensors are copied on the side stream; device tensors make the side stream wait for the current one) or
        None with `sampler` a callable that draws the batch on the side stream."""
        m = self.model
        main = torch.cuda.current_stream()
        tls = self._tls
        if getattr(tls, "plan_stream", None) is None:
            tls.plan_stream = torch.cuda.Stream(priority=-1)  # one high-priority stream per preparing host thread
        side = tls.plan_stream
        if batch is not None and batch.is_cuda:
            side.wait_stream(main)
        check_ids = True
        if batch is None:
            check_ids = False  # drawn from `positives` / `all_ids` below: in range by construction
        elif not batch.is_cuda:
            if batch.numel() and (int(batch.max()) >= NeighborTable.of(m.nbhds).n or int(batch.min()) < 0):
                raise IndexError("node id out of range")
            check_ids = False  # checked on the host copy, no device read needed
        timing = _PREP_TIMING
        t0 = time.perf_counter() if timing is not None else 0.0
        with torch.cuda.stream(side):
            if batch is None:
                batch = sampler()
            batch = batch.to("cuda", torch.int64, non_blocking=True)
            B = batch.shape[0]
            if timing is not None:
                side.synchronize(); t1 = time.perf_counter(); timing["sample"] += t1 - t0
            table = NeighborTable.of(m.nbhds)
            native_ok = isinstance(table, OnlineNeighbors) or (isinstance(table, NeighborTable) and table.nodes.is_cuda)
            if native_ok and os.environ.get("PS_PY_PREPARE") != "1":
                plan, triples, counts = prepare_native(batch.contiguous(), m.n_layers, m.T, table)  # one host call
                t2 = t1 if timing is not None else 0.0
            else:  # PS_PY_PREPARE=1 (or a host-resident table): composed from Python
                top, inv = torch.unique(batch.reshape(-1), return_inverse=True)
                if timing is not None:
                    t2 = time.perf_counter(); timing["unique"] += t2 - t1
                triples = inv.view(B, 3).to(torch.int32).contiguous()
                plan = build_plan(top, m.n_layers, m.T, table, need_backward=True, check_ids=check_ids)
                counts = torch.empty((3, top.numel()), dtype=torch.int32, device="cuda")
                nat.count_triples(triples, top.numel(), counts)
            ready = torch.cuda.Event()
            ready.record(side)
            if timing is not None:
                t3 = time.perf_counter(); timing["plan_host"] += t3 - t2
                side.synchronize(); timing["plan_drain"] += time.perf_counter() - t3; timing["n"] += 1
        return Prepared(batch=batch, triples=triples, plan=plan, counts=counts, ready=ready)

    def train_step(self, feats: torch.Tensor, batch, margin: float, reference_compat: bool = True, diagnostics: bool = False):
        """Forward of the distinct nodes of a batch, max-margin loss, backward into the parameters' .grad.
        `batch` is an int64 [B,3] tensor or a Prepared from prepare().  Returns (loss [1], embeddings of the
        distinct nodes [U, out], triples int32 [B,3] indexing them), all on the device.  The whole step is ONE host
        call (ps_train_step, csrc/step.cu); diagnostics=True also leaves [node-feature triplet loss, batch variance]
        (pinsage_training.py:200-212) in self.last_diag.  PS_PY_STEP=1 composes the same kernels from Python instead
        (Engine.forward / backward, what model(features, nodeset) + autograd use)."""
        prep = batch if isinstance(batch, Prepared) else self.prepare(batch)
        main = torch.cuda.current_stream()
        main.wait_event(prep.ready)
        for t in prep.tensors():
            t.record_stream(main)  # allocated on the side stream, consumed here
        if os.environ.get("PS_PY_STEP") != "1":
            return self._train_step_native(feats, prep, margin, reference_compat, diagnostics)
        self.last_diag = None
        out, ctx = self.forward(feats, prep.plan, keep=True)
        loss = torch.zeros(1, dtype=torch.float32, device="cuda")
        d_out = torch.zeros_like(out)
        nat.margin_loss_fwd_bwd(out, prep.triples, margin, 1.0, prep.counts if reference_compat else None, loss, d_out)
        grads = self.zero_grads()
        self.backward(ctx, d_out, grads)
        return loss, out, prep.triples

    def _train_step_native(self, feats, prep, margin, reference_compat, diagnostics):
        m = self.model
        in_dims, dh, do = self._dims()
        if feats.shape[1] < in_dims[0] or in_dims[0] % 4 or dh % 4 or do % 4:
            raise ValueError("feature / hidden / output dims must be multiples of 4 and features at least in_dim wide")
        L = len(prep.plan.layers)
        if L > nat.PS_MAX_LAYERS:
            raise ValueError(f"at most {nat.PS_MAX_LAYERS} layers")
        grads = self.zero_grads(zero=False)  # ps_train_step zeroes the flat buffer itself
        a = nat.StepArgsC()
        a.n_layers, a.T, a.in_dim, a.hidden_dim, a.out_dim = L, prep.plan.layers[0].nbz.shape[1], in_dims[0], dh, do
        a.feats, a.ld_feats = feats.data_ptr(), feats.stride(0)
        ptr = lambda t: t.data_ptr() if t is not None else None
        for l, lp in enumerate(prep.plan.layers):
            c = a.layers[l]
            c.n, c.nz = lp.n, lp.nz
            c.self_rows, c.nbz, c.w, c.zrows = ptr(lp.self_rows), ptr(lp.nbz), ptr(lp.w), ptr(lp.zrows)
            c.seg_off, c.pair_q, c.chunk_off, c.chunk_row = ptr(lp.seg_off), ptr(lp.pair_q), ptr(lp.chunk_off), ptr(lp.chunk_row)
            conv, pre, q = m.conv_layers[l], f"conv_layers.{l}.", a.params[l]
            q.Qw, q.Qb, q.Ww, q.Wb = conv.Q.weight.data_ptr(), conv.Q.bias.data_ptr(), conv.W.weight.data_ptr(), conv.W.bias.data_ptr()
            q.gQw, q.gQb, q.gWw, q.gWb = (grads[pre + k].data_ptr() for k in ("Q.weight", "Q.bias", "W.weight", "W.bias"))
        a.G1w, a.G1b, a.G2w = m.G1.weight.data_ptr(), m.G1.bias.data_ptr(), m.G2.weight.data_ptr()
        a.gG1w, a.gG1b, a.gG2w = grads["G1.weight"].data_ptr(), grads["G1.bias"].data_ptr(), grads["G2.weight"].data_ptr()
        a.triples, a.B = prep.triples.data_ptr(), prep.triples.shape[0]
        a.dup_counts = prep.counts.data_ptr() if reference_compat else None
        a.margin, a.feat_margin = float(margin), 0.0001
        a.flat_grad, a.n_params = self.flat_grad.data_ptr(), self.flat_grad.numel()
        need = nat.lib().ps_train_step_workspace(a)
        if need < 0:
            raise nat.NativeError(f"ps_train_step_workspace: {nat.lib().ps_last_error().decode()}")
        # One workspace per engine, grown in 64 MB steps and reused by every later step: frontier sizes change from step to
        # step, and a fresh torch.empty of a new size every step sends the caching allocator to cudaMalloc every so
        # often (a device synchronisation: single 25-180 ms steps in an otherwise 6.6 ms loop).  Steps run in stream
        # order on one stream, so reuse is safe.
        ws = getattr(self, "_step_ws", None)
        if ws is None or ws.numel() < need + 256 or ws.device != feats.device:
            ws = self._step_ws = torch.empty(((int(need) + 256 + (1 << 26) - 1) >> 26) << 26, dtype=torch.uint8, device="cuda")
        base = (ws.data_ptr() + 255) & ~255
        a.workspace, a.workspace_bytes = base, int(need)
        out = torch.empty(3 if diagnostics else 1, dtype=torch.float32, device="cuda")  # [loss, feature loss, variance]
        a.loss_out = out.data_ptr()
        if diagnostics:
            a.batch, a.diag_out = prep.batch.data_ptr(), out.data_ptr() + 4
        emb_ptr = nat.c_void_p()
        a.emb_out = nat.ctypes.pointer(emb_ptr)
        # data parallel (ps_dist.GradSync): an event the step records once only layer 0's Q gradients are still being computed
        self.upper_grads_event, self.upper_grads_offset = None, 0
        if getattr(self, "want_upper_grads_event", False) and L >= 2:
            ev = getattr(self, "_upper_ev", None)
            if ev is None:
                ev = self._upper_ev = torch.cuda.Event()
                ev.record()  # creates the underlying cudaEvent_t
            a.upper_grads_event = ev.cuda_event
            self.upper_grads_event = ev
            q0 = m.conv_layers[0].Q  # flat order = named_parameters(): conv_layers.0.Q.weight, Q.bias come first
            self.upper_grads_offset = q0.weight.numel() + q0.bias.numel()
        nat.train_step(a, launches=10 + 13 * L)
        n_top = prep.plan.layers[-1].n
        off = emb_ptr.value - ws.data_ptr()
        emb = ws[off: off + n_top * do * 4].view(torch.float32).view(n_top, do).clone()  # the workspace is reused by the next step
        self.last_diag = out[1:] if diagnostics else None
        return out[:1], emb, prep.triples

    @torch.no_grad()
    def embed(self, feats: torch.Tensor, nodes: torch.Tensor) -> torch.Tensor:
        """Inference: embeddings of `nodes` (int64 on device, duplicates allowed)."""
        m = self.model
        top, inv = torch.unique(nodes, return_inverse=True)
        plan = build_plan(top, m.n_layers, m.T, NeighborTable.of(m.nbhds), need_backward=False)
        out, _ = self.forward(feats, plan, keep=False)
        return out[inv]


    @torch.no_grad()
    def embed_range(self, feats: torch.Tensor, lo: int, hi: int, chunk: int = 1 << 18, stats: Optional[dict] = None,
                    gather_layer=None) -> torch.Tensor:
        """Full-graph inference for the node range [lo, hi): embeddings float32 [hi-lo, out_dim] on the device.
        Result-identical to embed(arange(lo, hi)) (per row the same kernels do the same arithmetic) but organised
        layer by layer over the range's T-hop closure instead of one frontier plan: a dense bool mask per layer marks
        the nodes whose layer output is needed, Q is applied ONCE per needed input row of a layer, and targets are
        aggregated in chunks against node-indexed activation tables.  This is the shard a rank owns in node-range
        sharded inference (BASELINE.json configs[3]); nothing is exchanged between ranks.

        gather_layer (optional): a callable(table [N, out_dim], lo, hi) that fills the rows outside [lo, hi) of a
        layer's node-indexed output table from the other ranks (an all-gather).  With it every rank computes every
        layer for ITS range only (no closure, 1/world of the work) and the result is still identical: the closure
        rows it would have recomputed arrive from their owners."""
        m = self.model
        table = NeighborTable.of(m.nbhds)
        if isinstance(table, OnlineNeighbors):
            table = table.materialize(m.T)
        N, T, L = table.n, m.T, m.n_layers
        in_dims, dh, do = self._dims()
        dev = feats.device
        if not (0 <= lo <= hi <= N):
            raise IndexError("node range out of bounds")
        if hi == lo:
            return torch.empty((0, do), dtype=torch.float32, device=dev)
        if gather_layer is not None and do <= 128 and do < dh:
            return self._embed_range_exchange(feats, lo, hi, chunk, stats, gather_layer, table)

        def mark_neighbours(mask, idx):
            for i in range(0, idx.numel(), chunk):
                mask[table.nodes[idx[i:i + chunk], :T].reshape(-1).long()] = True

        need = [None] * L
        top = torch.zeros(N, dtype=torch.bool, device=dev)
        top[lo:hi] = True
        need[L - 1] = top
        for l in range(L - 1, 0, -1):
            if gather_layer is not None or (lo == 0 and hi == N):
                need[l - 1] = top  # owners compute and the exchange delivers the rest / the shard is the whole graph
                continue
            nxt = need[l].clone()
            mark_neighbours(nxt, need[l].nonzero().squeeze(1))
            need[l - 1] = nxt
        h_prev = feats
        full_T = T == table.Tp  # table rows are exactly the T neighbours: row slices are views
        for l in range(L):
            conv = m.conv_layers[l]
            din = in_dims[l]
            # targets: a contiguous range when the mask is the shard itself (top layer; every layer with an exchange or on one
            # GPU), else the set bits of the closure mask
            if need[l] is top:
                t_lo, t_hi, targets = lo, hi, None
                n_t = hi - lo
            else:
                targets = need[l].nonzero().squeeze(1)
                n_t = targets.numel()
            # rows that get the Q transform: when the targets reference (nearly) every row anyway -- 3+ references per
            # row on average -- all N rows are transformed in node order: no mask, no position map, no gather in the GEMM
            dense_z = n_t * T >= 3 * N
            if dense_z:
                zrows, zpos, nz = None, None, N
            else:
                zmask = torch.zeros(N, dtype=torch.bool, device=dev)
                mark_neighbours(zmask, targets if targets is not None else torch.arange(t_lo, t_hi, device=dev))
                zrows = zmask.nonzero().squeeze(1).to(torch.int32)
                zpos = (torch.cumsum(zmask, 0, dtype=torch.int32) - 1)
                del zmask
                nz = zrows.numel()
            projected = do <= 128 and do < dh  # aggregate do-wide projections instead of dh-wide activations (_conv_projected)
            if projected:
                z = self._projected_rows(h_prev, din, nz, zrows, conv, l)
                zw, w_cat = do, self._projected_weight(conv, din)
            else:
                z = torch.empty((nz, dh), dtype=torch.float32, device=dev)
                nat.gemm(h_prev, conv.Q.weight, z, nz, dh, din, p_rows=zrows, bias=conv.Q.bias, act=1, tag=f"gemm_q_fwd_l{l}")
                zw, w_cat = dh, conv.W.weight
            last = l == L - 1
            h = torch.empty((hi - lo, do) if last else (N, do), dtype=torch.float32, device=dev)
            for i in range(0, n_t, chunk):
                n = min(chunk, n_t - i)
                if targets is None:  # rows t_lo + i .. : slices instead of gathers / scatters
                    r0 = t_lo + i
                    self_rows = torch.arange(r0, r0 + n, dtype=torch.int32, device=dev)
                    nb_ids = table.nodes[r0:r0 + n, :T]
                    w = table.w[r0:r0 + n, :T]
                    out = h[r0 - lo: r0 - lo + n] if last else h[r0:r0 + n]
                else:
                    c = targets[i:i + n]
                    self_rows = c.to(torch.int32)
                    nb_ids = table.nodes[c, :T]
                    w = table.w[c, :T]
                    out = torch.empty((n, do), dtype=torch.float32, device=dev)
                if not full_T or targets is not None:
                    nb_ids, w = nb_ids.contiguous(), w.contiguous()
                nbz = nb_ids if zpos is None else zpos[nb_ids.reshape(-1).long()].view(n, T)
                cat = torch.empty((n, din + zw), dtype=torch.float32, device=dev)
                inv = torch.empty((n,), dtype=torch.float32, device=dev)
                nat.aggregate_fwd(h_prev, self_rows, din, z, nbz, w, zw, cat, inv, tag=f"aggregate_fwd_l{l}")
                if do <= 128:
                    nat.gemm(cat, w_cat, out, n, do, din + zw, bias=conv.W.bias, act=1, l2norm=True, tag=f"gemm_w_fwd_l{l}")
                else:
                    nat.gemm(cat, w_cat, out, n, do, din + zw, bias=conv.W.bias, act=1)
                    nat.l2norm_rows(out, torch.empty((n,), dtype=torch.float32, device=dev))
                if targets is not None:
                    if last:
                        h[c - lo] = out
                    else:
                        h[c] = out
            if stats is not None:
                stats[f"layer{l}"] = {"targets": int(n_t), "z_rows": int(nz)}
            del z, zpos, zrows
            if gather_layer is not None and not last:
                gather_layer(h, lo, hi)
            h_prev = h
        n_top = hi - lo
        a1 = torch.empty((n_top, do), dtype=torch.float32, device=dev)
        nat.gemm(h_prev, m.G1.weight, a1, n_top, do, do, bias=m.G1.bias, act=1)
        out = torch.empty((n_top, do), dtype=torch.float32, device=dev)
        nat.gemm(a1, m.G2.weight, out, n_top, do, do)
        return out


def _embed_range_exchange(self, feats, lo, hi, chunk, stats, gather_layer, table):
    """embed_range with an exchange step, everything sharded: per layer a rank transforms and projects ITS rows only
    (zp[lo:hi] = leaky(Q x + b) W2^T), the projected table [N, out_dim] is all-gathered (gather_layer), and the rank
    aggregates its own targets against it.  The next layer needs the layer output for the rank's own rows only, so the
    projection is the one thing exchanged: L all-gathers of [N, out_dim] fp32, 1 / world of every kernel's work per rank.
    Row by row the arithmetic is that of the other paths (same kernels, same K order): results are identical."""
    m = self.model
    N, T, L = table.n, m.T, m.n_layers
    in_dims, dh, do = self._dims()
    dev = feats.device
    n_own = hi - lo
    full_T = T == table.Tp
    h_own = None
    for l in range(L):
        conv, din = m.conv_layers[l], in_dims[l]
        src = feats[lo:hi] if l == 0 else h_own
        zp = torch.empty((N, do), dtype=torch.float32, device=dev)
        self._projected_rows(src, din, n_own, None, conv, l, out=zp[lo:hi])
        gather_layer(zp, lo, hi)
        w_cat = self._projected_weight(conv, din)
        h_new = torch.empty((n_own, do), dtype=torch.float32, device=dev)
        hin = feats if l == 0 else h_own      # layer 0 reads self rows by node id, deeper layers by position in the shard
        base = lo if l == 0 else 0
        for i in range(0, n_own, chunk):
            n = min(chunk, n_own - i)
            self_rows = torch.arange(base + i, base + i + n, dtype=torch.int32, device=dev)
            nbz, w = table.nodes[lo + i: lo + i + n, :T], table.w[lo + i: lo + i + n, :T]
            if not full_T:
                nbz, w = nbz.contiguous(), w.contiguous()
            cat = torch.empty((n, din + do), dtype=torch.float32, device=dev)
            inv = torch.empty((n,), dtype=torch.float32, device=dev)
            nat.aggregate_fwd(hin, self_rows, din, zp, nbz, w, do, cat, inv, tag=f"aggregate_fwd_l{l}")
            nat.gemm(cat, w_cat, h_new[i:i + n], n, do, din + do, bias=conv.W.bias, act=1, l2norm=True, tag=f"gemm_w_fwd_l{l}")
        if stats is not None:
            stats[f"layer{l}"] = {"targets": int(n_own), "z_rows": int(n_own)}
        del zp
        h_own = h_new
    a1 = torch.empty((n_own, do), dtype=torch.float32, device=dev)
    nat.gemm(h_own, m.G1.weight, a1, n_own, do, do, bias=m.G1.bias, act=1)
    out = torch.empty((n_own, do), dtype=torch.float32, device=dev)
    nat.gemm(a1, m.G2.weight, out, n_own, do, do)
    return out


Engine._embed_range_exchange = _embed_range_exchange


class PinSageFunction(torch.autograd.Function):
    """Autograd bridge so that `model(features, nodeset)` composes with any torch loss, as
    the reference's nn.Module does.  The backward applies the reference's duplicate-node
    factor (pinsage_model.py:260,265): a node listed k times in `nodeset` receives k times
    the sum of its rows' gradients."""

    @staticmethod
    def forward(ctx, engine: Engine, feats, nodeset, reference_compat, *params):
        top, inv = torch.unique(nodeset, return_inverse=True)
        m = engine.model
        need = any(ctx.needs_input_grad[4:])  # (grad mode is off inside Function.forward)
        plan = build_plan(top, m.n_layers, m.T, NeighborTable.of(m.nbhds), need_backward=need)
        out, saved = engine.forward(feats, plan, keep=need)
        ctx.engine, ctx.saved, ctx.inv, ctx.n_top, ctx.compat = engine, saved, inv, top.numel(), reference_compat
        ctx.names = [n for n, _ in m.named_parameters()]
        return out[inv]

    @staticmethod
    def backward(ctx, g):
        if ctx.saved is None:
            raise RuntimeError("backward called twice or without saved activations")
        engine = ctx.engine
        d_out = torch.zeros((ctx.n_top, g.shape[1]), dtype=torch.float32, device=g.device)
        d_out.index_add_(0, ctx.inv, g.contiguous().to(torch.float32))
        if ctx.compat:
            d_out *= torch.bincount(ctx.inv, minlength=ctx.n_top).to(torch.float32)[:, None]
        grads = {n: torch.zeros_like(p) for n, p in engine.model.named_parameters()}
        engine.backward(ctx.saved, d_out, grads)
        ctx.saved = None
        return (None, None, None, None, *[grads[n] for n in ctx.names])
