"""Cosine k-nearest-neighbour search over an embedding table on the device: the N^2 * d part
of the reference's evaluation (`cosine_sim_ab` / `knn_from_emb`, baselines.py:69-103).
The similarity GEMM runs on tcgen05 (sm_100a).  Default path: ps_gemm_filter, whose epilogue keeps only the
candidates above a per-query threshold (no [queries, N] tile is ever written), then an exact top-(k+1) of the short
candidate lists.  Fallback / small problems: ps_gemm tile + ps_topk_rows (radix select, 4 streaming reads of the tile)."""
from __future__ import annotations

import torch

import ps_native as nat


def cosine_sim_ab(a, b, eps=1e-16):
    """sim[i, j] = a_i . b_j / (|a_i| |b_j| + eps)   (baselines.py:69-77)."""
    out_cpu = not a.is_cuda
    a = a.to("cuda", torch.float32).contiguous()
    b = b.to("cuda", torch.float32).contiguous()
    sim = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device="cuda")
    nat.gemm(a, b, sim, a.shape[0], b.shape[0], a.shape[1], tag="gemm_knn")
    sim /= a.norm(dim=1)[:, None] * b.norm(dim=1)[None, :] + eps
    return sim.cpu() if out_cpu else sim


# ---- fused search: no [queries, N] similarity tile -----------------------------------------------------------------
_SAMPLE_RANK = 128     # order statistic of the sample that becomes a query's threshold
_OVERSHOOT = 2.2       # expected candidates per query = _OVERSHOOT * (k + 1); relative spread ~ 1 / sqrt(_SAMPLE_RANK)
fused_stats = {"tiles": 0, "fallback_tiles": 0}


def _knn_tile_unfused(e, qe, k1):
    """[q, N] similarity tile through ps_gemm, then ps_topk_rows (4 streaming reads of the tile)."""
    n = e.shape[0]
    old_pack = nat.gemm_tc_pack(0)  # the all-rows operand is far larger than the query tile: stream it, do not pack it
    try:
        sim = torch.empty((qe.shape[0], n), dtype=torch.float32, device="cuda")
        nat.gemm(qe, e, sim, qe.shape[0], n, e.shape[1], tag="gemm_knn")
    finally:
        nat.gemm_tc_pack(old_pack)
    return nat.topk_rows(sim, k1) if k1 <= min(8192, n) else sim.topk(k1, dim=1, largest=True)


def _knn_tile_fused(e, qe, k1):
    """Top-k1 of every query of the tile without materialising the similarities:
      1. thresholds: similarities against a strided SAMPLE of the table (a view with a larger leading dimension,
         ~ _SAMPLE_RANK / (_OVERSHOOT k1) of the rows), per query the _SAMPLE_RANK-th largest of them;
      2. ps_gemm_filter: the full similarity GEMM whose epilogue appends (sim, row) to the query's candidate list when
         sim >= threshold (expected _OVERSHOOT * k1 candidates, spread ~9 %);
      3. exact top-k1 of every list (ps_topk_rows_mapped; ties by ascending row id, like the unfused path).
    Returns None when a list came out shorter than k1 or overflowed (the caller redoes the tile unfused): the result
    is exact whenever it is returned."""
    n, d = e.shape
    nq = qe.shape[0]
    stride = max(1, int(_OVERSHOOT * k1 / _SAMPLE_RANK))  # E[candidates] = _SAMPLE_RANK * stride ~ _OVERSHOOT * k1
    n_s = (n + stride - 1) // stride
    if stride < 4 or n_s < 4 * _SAMPLE_RANK or nq % 4 or nq < 64 or d % 4 or n < 1024:
        return None  # too little to gain (small table or large k): the sample would be most of the table
    sample = e[::stride]                                   # [n_s, d] view, leading dimension stride * d
    n_s4 = n_s - n_s % 4
    sim_s = torch.empty((nq, n_s4), dtype=torch.float32, device="cuda")
    old_pack = nat.gemm_tc_pack(0)
    try:
        nat.gemm(qe, sample, sim_s, nq, n_s4, d, tag="gemm_knn_sample")
    finally:
        nat.gemm_tc_pack(old_pack)
    thr = nat.topk_rows(sim_s, _SAMPLE_RANK)[0][:, -1].contiguous()
    cap = int(2 * _OVERSHOOT * k1) + 256
    cap += -cap % 4
    cnt, val, row = nat.gemm_filter(e, qe, thr, cap)
    lo, hi = (int(v) for v in torch.stack([cnt.min(), cnt.max()]).tolist())  # one host read per tile
    fused_stats["tiles"] += 1
    if lo < k1 or hi > cap:
        fused_stats["fallback_tiles"] += 1
        return None
    return nat.topk_rows_mapped(val, row, cnt, k1)


def knn_from_emb(emb, q, k, sim_func=None, q_tile=1024, fused=True):
    """(weights [len(q), k], nodes [len(q), k]): the k most cosine-similar rows of `emb` for
    every query row index in q; like the reference the top-(k+1) is taken and column 0
    (assumed to be the query itself) dropped (baselines.py:91-103).  Embedding dims are
    zero-padded to a multiple of 4 for the 128-bit loads.
    Rows are L2-normalised once, so a similarity tile is ONE GEMM (queries x all rows); the reference's
    `dot / (|a||b| + 1e-16)` differs from it by 1e-16 relative.  fused=True (default) never writes the
    [q_tile, N] tile: the GEMM's epilogue keeps only the candidates above a per-query threshold (_knn_tile_fused);
    small tables, huge k and the rare tile whose threshold missed go through the tile + ps_topk_rows path."""
    out_cpu = not emb.is_cuda
    e = emb.to("cuda", torch.float32)
    if e.shape[1] % 4:
        e = torch.nn.functional.pad(e, (0, 4 - e.shape[1] % 4))
    e = (e / e.norm(dim=1, keepdim=True).clamp_min(1e-30)).contiguous()
    q = torch.as_tensor(q).to("cuda", torch.int64)
    n, d = e.shape
    ws, ns = [], []
    for i in range(0, q.numel(), q_tile):
        qe = e[q[i:i + q_tile]].contiguous()
        res = None
        if fused and k + 1 <= min(4096, n):
            pad = -qe.shape[0] % 4  # the packed operand wants a multiple of 4 query rows: pad with copies of the last one
            qp = torch.cat([qe, qe[-1:].expand(pad, d)]) if pad else qe
            res = _knn_tile_fused(e, qp, k + 1)
            if res is not None and pad:
                res = (res[0][:-pad], res[1][:-pad])
        if res is None:
            res = _knn_tile_unfused(e, qe, k + 1)
        w, nb = res
        ws.append(w[:, 1:]); ns.append(nb[:, 1:])
    w, nb = torch.cat(ws, 0), torch.cat(ns, 0)
    return (w.cpu(), nb.cpu()) if out_cpu else (w, nb)
