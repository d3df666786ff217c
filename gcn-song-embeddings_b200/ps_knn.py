"""Cosine k-nearest-neighbour search over an embedding table on the device: the N^2 * d part
of the reference's evaluation (`cosine_sim_ab` / `knn_from_emb`, baselines.py:69-103).
The similarity tiles are ps_gemm calls (tcgen05 on sm_100a); the top-(k+1) selection per query tile is
ps_topk_rows (radix select + in-shared-memory sort, 4 streaming reads of the tile), which replaces torch.topk."""
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


def knn_from_emb(emb, q, k, sim_func=None, q_tile=1024):
    """(weights [len(q), k], nodes [len(q), k]): the k most cosine-similar rows of `emb` for
    every query row index in q; like the reference the top-(k+1) is taken and column 0
    (assumed to be the query itself) dropped (baselines.py:91-103).  Embedding dims are
    zero-padded to a multiple of 4 for the 128-bit loads.
    Rows are L2-normalised once, so a similarity tile is ONE GEMM (queries x all rows) with no pass over the
    [q_tile, N] tile afterwards; the reference's `dot / (|a||b| + 1e-16)` differs from it by 1e-16 relative."""
    out_cpu = not emb.is_cuda
    e = emb.to("cuda", torch.float32)
    if e.shape[1] % 4:
        e = torch.nn.functional.pad(e, (0, 4 - e.shape[1] % 4))
    e = (e / e.norm(dim=1, keepdim=True).clamp_min(1e-30)).contiguous()
    q = torch.as_tensor(q).to("cuda", torch.int64)
    n, d = e.shape
    ws, ns = [], []
    # the all-rows operand changes per call and is far larger than the query tile: stream it through the producer
    # warps instead of packing it as a "weight" image
    old_pack = nat.gemm_tc_pack(0)
    try:
        for i in range(0, q.numel(), q_tile):
            qe = e[q[i:i + q_tile]].contiguous()
            sim = torch.empty((qe.shape[0], n), dtype=torch.float32, device="cuda")
            nat.gemm(qe, e, sim, qe.shape[0], n, d, tag="gemm_knn")
            w, nb = nat.topk_rows(sim, k + 1) if k + 1 <= min(8192, n) else sim.topk(k + 1, dim=1, largest=True)
            ws.append(w[:, 1:]); ns.append(nb[:, 1:])
            del sim
    finally:
        nat.gemm_tc_pack(old_pack)
    w, nb = torch.cat(ws, 0), torch.cat(ns, 0)
    return (w.cpu(), nb.cpu()) if out_cpu else (w, nb)
