"""Cosine k-nearest-neighbour search over an embedding table on the device: the N^2 * d part
of the reference's evaluation (`cosine_sim_ab` / `knn_from_emb`, baselines.py:69-103).
The similarity tiles are ps_gemm calls (tcgen05 on sm_100a); the top-(k+1) selection per
query tile is torch.topk for now (a fused running-top-k epilogue is the next step,
SURVEY.md section 8f item 1)."""
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
    zero-padded to a multiple of 4 for the 128-bit loads."""
    out_cpu = not emb.is_cuda
    e = emb.to("cuda", torch.float32)
    if e.shape[1] % 4:
        e = torch.nn.functional.pad(e, (0, 4 - e.shape[1] % 4))
    e = e.contiguous()
    q = torch.as_tensor(q).to("cuda", torch.int64)
    ws, ns = [], []
    for i in range(0, q.numel(), q_tile):
        qe = e[q[i:i + q_tile]].contiguous()
        sim = cosine_sim_ab(qe, e)
        w, n = sim.topk(k + 1, dim=1, largest=True)
        ws.append(w[:, 1:]); ns.append(n[:, 1:])
    w, n = torch.cat(ws, 0), torch.cat(ns, 0)
    return (w.cpu(), n.cpu()) if out_cpu else (w, n)
