"""Next-song evaluation metrics with the reference's definitions (eval.py:227-250),
vectorised (the reference loops over test pairs in Python).  Device-agnostic torch."""
from __future__ import annotations

import torch


def hit_rate(knn_mat, test_positives, K):
    """Fraction of test pairs (q, pos) with pos among the first K neighbours of q
    (eval.py:227-238)."""
    knn = torch.as_tensor(knn_mat)[:, :K]
    tp = torch.as_tensor(test_positives).to(knn.device)
    hits = (knn[tp[:, 0]] == tp[:, 1:2]).any(1)
    return float(hits.sum()) / tp.shape[0]


def mrr(knn_mat, test_positives, K, scaling=1):
    """Mean of scaling / rank (1-based); rank = K when the positive is absent from the first
    K neighbours (eval.py:240-250)."""
    knn = torch.as_tensor(knn_mat)[:, :K]
    tp = torch.as_tensor(test_positives).to(knn.device)
    eq = knn[tp[:, 0]] == tp[:, 1:2]
    first = torch.argmax(eq.to(torch.int8), dim=1) + 1
    rank = torch.where(eq.any(1), first, torch.full_like(first, K)).to(torch.float64)
    return float((1.0 / (rank / scaling)).sum() / tp.shape[0])
