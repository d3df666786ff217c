"""Next-song evaluation slice of the reference's `eval` module (reference:
/root/reference/eval.py:31, 52-175, 227-250, 376-443): kNN-list precompute / cache, hit-rate,
MRR, the accuracy columns of the results table.  Beyond-accuracy metrics, the interactive
crawler and the LaTeX export are out of scope."""
from __future__ import annotations

import os
import time

import pandas as pd
import torch
from tqdm import tqdm

from baselines import EmbeddingModel
from ps_eval import hit_rate, mrr  # noqa: F401  (reference names; vectorised)

PRECOMP_K = 1000


def save_embedding(model, model_name, ids, save_dir):
    """eval.py:74-90."""
    save_dir = os.path.join(save_dir, "emb", model_name)
    os.makedirs(save_dir, exist_ok=True)
    emb_time = 0
    if len(os.listdir(save_dir)) == 0:
        t0 = time.time()
        emb = model.embed(torch.arange(0, len(ids), dtype=torch.int64))
        emb_time = time.time() - t0
        from pinsage_training import save_embedding_matrix
        save_embedding_matrix(save_dir, ids, emb)  # one tensor beside the per-track files (fast path of EmbLoader)
        for i in range(emb.shape[0]):
            torch.save(emb[i, :].clone().detach().cpu(), os.path.join(save_dir, ids[i] + ".pt"))
    return emb_time


def save_knn(model, model_name, ids, save_dir, train_time=0, emb_time=0, k=None):
    """kNN lists of all nodes in 1000-query batches, cached as the reference's 5-tuple
    (knn_w, knn_n, train_time, emb_time, knn_time) (eval.py:112-143)."""
    k = PRECOMP_K if k is None else k
    save_dir = os.path.join(save_dir, "knn")
    os.makedirs(save_dir, exist_ok=True)
    save_path = os.path.join(save_dir, model_name + ".pt")
    all_nodes = torch.arange(0, len(ids), dtype=torch.int64)
    if not os.path.isfile(save_path):
        knn_time, ws, ns = 0, [], []
        for i in tqdm(range(0, len(all_nodes), 1000), desc="Computing KNN"):
            t0 = time.time()
            w, n = model.knn(all_nodes[i:i + 1000], k)
            ws.append(w.cpu()); ns.append(n.cpu())
            knn_time += time.time() - t0
        torch.save((torch.cat(ws, 0), torch.cat(ns, 0), train_time, emb_time, knn_time), save_path)


def load_knn(model_name, ids, save_dir):
    return torch.load(os.path.join(save_dir, "knn", model_name + ".pt"))


def precompute_model(model, model_name, g, ids, train_pos, test_pos, features, save_dir, k=None):
    """Train, embed and build the kNN cache of one model unless cached (eval.py:52-71)."""
    if os.path.isfile(os.path.join(save_dir, "knn", model_name + ".pt")):
        return
    t0 = time.time()
    model.train(g, ids, train_pos, test_pos, features)
    train_time = time.time() - t0
    emb_time = save_embedding(model, model_name, ids, save_dir) if isinstance(model, EmbeddingModel) else 0
    save_knn(model, model_name, ids, save_dir, train_time=train_time, emb_time=emb_time, k=k)


class KnnDict(dict):
    """{model_name: (knn_weights, knn_indices)} with the timings on the side (the reference's
    LazyKnnDict, eval.py:177-202, without the laziness)."""

    def __init__(self):
        super().__init__()
        self.times = {}

    def get_times(self, model):
        return self.times[model]


def get_knn_dict(models, g, ids, train_pos, test_pos, features, save_dir, k=None):
    """eval.py:166-175."""
    out = KnnDict()
    for name, model in models.items():
        precompute_model(model, name, g, ids, train_pos, test_pos, features, save_dir, k=k)
        knn_w, knn_n, train_t, emb_t, knn_t = load_knn(name, ids, save_dir)
        out[name] = (knn_w, knn_n)
        out.times[name] = (train_t, emb_t, knn_t)
    return out


def low_degree_accuracy(knn_mat, g, test_positives, K, degree_thr=1, acc_func=mrr):
    """Accuracy restricted to test pairs whose query has at most degree_thr collections; 0 when no node
    qualifies (eval.py:376-389)."""
    deg = torch.as_tensor(g.in_degrees())[: knn_mat.shape[0]]
    if int((deg <= degree_thr).sum()) == 0:
        return 0
    tp = torch.as_tensor(test_positives)
    return acc_func(knn_mat, tp[deg[tp[:, 0]] <= degree_thr], K)


def low_co_accuracy(knn_mat, g, test_positives, K, co_thr=1, acc_func=mrr):
    """Accuracy restricted to test pairs whose query occurs in at most co_thr test pairs (eval.py:391-406: the row
    sums of the track-track co-occurrence count matrix of the test positives are the per-query pair counts)."""
    tp = torch.as_tensor(test_positives)
    co_counts = torch.bincount(tp[:, 0], minlength=knn_mat.shape[0])
    return acc_func(knn_mat, tp[co_counts[tp[:, 0]] <= co_thr], K)


def compute_results_table(knn_dict, test_positives, g, times=True, degree_thr=1):
    """hit-rate@10/100/500, MRR@1000, low-degree and low-co-occurrence MRR per model (eval.py:413-443)."""
    results = {}
    for model in knn_dict:
        _, knn_mat = knn_dict[model]
        row = {f"hr (k={k})": hit_rate(knn_mat, test_positives, k) for k in (10, 100, 500)}
        row["mrr"] = mrr(knn_mat, test_positives, 1000, 1)
        row["low-degree accuracy"] = low_degree_accuracy(knn_mat, g, test_positives, 1000, degree_thr=degree_thr, acc_func=mrr)
        row["low-co accuracy"] = low_co_accuracy(knn_mat, g, test_positives, 1000, co_thr=1, acc_func=mrr)
        if times and hasattr(knn_dict, "get_times"):
            row["t (train)"], row["t (emb)"], row["t (knn)"] = knn_dict.get_times(model)
        results[model] = row
    return pd.DataFrame.from_dict(results, orient="index")
