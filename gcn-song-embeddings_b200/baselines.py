"""PinSage-side slice of the reference's `baselines` module (reference:
/root/reference/baselines.py:33-103, 281-377): the recommender ABCs, cosine kNN from
embeddings (on the device, ps_knn), the per-track embedding loader, the PinSage wrapper and the
personalised-PageRank baseline (it is the same walker).
The other competing recommenders (Jaccard / node2vec / implicit CF / GraphSAGE) are out of
scope."""
from __future__ import annotations

import os
import time
from abc import ABC, abstractmethod

import torch
from tqdm import tqdm

import pinsage_model as psm
from pinsage_training import PinSage
from ps_knn import cosine_sim_ab, knn_from_emb  # noqa: F401  (re-exported reference names)


class PredictionModel(ABC):
    """Base recommender class (baselines.py:33-46)."""

    @abstractmethod
    def __init__(self):
        pass

    @abstractmethod
    def train(self, g, ids, train_set, test_set, features):
        pass

    @abstractmethod
    def knn(self, nodeset, k):
        pass


class EmbeddingModel(PredictionModel):
    """An embedding-based recommender (baselines.py:48-53)."""

    @abstractmethod
    def embed(self, nodeset):
        pass


class PersPageRank(PredictionModel):
    """Nearest graph neighbours by personalised PageRank, i.e. random walks with restarts
    (baselines.py:107-151).  The reference runs a copy of the PinSage walker in Python (n_hops 1000, alpha 0.85)
    and takes topk of the dense visit-probability row; here knn() is one ps_walk_topt launch (walk + visit counts +
    top-k fused on the device).  Order inside equal counts is (node id ascending); slots beyond the number of
    visited nodes carry weight 0."""

    def __init__(self):
        self.n_hops = 1000
        self.alpha = 0.85

    def visit_prob(self, g, nodeset, n_hops, alpha):
        """Dense float64 [len(nodeset), number_of_nodes] visit probabilities, self entry zeroed (baselines.py:114-143)."""
        return psm.sample_neighborhood(g, _n_items(g), nodeset, n_hops, alpha)

    def train(self, g, ids, train_set, test_set, features):
        self.g = g
        self.n_items = len(ids)

    def knn(self, nodeset, k):
        return psm.sample_neighborhood_topt(self.g, self.n_items, nodeset, self.n_hops, self.alpha, k)


def _n_items(g):
    return g.n_tracks if hasattr(g, "n_tracks") else g.number_of_nodes()


def _load_embeddings(ids, load_dir):
    """The embeddings of `ids` as one matrix (baselines.py:281-294): the single-tensor copy beside `load_dir` when it
    is there and lists exactly these ids (pinsage_training.save_embedding_matrix), else the per-track `<id>.pt` files."""
    from pinsage_training import load_embedding_matrix
    emb = load_embedding_matrix(load_dir, ids)
    if emb is not None:
        return emb
    emb_list = [torch.load(os.path.join(load_dir, track_id + ".pt")) for track_id in tqdm(ids, desc="Loading embeddings")]
    return torch.stack(emb_list, dim=0)


class EmbLoader(EmbeddingModel):
    """Load precomputed embeddings as a recommender method (baselines.py:297-328)."""

    def __init__(self, load_dir):
        self.load_dir = load_dir
        self.embedding = None

    def train(self, g, ids, train_set, test_set, features):
        print(f"Loading embeddings from {self.load_dir} ...")
        self.embedding = _load_embeddings(ids, self.load_dir)

    def embed(self, nodeset):
        return self.embedding[nodeset, :]

    def knn(self, nodeset, k):
        return knn_from_emb(self.embedding, nodeset, k)


class PinSageWrapper(EmbeddingModel):
    """Train PinSage behind the baseline interface (baselines.py:331-377).  train_params are
    applied as attributes after construction, as the reference does."""

    def __init__(self, train_params=None, run_name=None, log=True):
        self.embedding = None
        self.train_params = train_params if train_params else {}
        self.run_name = run_name if run_name else time.strftime("%X %x")
        self.log = log

    def train(self, g, ids, train_set, test_set, features):
        print("Training PinSage with parameters:")
        print(self.train_params)
        self.trainer = PinSage(g, len(ids), features, train_set, log=self.log, load_save=False)
        for param, value in self.train_params.items():
            setattr(self.trainer, param, value)
        self.trainer.run_name = self.run_name
        self.trainer.train()
        print("Embedding...")
        self.embedding = self.trainer.embed(torch.arange(len(ids)), bsize=None)

    def embed(self, nodeset):
        return self.embedding[nodeset, :]

    def knn(self, nodeset, k):
        return knn_from_emb(self.embedding, nodeset, k)
