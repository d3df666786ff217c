"""B200 drop-in for the reference's `spotify_graph` module: same dataset files, same node
numbering, same feature standardisation and positives split
(reference: /root/reference/spotify_graph.py:15-110); the graph object is a CSR-backed
PSGraph instead of a DGLGraph (the engine only needs adjacency, `nbhds_path`, `base_dir`).
Dataset statistics / plotting helpers of the reference are out of scope.

Files (schema of dataset_creation/get_data.py:102,211-214):
  tracks.json       {track_id: {...}}          node order = key order, tracks first
  collections.json  {collection_id: {...}}
  graph.json        {"tracks": [...], "collections": [...], "edges": [{"from": id, "to": id}, ...]}
                    (both directions listed, duplicates kept)
  <features_dir>/<track_id>.pt   1-D float tensor per track
  positives*.json   [{"a": track_id, "b": track_id}, ...]
"""
from __future__ import annotations

import json
import os
from os import path

import numpy as np
import torch

import ps_native
from ps_graph import PSGraph


def standardize_features(features):
    """(x - mean) / (std_unbiased + 1e-12) per column (spotify_graph.py:77-79).  With a CUDA device the statistics
    and the division run in HBM (ps_standardize, csrc/ingest.cu: fp64 column sums, two passes) and the standardised
    table comes back in the caller's placement (the trainer uploads it once anyway); without one, the same
    formula in framework ops (host-only tooling and the CPU tests)."""
    if torch.cuda.is_available() and features.dim() == 2 and features.shape[0] > 1:
        x = features.to("cuda", torch.float32, copy=True).contiguous()
        ps_native.standardize_(x, 1e-12)
        return x if features.is_cuda else x.cpu()
    mean = features.mean(dim=0)
    std = features.std(dim=0, unbiased=True) + 1e-12
    return (features - mean) / std


def _read_json(file_path):
    with open(file_path, "r", encoding="utf-8") as f:
        return json.load(f)


class SpotifyGraph():
    """Reads a dataset directory written by the reference's crawler; attribute names are the reference's
    (eval.py / dashboard.py read `tracks`, `collections`, `graph`, `base_dir`, `nbhds_path`, `features`, ...)."""

    def __init__(self, dir, features_dir):
        self.base_dir = dir
        join = lambda name: path.join(dir, name)
        self.nbhds_path = join("neighborhoods.pt")
        self.tracks_pth, self.col_pth, self.graph_pth = join("tracks.json"), join("collections.json"), join("graph.json")
        self.img_dir, self.clip_dir = join("images"), join("clips")
        print("Loading graph...")
        self.tracks = _read_json(self.tracks_pth)
        self.collections = _read_json(self.col_pth)
        self.graph = _read_json(self.graph_pth)
        self.ft_dir = features_dir if features_dir is not None and os.path.isdir(features_dir) else None
        self.features_dict = {}
        self._node_index = None

    def _index(self):
        """id -> node number: position in list(tracks) + list(collections) (spotify_graph.py:43-46,58)."""
        if self._node_index is None:
            self._node_index = {nid: i for i, nid in enumerate(list(self.tracks) + list(self.collections))}
        return self._node_index

    def to_dgl_graph(self):
        """(g, track_ids, col_ids, features): edges as listed (CSR built on the device, ps_csr_build); features
        standardised per column with the unbiased std + 1e-12 (ps_standardize) (spotify_graph.py:41-85)."""
        track_ids, col_ids = list(self.tracks), list(self.collections)
        index, edges = self._index(), self.graph["edges"]
        ends = np.fromiter((index[e[k]] for e in edges for k in ("from", "to")), dtype=np.int64, count=2 * len(edges)).reshape(-1, 2)
        g = PSGraph.from_edges(ends[:, 0], ends[:, 1], len(track_ids), len(col_ids), nbhds_path=self.nbhds_path, base_dir=self.base_dir)
        features = None
        if self.ft_dir:
            features = standardize_features(torch.stack([torch.load(path.join(self.ft_dir, t + ".pt")) for t in track_ids], dim=0))
        self.g, self.track_ids, self.col_ids, self.features = g, track_ids, col_ids, features
        return g, track_ids, col_ids, features

    def load_positives(self, pos_pth):
        """int64 [P, 2] node-number pairs of the {"a": id, "b": id} records (spotify_graph.py:88-100)."""
        pairs, index = _read_json(pos_pth), self._index()  # tracks come first, so node number == track number
        flat = np.fromiter((index[p[k]] for p in pairs for k in ("a", "b")), dtype=np.int64, count=2 * len(pairs))
        self.positives = torch.from_numpy(flat.reshape(-1, 2).copy())
        return self.positives

    def load_positives_split(self, pos_pth, split=0.7, shuffle=True, random_seed=42):
        """(train, test): rows permuted by RandomState(random_seed), the first `split` fraction trains
        (spotify_graph.py:102-110)."""
        pos = self.load_positives(pos_pth)
        if shuffle:
            pos = pos[torch.from_numpy(np.random.RandomState(random_seed).permutation(pos.shape[0]))]
        cut = int(split * pos.shape[0])
        return pos[:cut], pos[cut:]

    def song_info(self, index_id):
        tid = list(self.tracks)[index_id]
        t = self.tracks[tid]
        return f"{t.get('name', tid)} - {t.get('artist', '')}"
