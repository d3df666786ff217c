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

from ps_graph import PSGraph


class SpotifyGraph():

    def __init__(self, dir, features_dir):
        self.base_dir = dir
        self.nbhds_path = os.path.join(self.base_dir, "neighborhoods.pt")
        self.tracks_pth = path.join(dir, "tracks.json")
        self.col_pth = path.join(dir, "collections.json")
        self.graph_pth = path.join(dir, "graph.json")
        self.img_dir = path.join(dir, "images")
        self.clip_dir = path.join(dir, "clips")
        print("Loading graph...")
        with open(self.tracks_pth, "r", encoding="utf-8") as f:
            self.tracks = json.load(f)
        with open(self.col_pth, "r", encoding="utf-8") as f:
            self.collections = json.load(f)
        with open(self.graph_pth, "r", encoding="utf-8") as f:
            self.graph = json.load(f)
        self.ft_dir = features_dir if features_dir is not None and os.path.isdir(features_dir) else None
        self.features_dict = {}

    def to_dgl_graph(self):
        """(g, track_ids, col_ids, features): nodes are numbered by position in
        list(tracks) + list(collections); edges as listed; features standardised per column
        with the unbiased std + 1e-12 (spotify_graph.py:41-85)."""
        track_ids = list(self.tracks)
        col_ids = list(self.collections)
        index_map = {nid: i for i, nid in enumerate(track_ids + col_ids)}
        edges = self.graph["edges"]
        src = np.fromiter((index_map[e["from"]] for e in edges), dtype=np.int64, count=len(edges))
        dst = np.fromiter((index_map[e["to"]] for e in edges), dtype=np.int64, count=len(edges))
        g = PSGraph.from_edges(src, dst, len(track_ids), len(col_ids), nbhds_path=self.nbhds_path, base_dir=self.base_dir)
        if self.ft_dir:
            features = torch.stack([torch.load(os.path.join(self.ft_dir, t + ".pt")) for t in track_ids], dim=0)
            mean = features.mean(dim=0)
            std = features.std(dim=0, unbiased=True) + 1e-12
            features = (features - mean) / std
        else:
            features = None
        self.g, self.track_ids, self.col_ids, self.features = g, track_ids, col_ids, features
        return g, track_ids, col_ids, features

    def load_positives(self, pos_pth):
        """int64 [P, 2] index pairs (spotify_graph.py:88-100)."""
        with open(pos_pth, "r", encoding="utf-8") as f:
            positives = json.load(f)
        index_map = {nid: i for i, nid in enumerate(list(self.tracks))}
        a = torch.tensor([index_map[pair["a"]] for pair in positives], dtype=torch.int64)
        b = torch.tensor([index_map[pair["b"]] for pair in positives], dtype=torch.int64)
        pos = torch.stack((a, b), dim=1)
        self.positives = pos
        return pos

    def load_positives_split(self, pos_pth, split=0.7, shuffle=True, random_seed=42):
        """(train, test): RandomState(random_seed).permutation, first `split` fraction trains
        (spotify_graph.py:102-110)."""
        pos = self.load_positives(pos_pth)
        n = pos.shape[0]
        if shuffle:
            index = np.random.RandomState(random_seed).permutation(n)
            pos = pos[index, :]
        cut_point = int(split * n)
        return pos[:cut_point, :], pos[cut_point:, :]

    def song_info(self, index_id):
        track_ids = list(self.tracks)
        t = self.tracks[track_ids[index_id]]
        return f"{t.get('name', track_ids[index_id])} - {t.get('artist', '')}"
