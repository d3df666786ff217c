"""PinSage-side slice of the reference's `dashboard.py` (reference:
/root/reference/dashboard.py:48-79, 82-172, 175-191): `train` and `eval` for PinSage.
`prepare` (Spotify crawler + audio features) needs the network and is out of scope.

    python dashboard.py train|eval|all [DATA_DIR] [FEATURES_SUBDIR] [POSITIVES_FILE]
"""
from __future__ import annotations

import os
import sys

import eval as ev
import pinsage_training as pt
from baselines import EmbLoader
from pinsage_training import PinSage
from spotify_graph import SpotifyGraph

DATA_DIR = "./dataset_final_intersect"
FEATURES = "features_openl3"
POSITIVES = "positives_lfm.json"


def train_pinsage(data_dir=DATA_DIR, features=FEATURES, positives_file=POSITIVES, run_name=None, **overrides):
    """dashboard.py:48-79 for one feature set: train, then save the per-track embeddings."""
    dataset = SpotifyGraph(data_dir, os.path.join(data_dir, features))
    g, track_ids, col_ids, feats = dataset.to_dgl_graph()
    positives = dataset.load_positives(os.path.join(data_dir, positives_file))
    os.makedirs(pt.BASE_RUN_DIR, exist_ok=True)
    pinsage = PinSage(g, len(track_ids), feats, positives, log=overrides.pop("log", False), load_save=overrides.pop("load_save", True))
    setattr(pinsage, "run_name", run_name or f"pinsage_{features.replace('features_', '')}_ft")
    for k, v in overrides.items():
        setattr(pinsage, k, v)
    pinsage.train()
    pt.save_embeddings(pinsage, dataset)
    return pinsage


def eval_baselines(data_dir=DATA_DIR, features=FEATURES, positives_file=POSITIVES, run_name=None, save_dir="./eval_cache", k=None):
    """dashboard.py:82-172 restricted to the PinSage rows: kNN lists of the saved embeddings,
    hit-rate / MRR table on the 30 % test split."""
    dataset = SpotifyGraph(data_dir, os.path.join(data_dir, features))
    g, track_ids, col_ids, feats = dataset.to_dgl_graph()
    train_pos, test_pos = dataset.load_positives_split(os.path.join(data_dir, positives_file))
    run_name = run_name or f"pinsage_{features.replace('features_', '')}_ft"
    models = {"PinsageBase": EmbLoader(os.path.join(pt.BASE_RUN_DIR, run_name, "emb"))}
    knn_dict = ev.get_knn_dict(models, g, track_ids, train_pos, test_pos, feats, save_dir, k=k)
    table = ev.compute_results_table(knn_dict, test_pos, g, degree_thr=3)
    print(table)
    return table


if __name__ == "__main__":
    action = sys.argv[1] if len(sys.argv) > 1 else "all"
    args = sys.argv[2:5]
    if action in ("train", "all"):
        train_pinsage(*args)
    if action in ("eval", "all"):
        eval_baselines(*args)
