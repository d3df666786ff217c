"""B200 drop-in for the reference's `generate_positives` module (reference:
/root/reference/generate_positives.py:13-80): positive training pairs from the PPR neighbourhoods, i.e. a second
caller of the walker (`precompute_neighborhoods_topt`, one ps_walk_topt launch here), and uniformly random pairs.
Same function names, arguments and output files (`positives.json`, `positives_random.json`: [{"a": id, "b": id}])."""
from __future__ import annotations

import json
import os

import torch

import pinsage_model as psm
from spotify_graph import SpotifyGraph


def generate_positives_simple_walks(dataset, m, T):
    """m pairs (track, one of its T strongest PPR neighbours at a uniformly random rank)
    (generate_positives.py:13-45)."""
    print(f"\033[0;33mGenerating positive training pairs for {dataset.base_dir} with Personalized PageRank...\033[0m")
    g, track_ids, col_ids, _ = dataset.to_dgl_graph()
    _, nbhds = psm.precompute_neighborhoods_topt(g, len(track_ids), psm.DEF_HOPS, psm.DEF_ALPHA, psm.DEF_T_PRECOMP,
                                                 dataset.nbhds_path)
    rnd_ids = torch.randint(0, len(track_ids), (m,))
    rnd_rank = torch.randint(0, T, (m,))
    b = nbhds[rnd_ids, rnd_rank]
    return [{"a": track_ids[i], "b": track_ids[j]} for i, j in zip(rnd_ids.tolist(), b.tolist())]


def generate_positives(dataset_dir, n="auto", T=3):
    """generate_positives.py:47-56."""
    dataset = SpotifyGraph(dataset_dir, None)
    n = len(dataset.tracks) * 5 if n == "auto" else n
    positives = generate_positives_simple_walks(dataset, n, T)
    with open(os.path.join(dataset_dir, "positives.json"), "w", encoding="utf-8") as f:
        json.dump(positives, f, ensure_ascii=False, indent=2)


def generate_random_positives(dataset_dir, n="auto"):
    """generate_positives.py:58-75."""
    dataset = SpotifyGraph(dataset_dir, None)
    tracks = list(dataset.tracks.keys())
    n = len(tracks) * 2 if n == "auto" else n
    rand_a = torch.randint(0, len(tracks), (n,)).tolist()
    rand_b = torch.randint(0, len(tracks), (n,)).tolist()
    positives = [{"a": tracks[a], "b": tracks[b]} for a, b in zip(rand_a, rand_b)]
    with open(os.path.join(dataset_dir, "positives_random.json"), "w", encoding="utf-8") as f:
        json.dump(positives, f, ensure_ascii=False, indent=2)


if __name__ == "__main__":
    generate_positives("dataset_final_intersect")
    generate_random_positives("dataset_final_intersect")
