"""Synthetic track-collection datasets in the shapes BASELINE.json names (the reference's
own datasets are not in its checkout).  Pure torch index
work, device-agnostic: small graphs on the CPU for tests, the 1 M / 200 k / 40 M bench
graph directly in HBM.
"""
from __future__ import annotations

import torch

import ps_native
from ps_graph import PSGraph


def bipartite_csr(n_tracks: int, n_cols: int, n_edges: int, seed: int = 1234, device="cpu",
                  max_col_size: int = 5000, skew: float = 2.0):
    """Bipartite graph with Zipf-like collection sizes (shape 1.2, clipped to
    [2, max_col_size], scaled to ~n_edges memberships) and popularity-skewed track
    endpoints; duplicate memberships removed; every node has degree >= 1.
    Returns (indptr int64 [N+C+1], indices int32 [2E'], E')."""
    gen = torch.Generator(device=device).manual_seed(seed)
    u = torch.rand(n_cols, generator=gen, device=device, dtype=torch.float64)
    raw = (1.0 - u).clamp_min(1e-12).pow(-1.0 / 1.2)
    cap = min(max_col_size, n_tracks)
    sizes = (raw * (n_edges / raw.sum())).round().clamp(2, cap)
    # one rescale pass after clipping so the total lands near n_edges
    sizes = (sizes * (n_edges / sizes.sum())).round().clamp(2, cap).to(torch.int64)
    col = torch.repeat_interleave(torch.arange(n_cols, device=device), sizes)
    ut = torch.rand(col.numel(), generator=gen, device=device, dtype=torch.float64)
    relabel = torch.randperm(n_tracks, generator=gen, device=device)  # popular tracks scattered over the id space
    track = relabel[(ut.pow(skew) * n_tracks).to(torch.int64).clamp_max(n_tracks - 1)]
    # every track at least once
    missing = torch.ones(n_tracks, dtype=torch.bool, device=device)
    missing[track] = False
    miss = missing.nonzero().flatten()
    if miss.numel():
        track = torch.cat([track, miss])
        col = torch.cat([col, torch.randint(0, n_cols, (miss.numel(),), generator=gen, device=device)])
    key = torch.unique(col * n_tracks + track)
    del col, track, ut
    e = key.numel()
    # CSR halves built separately (a stable argsort over the 2E directed entries needs several times the memory):
    # collection rows list their tracks in ascending order (the order of `key`), track rows their collections.
    col_deg = torch.bincount(key // n_tracks, minlength=n_cols)
    col_rows = (key % n_tracks).to(torch.int32)
    key2 = torch.sort((key % n_tracks) * n_cols + key // n_tracks).values
    del key
    track_deg = torch.bincount(key2 // n_cols, minlength=n_tracks)
    track_rows = (key2 % n_cols + n_tracks).to(torch.int32)
    del key2
    n = n_tracks + n_cols
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(torch.cat([track_deg, col_deg]), 0)
    return indptr, torch.cat([track_rows, col_rows]), e


def make_graph(n_tracks, n_cols, n_edges, seed=1234, device="cpu", **kw) -> PSGraph:
    indptr, indices, _ = bipartite_csr(n_tracks, n_cols, n_edges, seed, device, **kw)
    g = PSGraph(indptr.cpu(), indices.cpu(), n_tracks, n_cols)
    if indptr.is_cuda:  # keep the device copy instead of uploading again
        g._handle = ps_native.GraphHandle(indptr, indices, n_tracks, n_cols)
    return g


def features(n_tracks, dim, seed=1, device="cpu"):
    """N(0,1) features standardised per column like SpotifyGraph.to_dgl_graph
    (spotify_graph.py:77-79)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn((n_tracks, dim), generator=gen, device=device, dtype=torch.float32)
    mean, std = x.mean(0), x.std(0, unbiased=True)
    return x.sub_(mean).div_(std + 1e-12)  # in place: the cfg4 table is 41 GB


def cooccurrence_positives(indptr, indices, n_tracks, n_pairs, seed=2):
    """Random (a, b) track pairs that share a collection (2-step co-occurrence), a != b."""
    device = indptr.device
    gen = torch.Generator(device=device).manual_seed(seed)
    track_entries = int(indptr[n_tracks])
    e = torch.randint(0, track_entries, (n_pairs,), generator=gen, device=device)
    a = torch.searchsorted(indptr[: n_tracks + 1], e, right=True) - 1
    c = indices[e].to(torch.int64)
    deg = indptr[c + 1] - indptr[c]
    r = (torch.rand(n_pairs, generator=gen, device=device, dtype=torch.float64) * deg).to(torch.int64).clamp_max(deg - 1)
    b = indices[indptr[c] + r].to(torch.int64)
    keep = a != b
    return torch.stack([a[keep], b[keep]], 1)


def write_dataset(out_dir, n_tracks, n_cols, n_edges, feat_dim, n_pos, seed=0, features_name="features_openl3",
                  positives_name="positives_lfm.json"):
    """Write a synthetic dataset in the reference's on-disk schema (tracks.json, collections.json, graph.json
    with both edge directions, one <track_id>.pt feature vector per track, positives json); see
    spotify_graph.py for the schema.  Small sizes only (JSON + one file per track)."""
    import json
    import os
    indptr, indices, _ = bipartite_csr(n_tracks, n_cols, n_edges, seed=seed)
    track_ids = [f"t{i:06d}" for i in range(n_tracks)]
    col_ids = [f"c{i:06d}" for i in range(n_cols)]
    all_ids = track_ids + col_ids
    deg = (indptr[1:] - indptr[:-1])
    src = torch.repeat_interleave(torch.arange(n_tracks + n_cols), deg)
    edges = [{"from": all_ids[a], "to": all_ids[b]} for a, b in zip(src.tolist(), indices.tolist())]
    os.makedirs(os.path.join(out_dir, features_name), exist_ok=True)
    with open(os.path.join(out_dir, "tracks.json"), "w") as f:
        json.dump({t: {"name": f"song {i}", "artist": f"artist {i % 7}"} for i, t in enumerate(track_ids)}, f)
    with open(os.path.join(out_dir, "collections.json"), "w") as f:
        json.dump({c: {"name": f"playlist {i}"} for i, c in enumerate(col_ids)}, f)
    with open(os.path.join(out_dir, "graph.json"), "w") as f:
        json.dump({"tracks": track_ids, "collections": col_ids, "edges": edges}, f)
    gen = torch.Generator().manual_seed(seed + 1)
    raw = torch.randn((n_tracks, feat_dim), generator=gen) * 2.0 + 0.5
    for i, t in enumerate(track_ids):
        torch.save(raw[i].clone(), os.path.join(out_dir, features_name, t + ".pt"))
    pos = cooccurrence_positives(indptr, indices, n_tracks, n_pos, seed=seed + 2)
    with open(os.path.join(out_dir, positives_name), "w") as f:
        json.dump([{"a": track_ids[a], "b": track_ids[b]} for a, b in pos.tolist()], f)
    return out_dir
