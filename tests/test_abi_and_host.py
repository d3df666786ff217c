"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares
(no compute calls without a GPU), and the host logic (CSR graph, frontier plans, batch
sampling, synthetic data) behaves like the reference's."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import ps_native
    header = open(os.path.join(ROOT, "include", "pinsage_b200.h")).read()
    declared = set(re.findall(r"\b(ps_[a-z0-9_]+)\s*\(", header))
    declared -= {"ps_graph_t", "ps_stream_t"}
    assert len(declared) >= 18
    lib = ps_native.load_library()  # dlopen only; raises AttributeError on a missing symbol
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(ps_native.EXPORTED_SYMBOLS)
    assert lib.ps_version() == 1


def test_no_cpu_fallback_without_device():
    import ps_native
    if torch.cuda.is_available():
        pytest.skip("device present")
    with pytest.raises(ps_native.NativeError, match="no CPU fallback"):
        ps_native._ensure_device()
    import pinsage_model as psm
    with pytest.raises(ps_native.NativeError):
        psm.PinSageModel(None, 10, 2, (8, 8, 4), 500, 0.85, 3, (torch.zeros(10, 3, dtype=torch.float64), torch.zeros(10, 3, dtype=torch.int64)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gcn-song-embeddings_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn


def test_psgraph_matches_dgl_semantics(golden):
    from ps_graph import PSGraph
    rng = np.random.RandomState(0)
    nt, nc = 30, 8
    t = rng.randint(0, nt, 100); c = rng.randint(0, nc, 100) + nt
    src, dst = np.r_[t, c], np.r_[c, t]
    g = PSGraph.from_edges(src, dst, nt, nc, nbhds_path="x.pt")
    assert g.number_of_nodes() == len(g) == nt + nc and g.nbhds_path == "x.pt"
    for v in (0, 5, nt + 2):
        want = dst[src == v]  # insertion order per source, multi-edges kept (DGL)
        assert np.array_equal(g.successors(v).numpy(), want)
    assert int(g.out_degrees().sum()) == 200 and torch.equal(g.in_degrees(), g.out_degrees())
    s2, d2 = g.edges()
    assert sorted(zip(s2.tolist(), d2.tolist())) == sorted(zip(src.tolist(), dst.tolist()))
    with pytest.raises(IndexError):
        PSGraph.from_edges([0, 99], [1, 2], nt, nc)


@pytest.mark.parametrize("tag,L,T", [("L2T3", 2, 3), ("L3T5", 3, 5), ("L2T10", 2, 10)])
def test_build_plan_matches_reference_frontier(golden, tag, L, T):
    """Engine plans (CPU tensors) reproduce relevant_nodes_per_layer_precomp on the distinct
    batch nodes, and the backward transpose lists every (target, slot) pair exactly once."""
    import ps_engine
    g = golden("frontier")
    tab = ps_engine.NeighborTable(torch.from_numpy(g["w"]), torch.from_numpy(g["nodes"]), device="cpu")
    top = torch.unique(torch.from_numpy(g[f"{tag}_nodeset"]))
    plan = ps_engine.build_plan(top, L, T, tab, need_backward=True)
    S = oracle.relevant_nodes_per_layer_precomp(top.numpy(), L, T, (g["w"], g["nodes"]))
    # lower layers of the reference frontier are those of the ORIGINAL (duplicated) nodeset too
    S_dup = [(g[f"{tag}_ns{l}"], g[f"{tag}_w{l}"], g[f"{tag}_nb{l}"]) for l in range(L)]
    for l in range(L - 1):
        assert np.array_equal(S[l][0], S_dup[l][0])
    for l, (ns, w, nb) in enumerate(S):
        lp = plan.layers[l]
        if l == 0:
            ids = lp.zrows.long()[lp.nbz.long()]
            assert np.array_equal(lp.self_rows.numpy(), ns)
        else:
            prev = torch.from_numpy(S[l - 1][0])
            ids = prev[lp.nbz.long()]
            assert np.array_equal(prev[lp.self_rows.long()].numpy(), ns)
        assert np.array_equal(ids.numpy(), nb)
        assert np.array_equal(lp.w.numpy(), w.astype(np.float32))
        flat = lp.nbz.reshape(-1)
        assert sorted(lp.pair_q.tolist()) == list(range(flat.numel()))
        for u in range(lp.nz):
            seg = lp.pair_q[lp.seg_off[u]:lp.seg_off[u + 1]].long()
            assert bool((flat[seg] == u).all())
    with pytest.raises(IndexError):
        ps_engine.build_plan(torch.tensor([10 ** 6]), L, T, tab, False)
    with pytest.raises(ValueError):
        ps_engine.build_plan(top, L, 999, tab, False)


def test_unique_inverse_dense_map_equals_torch_unique():
    import ps_engine
    torch.manual_seed(3)
    ids = torch.randint(0, 50_000, (200_000,), dtype=torch.int32)
    u, inv = ps_engine._unique_inverse(ids, 50_000, {})
    tu, tinv = torch.unique(ids, return_inverse=True)
    assert torch.equal(u, tu.to(torch.int64)) and torch.equal(inv, tinv)
    small = torch.randint(0, 50_000, (100,))
    u, inv = ps_engine._unique_inverse(small, 50_000, {})
    assert torch.equal(u[inv], small)


@pytest.mark.parametrize("n,P,B", [(500, 2000, 64), (200_000, 3_000_000, 512)])
def test_batch_sampling_properties(n, P, B):
    """Same guarantees as the reference's sample_batch with easy negatives: distinct rows of
    `positives`, distinct negatives outside the batch (pinsage_training.py:53-77)."""
    import pinsage_training as pst
    torch.manual_seed(0)
    positives = torch.unique(torch.randint(0, n, (P, 2)), dim=0)
    all_ids = torch.arange(n)
    batch, nodeset = pst.sample_batch(all_ids, positives, B, None, hard_negatives=False)
    assert batch.shape == (B, 3) and batch.dtype == torch.int64
    oracle.check_batch_properties(batch.numpy(), positives.numpy(), n)
    assert len({tuple(r) for r in batch[:, :2].tolist()}) == B  # pairs drawn without repetition
    assert torch.equal(nodeset, batch.flatten().unique())
    # hard negatives: rank window and the reference's row quirk
    nb = torch.randint(0, n, (n, 100))
    hb, _ = pst.sample_batch(all_ids, positives, B, (None, nb), hard_negatives=True, hn_min=10, hn_max=20)
    assert all(int(hb[i, 2]) in nb[i, 10:20].tolist() for i in range(B))  # rows 0..B-1, pinsage_training.py:84
    fixed, _ = pst.sample_hard_negatives(all_ids, hb[:, :2], (None, nb), 10, 20, reference_compat=False)
    assert all(int(fixed[i, 2]) in nb[int(fixed[i, 0]), 10:20].tolist() for i in range(B))


def test_distinct_randint_is_uniform_without_replacement():
    import pinsage_training as pst
    torch.manual_seed(1)
    hits = torch.zeros(100_000)
    for _ in range(50):
        s = pst._distinct_randint(100_000, 2000, "cpu")
        assert s.numel() == 2000 and s.unique().numel() == 2000
        hits[s] += 1
    assert abs(float(hits.mean()) - 1.0) < 1e-6 and float(hits.max()) <= 8


def test_synthetic_graph_properties():
    import ps_synth
    nt, nc = 5000, 800
    indptr, indices, e = ps_synth.bipartite_csr(nt, nc, 60_000, seed=9)
    deg = indptr[1:] - indptr[:-1]
    assert int(deg.min()) >= 1 and indices.numel() == 2 * e == int(indptr[-1])
    t_side = indices[: int(indptr[nt])]; c_side = indices[int(indptr[nt]):]
    assert int(t_side.min()) >= nt and int(c_side.max()) < nt  # bipartite
    src = torch.repeat_interleave(torch.arange(nt + nc), deg)
    fwd = set(zip(src.tolist(), indices.tolist()))
    assert all((b, a) in fwd for a, b in list(fwd)[:2000])  # both directions listed
    assert len(fwd) == 2 * e  # no duplicate memberships
    pos = ps_synth.cooccurrence_positives(indptr, indices, nt, 500)
    for a, b in pos[:50].tolist():
        ca = set(indices[indptr[a]:indptr[a + 1]].tolist()); cb = set(indices[indptr[b]:indptr[b + 1]].tolist())
        assert a != b and ca & cb
    f = ps_synth.features(nt, 16)
    assert torch.allclose(f.mean(0), torch.zeros(16), atol=1e-5) and torch.allclose(f.std(0), torch.ones(16), atol=1e-4)


def test_metrics_match_oracle(golden):
    g = golden("metrics_knn")
    import ps_eval
    knn = torch.from_numpy(g["rnd_knn"]); pos = torch.from_numpy(g["rnd_pos"])
    for K, hr, m in zip(g["rnd_K"], g["rnd_hr"], g["rnd_mrr"]):
        assert ps_eval.hit_rate(knn, pos, int(K)) == pytest.approx(hr, abs=1e-12)
        assert ps_eval.mrr(knn, pos, int(K)) == pytest.approx(m, abs=1e-12)
    for K, hr, m in zip(g["toy_K"], g["toy_hr"], g["toy_mrr"]):
        assert ps_eval.hit_rate(torch.from_numpy(g["toy_knn"]), torch.from_numpy(g["toy_pos"]), int(K)) == pytest.approx(hr)
        assert ps_eval.mrr(torch.from_numpy(g["toy_knn"]), torch.from_numpy(g["toy_pos"]), int(K)) == pytest.approx(m)


def _rewrite_dataset(g, d):
    """Recreate the dataset files of tests/golden/dataset.npz (the content the reference's loader was run on)."""
    import json
    os.makedirs(os.path.join(d, "features_openl3"), exist_ok=True)
    json.dump({t: {"name": t, "artist": "a"} for t in g["track_ids"].tolist()}, open(os.path.join(d, "tracks.json"), "w"))
    json.dump({c: {} for c in g["col_ids"].tolist()}, open(os.path.join(d, "collections.json"), "w"))
    json.dump({"tracks": g["track_ids"].tolist(), "collections": g["col_ids"].tolist(),
               "edges": [{"from": a, "to": b} for a, b in zip(g["edges_from"].tolist(), g["edges_to"].tolist())]},
              open(os.path.join(d, "graph.json"), "w"))
    for t, row in zip(g["track_ids"].tolist(), g["raw_features"]):
        torch.save(torch.from_numpy(row.copy()), os.path.join(d, "features_openl3", t + ".pt"))
    json.dump([{"a": a, "b": b} for a, b in zip(g["pos_a"].tolist(), g["pos_b"].tolist())], open(os.path.join(d, "positives_lfm.json"), "w"))


def test_spotify_graph_matches_reference_loader(golden, tmp_path):
    """Drop-in SpotifyGraph == the reference's loader on the same files: node numbering, adjacency (order of
    successors included), standardised features, positives and the 70/30 split."""
    from spotify_graph import SpotifyGraph
    g = golden("dataset")
    d = str(tmp_path)
    _rewrite_dataset(g, d)
    ds = SpotifyGraph(d, os.path.join(d, "features_openl3"))
    graph, track_ids, col_ids, features = ds.to_dgl_graph()
    assert track_ids == g["track_ids"].tolist() and col_ids == g["col_ids"].tolist()
    assert np.array_equal(graph.indptr.numpy(), g["indptr"]) and np.array_equal(graph.indices.numpy(), g["indices"])
    assert os.path.basename(graph.nbhds_path) == str(g["nbhds_path_tail"]) and graph.base_dir == d
    assert np.allclose(features.numpy(), g["features"], rtol=1e-6, atol=1e-7)
    pos = ds.load_positives(os.path.join(d, "positives_lfm.json"))
    assert np.array_equal(pos.numpy(), g["positives"])
    train, test = ds.load_positives_split(os.path.join(d, "positives_lfm.json"))
    assert np.array_equal(train.numpy(), g["train"]) and np.array_equal(test.numpy(), g["test"])


def test_results_table_matches_reference(golden):
    """eval.compute_results_table == the reference's (eval.py:413-443) on its own output: HR@10/100/500, MRR@1000,
    low-degree MRR (two thresholds) and low-co-occurrence MRR."""
    import eval as ev
    from ps_graph import PSGraph
    g = golden("results_table")
    graph = PSGraph.from_edges(g["from"], g["to"], int(g["n_tracks"]), int(g["n_cols"]))
    kd = ev.KnnDict(); kd["m"] = (None, torch.from_numpy(g["knn"]))
    for thr in (1, 3):
        table = ev.compute_results_table(kd, torch.from_numpy(g["test_pos"]), graph, times=False, degree_thr=thr)
        for col in ("hr (k=10)", "hr (k=100)", "hr (k=500)", "mrr", "low-degree accuracy", "low-co accuracy"):
            assert float(table.loc["m", col]) == pytest.approx(float(g[f"thr{thr}/{col}"]), abs=1e-12), (thr, col)


def test_results_table_host_logic():
    import eval as ev
    from ps_graph import PSGraph
    knn = torch.tensor([[1, 2, 3], [0, 2, 3], [3, 1, 0], [2, 0, 1]])
    test_pos = torch.tensor([[0, 2], [1, 3], [2, 0], [3, 3]])
    g = PSGraph.from_edges([0, 4, 1, 4, 2, 4, 3, 5, 0, 5], [4, 0, 4, 1, 4, 2, 5, 3, 5, 0], 4, 2)
    kd = ev.KnnDict(); kd["m"] = (None, knn); kd.times["m"] = (1.0, 2.0, 3.0)
    table = ev.compute_results_table(kd, test_pos, g, degree_thr=1)
    assert table.loc["m", "hr (k=10)"] == pytest.approx(0.75)
    assert table.loc["m", "mrr"] == pytest.approx((1 / 2 + 1 / 3 + 1 / 3 + 1 / 1000) / 4)
    assert table.loc["m", "low-degree accuracy"] == pytest.approx((1 / 3 + 1 / 3 + 1 / 1000) / 3)  # node 0 has degree 2
    assert table.loc["m", "t (knn)"] == 3.0


def test_batch_construction_equals_reference_draws(golden):
    """sample_batch (easy and hard negatives, incl. the reference's row-gather quirk under reference_compat) and
    batch_variance reproduce the reference's outputs under the same torch seed (pinsage_training.py:53-103)."""
    import pinsage_training as pst
    g = golden("batches")
    n, B = int(g["n"]), int(g["B"])
    positives = torch.from_numpy(g["positives"]); all_ids = torch.arange(n, dtype=torch.int64)
    nbhds = (None, torch.from_numpy(g["nb_nodes"]))
    for seed in (0, 1, 2):
        torch.manual_seed(seed)
        b, ns = pst.sample_batch(all_ids, positives, B, nbhds, hard_negatives=False)
        assert np.array_equal(b.numpy(), g[f"easy{seed}"]) and np.array_equal(ns.numpy(), g[f"easy{seed}_nodeset"])
        torch.manual_seed(seed)
        b, ns = pst.sample_batch(all_ids, positives, B, nbhds, hard_negatives=True, hn_min=10, hn_max=100)
        assert np.array_equal(b.numpy(), g[f"hard{seed}"]) and np.array_equal(ns.numpy(), g[f"hard{seed}_nodeset"])
    # the fixed variant gathers the queries' own rows
    torch.manual_seed(0)
    pos_batch = pst.sample_positives_with_rep(positives, B)
    fixed, _ = pst.sample_hard_negatives(all_ids, pos_batch, nbhds, 10, 100, reference_compat=False)
    assert all(int(fixed[i, 2]) in g["nb_nodes"][int(fixed[i, 0]), 10:100] for i in range(B))
    assert float(pst.batch_variance(torch.from_numpy(g["var_h"]))) == pytest.approx(float(g["var"]), rel=1e-6)


def test_bench_roofline_helpers():
    """bench.py's roofline arithmetic (no GPU): the aggregation kernel is reported in SURVEY 8d's algorithmic unit with the
    kernel-addressed and DRAM-measured figures beside it; a GEMM against the tensor peak; the walker against HBM and the
    measured random-access ceiling."""
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv
    peaks = {"hbm_gbs": 6548.5, "tflops": 1381.6, "source": "measured"}
    n, T, din, dh, do = 98_499, 50, 256, 512, 128
    per_target_kernel = T * dh * 4 + din * 4 + T * 8 + 4 + (din + dh) * 4 + 4
    summary = {"aggregate_fwd_l0": {"ms": 1.166 * 30, "launches": 30, "flops": 0.0, "bytes": 30.0 * n * per_target_kernel},
               "gemm_q_wgrad_l0": {"ms": 1.07 * 30, "launches": 30, "flops": 30 * 1.72e11, "bytes": 0.0}}
    r = bench.roofline_from_profile(summary, 30, peaks, dims=(din, dh, do, T))
    assert r["kernel"] == "aggregate_fwd_l0" and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert r["algorithmic_bytes_per_target"] == (T + 1) * din * 4 + T * 8 + do * 4 == 53_136
    assert abs(r["achieved"] - n * 53_136 / 1.166e-3 / 1e9) < 1.0 and abs(r["frac"] - r["achieved"] / 6548.5) < 1e-3
    assert r["kernel_addressed_gbs"] > r["achieved"] and r["targets_per_launch"] == n
    if r["traffic"]:
        assert abs(r["frac_dram_measured"] - r["traffic"] / 1.166e-3 / 1e9 / 6548.5) < 1e-3
    summary["gemm_q_wgrad_l0"]["ms"] = 2.0 * 30
    g = bench.roofline_from_profile(summary, 30, peaks, dims=(din, dh, do, T))
    assert g["kernel"] == "gemm_q_wgrad_l0" and g["bound"] == "tensor" and abs(g["achieved"] - 1.72e11 / 2.0e-3 / 1e12) < 0.01
    w = bench.walk_roofline(1_000_000, 500, 9.87, peaks)
    assert abs(w["frac_of_hbm"] - 5e8 * 28 / 9.87e-3 / 1e9 / 6548.5) < 1e-3
    if "random_access_bound" in w:
        assert 0.5 < w["random_access_bound"]["frac_of_bound"] <= 1.05 and w["frac_of_hbm_dram_measured"] > w["frac_of_hbm"]
