"""GPU parity tests for K1/K2 (walker + top-T) through the C ABI (ctypes)."""
import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nat():
    import ps_native
    return ps_native


@pytest.fixture(params=[0, 1], ids=["sort_kernel", "hash_kernel"], autouse=True)
def walk_algo(request, nat):
    """Every test of this module runs on both implementations of ps_walk_topt / ps_trace_topt (ps_walk_algo): the
    sort-based kernel (n_hops <= 512, T <= 256) and the hash-table kernel (forced; it also takes the larger shapes)."""
    old = nat.walk_algo(request.param)
    yield request.param
    nat.walk_algo(old)


def _graph(nat, indptr, indices, n_tracks):
    n_cols = len(indptr) - 1 - n_tracks
    return nat.GraphHandle(torch.from_numpy(np.asarray(indptr)), torch.from_numpy(np.asarray(indices)), n_tracks, n_cols)


@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("T", [3, 100])
def test_trace_topt_matches_reference(nat, golden, tag, T):
    """K2 fed the reference's own traces: weights bit-equal, neighbour sets equal modulo
    tie groups / zero-weight fillers."""
    g = golden("walk_topt")
    trace = torch.from_numpy(g[f"{tag}_trace"].astype(np.int64))
    nodeset = torch.from_numpy(g[f"{tag}_nodeset"])
    w, nb = nat.trace_topt(trace, nodeset, T)
    w, nb = w.cpu().numpy(), nb.cpu().numpy()
    oracle.check_topt_against_reference(g[f"{tag}_w_T{T}"], g[f"{tag}_nb_T{T}"], w, nb, trace.numpy(), nodeset.numpy())
    ow, onb = oracle.topt_from_trace(trace.numpy(), nodeset.numpy(), T)
    assert np.array_equal(ow.view(np.int64), w.view(np.int64))  # bit-exact vs the oracle, order included
    assert np.array_equal(onb, nb)


@pytest.mark.parametrize("n_hops,alpha,fixed_len,T", [(500, 0.85, 0, 100), (37, 0.5, 0, 5), (1000, 0.85, 0, 50),
                                                      (64, 0.0, 0, 10), (96, 1.0, 0, 10), (400, 0.85, 4, 50), (15, 0.85, 5, 3),
                                                      (500, 0.85, 0, 3), (512, 0.3, 0, 17), (256, 0.85, 0, 200), (129, 0.6, 0, 256),
                                                      (300, 0.85, 3, 50)])
def test_walker_bit_exact_vs_oracle(nat, golden, n_hops, alpha, fixed_len, T):
    """K1: the Philox walker's traces equal the CPU restatement's bit for bit, for any
    launch shape, and the fused top-T equals the oracle's reduction of that trace."""
    g = golden("walk_topt")
    for tag in ("a", "b"):
        nt = int(g[f"{tag}_n_tracks"])
        gh = _graph(nat, g[f"{tag}_indptr"], g[f"{tag}_indices"], nt)
        src = torch.arange(0, nt, 3 if tag == "a" else 41)
        out = nat.walk_topt(gh, src, n_hops, alpha, T, seed=0xC0FFEE1234, fixed_len=fixed_len, want_i32=True, want_trace=True)
        want = oracle.do_random_walks_philox(g[f"{tag}_indptr"], g[f"{tag}_indices"], src.numpy(), n_hops, alpha,
                                             0xC0FFEE1234, fixed_len)
        assert np.array_equal(out["trace"].cpu().numpy().astype(np.int64), want)
        ow, onb = oracle.topt_from_trace(want, src.numpy(), T)
        assert np.array_equal(out["weights"].cpu().numpy().view(np.int64), ow.view(np.int64))
        assert np.array_equal(out["nodes"].cpu().numpy(), onb)
        assert np.array_equal(out["nodes_i32"].cpu().numpy().astype(np.int64), onb)
        assert np.array_equal(out["weights_f32"].cpu().numpy(), ow.astype(np.float32))


def test_walker_matches_reference_distribution(nat, golden):
    """Statistical parity with the reference's mt19937 walker (visit histograms)."""
    g = golden("walk_dist")
    nt = int(g["n_tracks"])
    gh = _graph(nat, g["indptr"], g["indices"], nt)
    n_hops = int(g["n_hops"])
    # pool 5 walks of n_hops/5 (every restart returns to the source, so pooled visit counts follow the same law)
    parts = [nat.walk_topt(gh, torch.from_numpy(g["nodeset"]), n_hops // 5, 0.85, 8, seed=99 + k, want_trace=True) for k in range(5)]
    trace = np.concatenate([p["trace"].cpu().numpy() for p in parts], axis=1)
    out = nat.trace_topt(torch.from_numpy(trace.astype(np.int64))[:, :16384], torch.from_numpy(g["nodeset"]), 8)
    out = {"nodes": out[1]}
    for i in range(len(g["nodeset"])):
        ours = np.bincount(trace[i], minlength=g["counts"].shape[1]) / n_hops
        ref = g["counts"][i] / n_hops
        assert 0.5 * np.abs(ours - ref).sum() < 0.05
    # top-8 neighbourhoods overlap the reference's heavily
    ref_top = np.argsort(-g["counts"].astype(np.float64) * (np.arange(g["counts"].shape[1])[None] != g["nodeset"][:, None]), axis=1)[:, :8]
    got = out["nodes"].cpu().numpy()
    overlap = np.mean([len(set(ref_top[i]) & set(got[i])) / 8 for i in range(len(got))])
    assert overlap >= 0.7


def test_walker_properties_large(nat):
    """Size-independent properties on a graph far beyond what the oracle handles quickly."""
    import ps_synth
    n_tracks, n_cols = 200_000, 40_000
    g = ps_synth.make_graph(n_tracks, n_cols, 4_000_000, seed=5, device="cuda")
    src = torch.arange(0, n_tracks, 7, device="cuda")
    out = nat.walk_topt(g.device(), src, 500, 0.85, 100, seed=1, want_i32=True)
    w, nb = out["weights"], out["nodes"]
    assert bool((w[:, :-1] >= w[:, 1:]).all())                      # sorted descending
    assert bool((w.sum(1) <= 1.0 + 1e-12).all()) and bool((w >= 0).all())
    assert bool(((w * 500).round() == w * 500).all())               # exact multiples of 1/n_hops
    assert bool((nb >= 0).all()) and bool((nb < n_tracks).all())    # items only, never collections
    assert bool(((nb != src[:, None]) | (w == 0)).all())            # self only in zero-weight slots
    s, _ = torch.sort(torch.where(w > 0, nb, -torch.arange(1, 101, device="cuda")[None, :].expand_as(nb)), dim=1)
    assert bool((s[:, 1:] != s[:, :-1]).all())                      # no duplicate neighbours
    again = nat.walk_topt(g.device(), src, 500, 0.85, 100, seed=1)
    assert torch.equal(again["nodes"], nb) and torch.equal(again["weights"], w)  # deterministic in the seed
    other = nat.walk_topt(g.device(), src, 500, 0.85, 100, seed=2)
    assert not torch.equal(other["nodes"], nb)


@pytest.mark.parametrize("wide_indptr", [False, True])
def test_walker_bit_exact_on_large_graph(nat, wide_indptr):
    """Traces and top-T of 384 sources of a 200 k-track / 4 M-edge graph (skewed degrees, hub playlists of 5000
    tracks) equal the oracle's bit for bit -- on the 4-byte row-offset path and on the 8-byte one that graphs with
    >= 2^32 CSR entries take (BASELINE.json configs[3]), forced here through ps_graph_use_indptr32."""
    import ps_synth
    n_tracks, n_cols = 200_000, 40_000
    g = ps_synth.make_graph(n_tracks, n_cols, 4_000_000, seed=5, device="cuda")
    gh = g.device()
    indptr, indices = g.indptr.numpy(), g.indices.numpy()
    deg = np.diff(indptr[: n_tracks + 1])
    src = np.unique(np.concatenate([np.argsort(-deg)[:64], np.argsort(deg)[:64], np.random.RandomState(1).randint(0, n_tracks, 256)]))
    old = gh.use_indptr32(not wide_indptr)
    try:
        for n_hops, alpha, fixed_len, T in ((500, 0.85, 0, 100), (300, 0.85, 3, 50)):
            out = nat.walk_topt(gh, torch.from_numpy(src), n_hops, alpha, T, seed=0xBADC0DE5, fixed_len=fixed_len, want_i32=True, want_trace=True)
            want = oracle.do_random_walks_philox(indptr, indices, src, n_hops, alpha, 0xBADC0DE5, fixed_len)
            assert np.array_equal(out["trace"].cpu().numpy().astype(np.int64), want)
            ow, onb = oracle.topt_from_trace(want, src, T)
            assert np.array_equal(out["weights"].cpu().numpy().view(np.int64), ow.view(np.int64))
            assert np.array_equal(out["nodes"].cpu().numpy(), onb)
            assert np.array_equal(out["nodes_i32"].cpu().numpy().astype(np.int64), onb)
    finally:
        gh.use_indptr32(old)
    # both paths give the same table over ALL sources (size-independent: path A == path B)
    all_src = torch.arange(0, n_tracks, device="cuda")
    a = nat.walk_topt(gh, all_src, 500, 0.85, 100, seed=3, want_i64=False, want_i32=True)
    gh.use_indptr32(False)
    b = nat.walk_topt(gh, all_src, 500, 0.85, 100, seed=3, want_i64=False, want_i32=True)
    gh.use_indptr32(True)
    assert torch.equal(a["nodes_i32"], b["nodes_i32"]) and torch.equal(a["weights_f32"], b["weights_f32"])


def test_graph_with_dead_end_is_rejected(nat):
    indptr = torch.tensor([0, 1, 1, 2])  # track 1 has no successors
    indices = torch.tensor([2, 0], dtype=torch.int32)
    with pytest.raises(nat.NativeError, match="no successors"):
        nat.GraphHandle(indptr, indices, 2, 1)


def test_empty_nodeset(nat, golden):
    g = golden("walk_dist")
    gh = _graph(nat, g["indptr"], g["indices"], int(g["n_tracks"]))
    out = nat.walk_topt(gh, torch.zeros(0, dtype=torch.int64), 500, 0.85, 10, seed=1)
    assert out["nodes"].shape == (0, 10)


def test_module_level_api(golden):
    """The reference-named entry points of pinsage_model (signatures and formats)."""
    import pinsage_model as psm
    from ps_graph import PSGraph
    g = golden("walk_topt")
    nt = int(g["a_n_tracks"])
    pg = PSGraph(g["a_indptr"], g["a_indices"], nt, len(g["a_indptr"]) - 1 - nt)
    nodeset = torch.from_numpy(g["a_nodeset"])
    trace = psm.do_random_walks(pg, nodeset, 500, 0.85, seed=3)
    assert trace.dtype == torch.int64 and trace.shape == (len(nodeset), 500) and not trace.is_cuda
    w, nb = psm.sample_neighborhood_topt(pg, nt, nodeset, 500, 0.85, 10, seed=3)
    ow, onb = oracle.topt_from_trace(trace.numpy(), nodeset.numpy(), 10)
    assert np.array_equal(w.numpy().view(np.int64), ow.view(np.int64)) and np.array_equal(nb.numpy(), onb)
    prob = psm.sample_neighborhood(pg, nt, nodeset, 500, 0.85, seed=3)
    assert prob.shape == (len(nodeset), pg.number_of_nodes()) and prob.dtype == torch.float64
    tw, tn = prob.topk(10, 1)
    assert torch.equal(tw, w)
    weights, nodes = psm.precompute_neighborhoods_topt(pg, nt, 500, 0.85, 100, None, seed=4)
    assert weights.shape == (nt, 100) and weights.dtype == torch.float64 and nodes.dtype == torch.int64


def test_sharded_precompute_equals_single_pass(tmp_path):
    """ps_dist.precompute_neighborhoods_sharded (rank r walks its node range, one all-gather): the shards of a 3-rank
    split, walked separately, tile exactly the table of a single pass with the same key (draws are keyed by
    (seed, source, step), never by launch shape), and the world-size-1 call returns the reference's tuple format
    (pinsage_model.py:109-132) equal to precompute_neighborhoods_topt's."""
    import ps_dist
    import ps_native
    import ps_synth
    import pinsage_model as psm
    n_tracks = 5000
    g = ps_synth.make_graph(n_tracks, 700, 60_000, seed=21)
    w1, n1 = psm.precompute_neighborhoods_topt(g, n_tracks, 300, 0.85, 40, None, seed=77)
    path = str(tmp_path / "nb.pt")
    w2, n2 = ps_dist.precompute_neighborhoods_sharded(g, n_tracks, 300, 0.85, 40, path, rank=0, world_size=1, seed=77)
    assert w2.dtype == torch.float64 and n2.dtype == torch.int64 and torch.equal(w1, w2) and torch.equal(n1, n2)
    assert torch.equal(w2._ps_table.nodes.cpu().long(), n1)
    lw, ln = torch.load(path)
    assert torch.equal(lw, w1) and torch.equal(ln, n1)
    w3, _ = ps_dist.precompute_neighborhoods_sharded(g, n_tracks, 300, 0.85, 40, path, rank=0, world_size=1, seed=5)  # cache hit
    assert torch.equal(w3, w1)
    parts = []
    for r in range(3):
        lo, hi = ps_dist.shard_range(n_tracks, r, 3)
        parts.append(ps_native.walk_topt(g.device(), torch.arange(lo, hi), 300, 0.85, 40, 77, want_i64=True))
    assert torch.equal(torch.cat([p["nodes"] for p in parts]).cpu(), n1)
    assert torch.equal(torch.cat([p["weights"] for p in parts]).cpu(), w1)
