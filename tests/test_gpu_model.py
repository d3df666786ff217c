"""GPU parity of the model path (ConvLayer, PinSageModel forward/backward, fused train
step, trainer) against golden vectors produced by the unmodified reference and against
the CPU oracle on fresh random inputs.  Tolerance: 1e-4 relative (north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def rel(a, b):
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).double().cpu()
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _inputs(g):
    seed, n, dims, L, T = int(g["seed"]), int(g["n"]), tuple(int(x) for x in g["dims"]), int(g["L"]), int(g["T"])
    rng = np.random.RandomState(seed)
    features = torch.tensor(rng.standard_normal((n, dims[0])), dtype=torch.float32)
    params = oracle.make_params(L, dims, np.random.RandomState(seed + 1))
    nbhds = (torch.from_numpy(g["w"]), torch.from_numpy(g["nodes"]))
    return features, params, nbhds, n, L, T, dims


def _model(n, L, dims, T, nbhds, params):
    import pinsage_model as psm
    m = psm.PinSageModel(None, n, L, dims, 500, 0.85, T, nbhds)
    m.load_state_dict(params)
    return m


@pytest.mark.parametrize("tag", ["small", "l3", "default"])
def test_forward_and_autograd_vs_reference(golden, tag):
    import pinsage_model as psm
    g = golden(f"model_{tag}")
    features, params, nbhds, n, L, T, dims = _inputs(g)
    full = tag != "default"
    model = _model(n, L, dims, T, nbhds, params)
    batch = torch.from_numpy(g["batch"])
    # ConvLayer.forward with the reference's signature
    S = psm.relevant_nodes_per_layer_precomp(batch[:, 0], L, T, nbhds)
    ns0, w0, nb0 = S[0]
    c0 = model.conv_layers[0](features, ns0, nb0, w0)
    assert not c0.is_cuda and rel(c0, g["conv0_out"]) < RTOL
    # one call + linear functional: embeddings and the duplicate-node gradient factor
    emb = model(features, batch[:, 0])
    assert not emb.is_cuda and emb.shape == (batch.shape[0], dims[2])
    assert rel(emb.detach(), g["emb_q"]) < RTOL
    (emb * torch.from_numpy(g["R"])).sum().backward()
    for k, p in model.named_parameters():
        got = p.grad if full else p.grad.reshape(-1)[::97]
        assert rel(got, g[f"lin_grad/{k}"]) < RTOL, k
    # three separate calls + torch loss through autograd (exactly the reference's train_batch graph)
    import pinsage_training as pst
    for mtag, margin in (("m1e-5", 1e-5), ("m0.5", 0.5)):
        model.zero_grad()
        hq, hp, hn = model(features, batch[:, 0]), model(features, batch[:, 1]), model(features, batch[:, 2])
        loss = pst.max_margin_loss(hq, hp, hn, margin)
        loss.backward()
        assert abs(float(loss) - float(g[f"{mtag}/loss"])) < RTOL * abs(float(g[f"{mtag}/loss"])) + 1e-9
        assert rel(hp.detach(), g[f"{mtag}/hp"]) < RTOL
        for k, p in model.named_parameters():
            got = p.grad if full else p.grad.reshape(-1)[::97]
            assert rel(got, g[f"{mtag}/grad/{k}"]) < 5 * RTOL, (mtag, k)


@pytest.mark.parametrize("tag", ["small", "l3", "default"])
def test_fused_train_step_vs_reference(golden, tag):
    """Engine.train_step (shared frontier + CUDA loss/backward) reproduces the reference's
    loss and parameter gradients, duplicate factor included."""
    g = golden(f"model_{tag}")
    features, params, nbhds, n, L, T, dims = _inputs(g)
    full = tag != "default"
    model = _model(n, L, dims, T, nbhds, params)
    feats = model.engine.features(features)
    batch = torch.from_numpy(g["batch"]).cuda()
    for mtag, margin in (("m1e-5", 1e-5), ("m0.5", 0.5)):
        loss, emb, triples = model.engine.train_step(feats, batch, margin, reference_compat=True)
        assert abs(float(loss) - float(g[f"{mtag}/loss"])) < RTOL * abs(float(g[f"{mtag}/loss"])) + 1e-9
        assert rel(emb[triples[:, 0].long()], g[f"{mtag}/hq"]) < RTOL
        assert rel(emb[triples[:, 2].long()], g[f"{mtag}/hn"]) < RTOL
        for k, p in model.named_parameters():
            got = p.grad if full else p.grad.reshape(-1)[::97]
            assert rel(got, g[f"{mtag}/grad/{k}"]) < 5 * RTOL, (mtag, k)


def test_trainer_steps_vs_reference(golden, tmp_path, monkeypatch):
    """Three optimiser steps of the drop-in trainer == the reference trainer's."""
    import pinsage_training as pst
    from ps_graph import PSGraph
    g = golden("train_steps")
    n, din = int(g["n"]), int(g["din"])
    rng = np.random.RandomState(21)
    features = torch.tensor(rng.standard_normal((n, din)), dtype=torch.float32)
    params = oracle.make_params(2, (din, 512, 128), np.random.RandomState(22))
    monkeypatch.chdir(tmp_path)
    nb_path = str(tmp_path / "neighborhoods.pt")
    torch.save((torch.from_numpy(g["w"]), torch.from_numpy(g["nodes"])), nb_path)
    graph = PSGraph(np.arange(n + 2), np.r_[np.full(n, n), 0].astype(np.int32), n, 1, nbhds_path=nb_path)
    positives = torch.from_numpy(rng.randint(0, n, size=(500, 2)).astype(np.int64))
    trainer = pst.PinSage(graph, n, features, positives, log=False, load_save=False)
    trainer.model.load_state_dict(params)
    losses = [float(trainer.train_batch(torch.from_numpy(b))[0]) for b in g["batches"]]
    assert np.allclose(losses, g["losses"], rtol=2e-3, atol=1e-8)
    for k, p in trainer.model.named_parameters():
        assert rel(p.detach().reshape(-1)[::53], g[f"param_sub/{k}"]) < RTOL, k
    emb = trainer.embed(torch.arange(0, 40))
    assert not emb.is_cuda and rel(emb, g["emb_after"]) < 1e-3
    # checkpoint format round trip (state.pt keys of pinsage_training.py:288-295)
    trainer.save_model()
    prog = torch.load(os.path.join(pst.BASE_RUN_DIR, trainer.run_name, "state.pt"))
    assert set(prog) == {"epochs_done", "batches_done", "model_state", "optimizer_state"}
    assert set(prog["model_state"]) == set(params)


def test_frontier_api_vs_reference(golden):
    import pinsage_model as psm
    g = golden("frontier")
    nbhds = (torch.from_numpy(g["w"]), torch.from_numpy(g["nodes"]))
    for tag, (L, T) in {"L2T3": (2, 3), "L3T5": (3, 5), "L2T10": (2, 10)}.items():
        S = psm.relevant_nodes_per_layer_precomp(torch.from_numpy(g[f"{tag}_nodeset"]), L, T, nbhds)
        for l, (ns, w, nb) in enumerate(S):
            assert np.array_equal(ns.numpy(), g[f"{tag}_ns{l}"]) and np.array_equal(nb.numpy(), g[f"{tag}_nb{l}"])
            assert np.array_equal(w.numpy().view(np.int64), g[f"{tag}_w{l}"].view(np.int64))


@pytest.mark.parametrize("n,dims,L,T,B", [(3000, (256, 512, 128), 2, 10, 200), (1500, (128, 256, 64), 3, 4, 64),
                                         (800, (512, 1024, 512), 1, 6, 50)])
def test_against_oracle_random(n, dims, L, T, B):
    """Fresh random inputs at larger sizes: embeddings, loss and gradients vs the CPU oracle."""
    rng = np.random.RandomState(n)
    features = torch.tensor(rng.standard_normal((n, dims[0])), dtype=torch.float32)
    nodes = np.stack([rng.choice(n, size=max(T, 8), replace=False) for _ in range(n)]).astype(np.int64)
    w = (np.sort(rng.randint(1, 60, size=nodes.shape), axis=1)[:, ::-1] / 500.0).copy()
    nbhds = (torch.from_numpy(w), torch.from_numpy(nodes))
    params = oracle.make_params(L, dims, np.random.RandomState(n + 1))
    batch = rng.randint(0, n, size=(B, 3)).astype(np.int64)
    batch[1] = batch[0]  # a fully duplicated triple
    o_loss, o_grads, (o_hq, _, o_hn) = oracle.train_batch_grads(params, features, batch, nbhds, T, L, 0.05)
    model = _model(n, L, dims, T, nbhds, params)
    feats = model.engine.features(features)
    loss, emb, triples = model.engine.train_step(feats, torch.from_numpy(batch).cuda(), 0.05, True)
    assert abs(float(loss) - float(o_loss)) < RTOL * abs(float(o_loss)) + 1e-9
    assert rel(emb[triples[:, 0].long()], o_hq) < RTOL and rel(emb[triples[:, 2].long()], o_hn) < RTOL
    for k, p in model.named_parameters():
        assert rel(p.grad, o_grads[k]) < 5 * RTOL, k
    # inference path == training forward
    out = model.engine.embed(feats, torch.from_numpy(batch[:, 0]).cuda())
    assert rel(out, o_hq) < RTOL
    # without the compat factor the gradient is the plain autograd one: differs when duplicates exist
    model.engine.train_step(feats, torch.from_numpy(batch).cuda(), 0.05, False)
    assert rel(model.G2.weight.grad, o_grads["G2.weight"]) > 1e-3


def test_state_dict_roundtrip_with_reference_keys(golden):
    g = golden("model_small")
    features, params, nbhds, n, L, T, dims = _inputs(g)
    model = _model(n, L, dims, T, nbhds, params)
    sd = model.state_dict()
    assert list(sd) == list(params)
    for k in params:
        assert torch.equal(sd[k].cpu(), params[k])


def test_embed_range_equals_frontier_embed():
    """Layer-wise node-range inference (Engine.embed_range, the shard of BASELINE.json configs[3]) gives the same
    embeddings as the per-batch frontier path (PinSage.embed), for 2 and 3 layers, ragged ranges and tiny chunks."""
    import ps_synth
    import pinsage_model as psm
    from oracle import oracle
    n_tracks = 900
    g = ps_synth.make_graph(n_tracks, 120, 9000, seed=5)
    feats = ps_synth.features(n_tracks, 64, seed=6).cuda()
    nb = psm.sample_neighborhood_topt(g, n_tracks, torch.arange(n_tracks), 300, 0.85, 12, seed=3)
    for L, T, dims in ((2, 5, (64, 96, 32)), (3, 4, (64, 48, 64))):
        model = psm.PinSageModel(g, n_tracks, L, dims, 300, 0.85, T, nb)
        model.load_state_dict(oracle.make_params(L, dims, np.random.RandomState(L)))
        want = model.engine.embed(feats, torch.arange(n_tracks, device="cuda"))
        for lo, hi, chunk in ((0, n_tracks, 1 << 18), (0, n_tracks, 97), (123, 457, 64), (899, 900, 8), (10, 10, 8)):
            st = {}
            got = model.engine.embed_range(feats, lo, hi, chunk=chunk, stats=st)
            assert got.shape == (hi - lo, dims[2])
            if hi > lo:
                assert torch.allclose(got, want[lo:hi], rtol=1e-5, atol=1e-6), float((got - want[lo:hi]).abs().max())


def test_online_sampling_matches_table_given_same_walks():
    """Online neighbourhoods (walker inside the forward, pinsage_model.py:142-154) drive the same engine: a table
    precomputed with the Philox key of a plan reproduces that plan's online forward bit for bit (walks are keyed by
    (seed, source, step), not by launch shape), and a train step runs end to end in online mode."""
    import ps_native as nat
    import ps_synth
    import pinsage_model as psm
    from ps_engine import NeighborTable, build_plan
    from oracle import oracle
    n_tracks, dims, L, T = 700, (64, 96, 32), 2, 6
    g = ps_synth.make_graph(n_tracks, 100, 7000, seed=8)
    feats = ps_synth.features(n_tracks, 64, seed=9).cuda()
    model = psm.PinSageModel(g, n_tracks, L, dims, 200, 0.85, T, None)   # nbhds=None -> online
    model.load_state_dict(oracle.make_params(L, dims, np.random.RandomState(3)))
    online = model.nbhds
    top = torch.arange(0, n_tracks, 9, device="cuda")
    plan = build_plan(top, L, T, online, need_backward=False)
    out_online, _ = model.engine.forward(feats, plan, keep=False)
    ref = nat.walk_topt(g.device(), torch.arange(n_tracks), 200, 0.85, T, online._seed, want_i64=False, want_i32=True)
    table = NeighborTable.__new__(NeighborTable)
    table.nodes, table.w, table.n, table.Tp, table.scratch = ref["nodes_i32"], ref["weights_f32"], n_tracks, T, {}
    plan2 = build_plan(top, L, T, table, need_backward=False)
    out_table, _ = model.engine.forward(feats, plan2, keep=False)
    assert torch.equal(out_online, out_table) and torch.isfinite(out_online).all()
    seed_before = online._seed
    batch = torch.randint(0, n_tracks, (32, 3), device="cuda")
    loss, emb, triples = model.engine.train_step(feats, batch, 0.1, True)
    assert online._seed != seed_before  # fresh walks per step
    assert torch.isfinite(loss).all() and all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


@pytest.mark.parametrize("n_tracks,T,L,ntop", [(900, 5, 2, 100), (5000, 12, 3, 700), (300, 3, 2, 1)])
def test_native_plan_builder_equals_torch_path(n_tracks, T, L, ntop):
    """csrc/plan.cu (ps_plan_layer / ps_plan_transpose) builds exactly the plan of the framework-op path (which the
    CPU tests pin against the reference's relevant_nodes_per_layer_precomp): frontiers, positions, the stable pair
    order, segment offsets, work chunks and the chunk -> row map."""
    from ps_engine import NeighborTable, build_plan
    rng = np.random.RandomState(n_tracks)
    nodes = torch.from_numpy(np.stack([rng.choice(n_tracks, size=T + 2, replace=False) for _ in range(n_tracks)]).astype(np.int64))
    w = torch.from_numpy(rng.randint(1, 40, size=(n_tracks, T + 2)).astype(np.float64) / 500)
    top = torch.from_numpy(np.sort(rng.choice(n_tracks, size=ntop, replace=False)).astype(np.int64))
    dev = build_plan(top.cuda(), L, T, NeighborTable(w, nodes, device="cuda"), need_backward=True)
    cpu = build_plan(top, L, T, NeighborTable(w, nodes, device="cpu"), need_backward=True)
    for a, b in zip(dev.layers, cpu.layers):
        assert (a.n, a.nz) == (b.n, b.nz)
        for name in ("self_rows", "nbz", "w", "zrows", "seg_off", "pair_q", "chunk_off"):
            x, y = getattr(a, name), getattr(b, name)
            assert (x is None) == (y is None), name
            if x is not None:
                assert torch.equal(x.cpu().to(y.dtype), y), name
        total = int(b.chunk_off[-1])
        assert torch.equal(a.chunk_row.cpu()[:total], b.chunk_row[:total])
