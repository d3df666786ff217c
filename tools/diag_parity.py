"""Development aid (GPU): prints the actual parity errors of the model path against the golden vectors of the
reference and against the CPU oracle at the bench's `micro` shape -- whole-tensor norm ratio AND max-abs-scaled
per-element error, per parameter.  Not a test; tests/test_gpu_model.py asserts the bars."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "gcn-song-embeddings_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np
import torch

from oracle import oracle
import pinsage_model as psm
import ps_native
import ps_synth


def errs(a, b):
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).double().cpu().reshape(-1)
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b).double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-300)), float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def golden_cfg(tag):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"model_{tag}.npz"))
    seed, n, dims, L, T = int(g["seed"]), int(g["n"]), tuple(int(x) for x in g["dims"]), int(g["L"]), int(g["T"])
    rng = np.random.RandomState(seed)
    features = torch.tensor(rng.standard_normal((n, dims[0])), dtype=torch.float32)
    params = oracle.make_params(L, dims, np.random.RandomState(seed + 1))
    nbhds = (torch.from_numpy(g["w"]), torch.from_numpy(g["nodes"]))
    m = psm.PinSageModel(None, n, L, dims, 500, 0.85, T, nbhds)
    m.load_state_dict(params)
    feats = m.engine.features(features)
    batch = torch.from_numpy(g["batch"]).cuda()
    full = tag != "default"
    for mtag, margin in (("m1e-5", 1e-5), ("m0.5", 0.5)):
        loss, emb, triples = m.engine.train_step(feats, batch, margin, True)
        print(f"[golden {tag} {mtag}] loss rel {abs(float(loss) - float(g[f'{mtag}/loss'])) / abs(float(g[f'{mtag}/loss'])):.2e}  emb {errs(emb[triples[:, 0].long()], g[f'{mtag}/hq'])}")
        for k, p in m.named_parameters():
            got = p.grad if full else p.grad.reshape(-1)[::97]
            print(f"    {k:28s} norm-rel {errs(got, g[f'{mtag}/grad/{k}'])[0]:.2e}  max-abs-scaled {errs(got, g[f'{mtag}/grad/{k}'])[1]:.2e}")


def micro(B=256, backend=0, margin=1e-5):
    ps_native.gemm_backend(backend)
    N, C, E, din, T, L = 20000, 4000, 400000, 256, 50, 2
    g = ps_synth.make_graph(N, C, E, seed=1234, device="cuda")
    feats = ps_synth.features(N, din, seed=1, device="cuda")
    out = ps_native.walk_topt(g.device(), torch.arange(N, device="cuda"), 500, 0.85, 100, seed=11)
    nbhds = (out["weights"].cpu(), out["nodes"].cpu())
    pos = ps_synth.cooccurrence_positives(g.indptr, g.indices, N, 200000, seed=2)
    rng = np.random.RandomState(3)
    pairs = pos[torch.from_numpy(rng.choice(pos.shape[0], B, replace=False))].numpy()
    batch = np.concatenate([pairs, rng.randint(0, N, size=(B, 1))], 1).astype(np.int64)
    dims = (din, 512, 128)
    params = oracle.make_params(L, dims, np.random.RandomState(0))
    m = psm.PinSageModel(g, N, L, dims, 500, 0.85, T, nbhds)
    m.load_state_dict(params)
    loss, emb, triples = m.engine.train_step(feats, torch.from_numpy(batch).cuda(), margin, True)
    o_loss, o_grads, (o_hq, o_hp, o_hn) = oracle.train_batch_grads(params, feats.cpu(), batch, nbhds, T, L, margin)
    print(f"[micro B={B} backend={backend} margin={margin}] loss {float(loss):.8e} oracle {float(o_loss):.8e} rel {abs(float(loss) - float(o_loss)) / abs(float(o_loss)):.2e}")
    print("    emb q", errs(emb[triples[:, 0].long()], o_hq), "neg", errs(emb[triples[:, 2].long()], o_hn))
    for k, p in m.named_parameters():
        e = errs(p.grad, o_grads[k])
        print(f"    {k:28s} norm-rel {e[0]:.2e}  max-abs-scaled {e[1]:.2e}  |g| {float(o_grads[k].norm()):.3e}")
    ps_native.gemm_backend(0)


if __name__ == "__main__":
    for tag in ("small", "l3", "default"):
        golden_cfg(tag)
    micro(256, 0, 1e-5)
    micro(256, 1, 1e-5)
    micro(256, 0, 0.1)
