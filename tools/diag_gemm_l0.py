"""Development aid (GPU): every GEMM of one micro-shape train step checked against an fp64 torch product of ITS OWN
inputs (captured at call time), to find which call loses accuracy on real data."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "gcn-song-embeddings_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np
import torch

from oracle import oracle
import pinsage_model as psm
import ps_native as nat
import ps_synth

N, C, E, din, T, L, B = 20000, 4000, 400000, 256, 50, 2, 256
g = ps_synth.make_graph(N, C, E, seed=1234, device="cuda")
feats = ps_synth.features(N, din, seed=1, device="cuda")
out = nat.walk_topt(g.device(), torch.arange(N, device="cuda"), 500, 0.85, 100, seed=11)
nbhds = (out["weights"].cpu(), out["nodes"].cpu())
pos = ps_synth.cooccurrence_positives(g.indptr, g.indices, N, 200000, seed=2)
rng = np.random.RandomState(3)
pairs = pos[torch.from_numpy(rng.choice(pos.shape[0], B, replace=False))].numpy()
batch = np.concatenate([pairs, rng.randint(0, N, size=(B, 1))], 1).astype(np.int64)
dims = (din, 512, 128)
m = psm.PinSageModel(g, N, L, dims, 500, 0.85, T, nbhds)
m.load_state_dict(oracle.make_params(L, dims, np.random.RandomState(0)))

orig = nat._gemm
log = []


def spy(P, Q, C, M, N_, K, p_kmajor, q_kmajor, p_rows, q_rows, bias, act, l2norm, norm_out, accumulate, splits, mask=None):
    A = P[p_rows.long()] if (p_rows is not None and p_kmajor) else P
    A = A[:M, :K] if p_kmajor else (A[p_rows.long()] if p_rows is not None else A)[:K, :M].t()
    Bm = Q[:N_, :K] if q_kmajor else (Q[q_rows.long()] if q_rows is not None else Q)[:K, :N_].t()
    want = A.double() @ Bm.double().t()
    absprod = A.double().abs() @ Bm.double().abs().t()
    if bias is not None:
        want = want + bias.double()
    before = C.clone() if (accumulate or (act == 2 and mask is None)) else None
    orig(P, Q, C, M, N_, K, p_kmajor, q_kmajor, p_rows, q_rows, bias, act, l2norm, norm_out, accumulate, splits, mask)
    got = C.double()
    if accumulate:
        got = got - before.double()
    elif act == 1:
        want = torch.nn.functional.leaky_relu(want, 0.01)
        if l2norm:
            want = want / want.norm(dim=1, keepdim=True)
    elif act == 2:
        if mask is not None:
            bits = ((mask.view(M, N_ // 32, 1) >> torch.arange(32, device="cuda", dtype=torch.int32)) & 1).reshape(M, N_).bool()
        else:
            bits = before > 0
        want = want * torch.where(bits, 1.0, 0.01).double()
    err = (got - want)
    log.append((spy.tag, M, N_, K, splits, float(err.norm() / want.norm()), float(err.abs().max() / want.abs().max()),
                float((err.abs() / absprod.clamp_min(1e-300)).max()), float(want.abs().mean() / absprod.mean())))


def gemm(P, Q, C, M, N_, K, *, p_kmajor=True, q_kmajor=True, p_rows=None, q_rows=None, bias=None, act=0, l2norm=False,
         norm_out=None, accumulate=False, splits=1, mask=None, tag="gemm"):
    spy.tag = tag
    spy(P, Q, C, M, N_, K, p_kmajor, q_kmajor, p_rows, q_rows, bias, act, l2norm, norm_out, accumulate, splits, mask)


nat.gemm = gemm
for backend in (0, 1):
    nat.gemm_backend(backend)
    log.clear()
    m.engine.train_step(feats, torch.from_numpy(batch).cuda(), 1e-5, True)
    torch.cuda.synchronize()
    print(f"== backend {backend}: tag M N K splits | norm-rel | max-abs-scaled | max err/(|A||B|) | cancellation mean|C|/mean(|A||B|)")
    for r in log:
        print(f"   {r[0]:22s} {r[1]:6d} {r[2]:4d} {r[3]:6d} {r[4]:3d} | {r[5]:.2e} | {r[6]:.2e} | {r[7]:.2e} | {r[8]:.2e}")
