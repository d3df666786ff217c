"""Development aid (GPU): compare layer-0 backward intermediates between the tcgen05 and CUDA-core back-ends."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "gcn-song-embeddings_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np, torch
from oracle import oracle
import pinsage_model as psm, ps_native as nat, ps_synth

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

orig_gemm = nat.gemm
cap = {}
def gemm(P, Q, C, M, N_, K, **kw):
    tag = kw.get("tag", "gemm")
    if tag == "gemm_agg_dgrad_l0":
        cap["s_buf"] = P.clone(); cap["z_before"] = C.clone()
        mask = kw.get("mask")
        if mask is not None:
            bits = ((mask.view(M, N_ // 32, 1) >> torch.arange(32, device="cuda", dtype=torch.int32)) & 1).reshape(M, N_).bool()
            cap["mask_mismatch"] = int((bits != (C > 0)).sum()); cap["z_zero"] = int((C == 0).sum())
    orig_gemm(P, Q, C, M, N_, K, **kw)
    if tag == "gemm_agg_dgrad_l0":
        cap["dz"] = C.clone()
    if tag == "gemm_q_fwd_l0":
        cap["z_fwd"] = C.clone()
nat.gemm = gemm
import ps_engine
res = {}
for backend in (0, 1, 0):
    nat.gemm_backend(backend)
    cap.clear()
    m.engine.train_step(feats, torch.from_numpy(batch).cuda(), 1e-5, True)
    torch.cuda.synchronize()
    res.setdefault(backend, []).append({**{k: v for k, v in cap.items()}, "gq": m.conv_layers[0].Q.weight.grad.clone(), "gb": m.conv_layers[0].Q.bias.grad.clone()})
    print(backend, "mask_mismatch", cap.get("mask_mismatch"), "z_zero", cap.get("z_zero"))
a, b, a2 = res[0][0], res[1][0], res[0][1]
def rel(x, y): return float((x.double() - y.double()).norm() / y.double().norm())
for k in ("z_fwd", "z_before", "s_buf", "dz", "gq", "gb"):
    print(k, "tc vs simt", rel(a[k], b[k]), " tc run1 vs run2", rel(a[k], a2[k]))
d = (a["dz"].double() - b["dz"].double()).abs()
print("dz diff: max", float(d.max()), "ref max", float(b["dz"].abs().max()), "rows with diff>1e-3*max:", int((d.max(1).values > 1e-3 * b["dz"].abs().max()).sum()), "of", d.shape[0])
rows = (d.max(1).values > 1e-3 * b["dz"].abs().max()).nonzero().flatten()[:20]
print("bad rows", rows.tolist())
for r in rows[:5].tolist():
    bad_cols = (d[r] > 1e-3 * b["dz"].abs().max()).nonzero().flatten()
    print("row", r, "n bad cols", bad_cols.numel(), "cols", bad_cols[:16].tolist(), "tc", a["dz"][r, bad_cols[:4]].tolist(), "simt", b["dz"][r, bad_cols[:4]].tolist(),
          "z_before tc", a["z_before"][r, bad_cols[:4]].tolist())
ds = (a["s_buf"].double() - b["s_buf"].double()).abs()
print("s_buf diff max", float(ds.max()), "ref max", float(b["s_buf"].abs().max()))
