"""kNN microbenchmark (development tool): one 1024-query tile against N embeddings, k = 1000 (eval.PRECOMP_K).
usage: python tools/knn_microbench.py [--n 1000000]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gcn-song-embeddings_b200"))
import torch
import ps_native as nat
import ps_knn

ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=1_000_000); ap.add_argument("--k", type=int, default=1000)
args = ap.parse_args()
torch.manual_seed(0)
emb = torch.randn(args.n, 128, device="cuda")
q = torch.randint(0, args.n, (1024,), device="cuda")

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

sim = ps_knn.cosine_sim_ab(emb[q].contiguous(), emb)
en = (emb / emb.norm(dim=1, keepdim=True)).contiguous(); qe = en[q].contiguous()
out = torch.empty((1024, args.n), device="cuda")
for pack in (1, 0):
    nat.gemm_tc_pack(pack)
    print(f"N={args.n}: similarity GEMM [1024 x N x 128], weight packing {pack}: {timed(lambda: nat.gemm(qe, en, out, 1024, args.n, 128)):.2f} ms")
nat.gemm_tc_pack(1)
print(f"  ps_topk_rows {timed(lambda: nat.topk_rows(sim, args.k + 1)):.2f} ms   torch.topk {timed(lambda: sim.topk(args.k + 1, dim=1)):.2f} ms")
print(f"  knn_from_emb (tile + ps_topk_rows) {timed(lambda: ps_knn.knn_from_emb(emb, q, args.k, fused=False)):.2f} ms per 1024 queries")
ps_knn.fused_stats.update(tiles=0, fallback_tiles=0)
print(f"  knn_from_emb (fused: sample -> ps_gemm_filter -> ps_topk_rows_mapped) {timed(lambda: ps_knn.knn_from_emb(emb, q, args.k)):.2f} ms per 1024 queries   {ps_knn.fused_stats}")
k1 = args.k + 1
thr = nat.topk_rows(sim[:, ::17].contiguous()[:, : (sim[:, ::17].shape[1] // 4) * 4], 128)[0][:, -1].contiguous()
cap = 4660
print(f"  ps_gemm_filter alone {timed(lambda: nat.gemm_filter(en, qe, thr, cap)):.2f} ms")
cnt, val, row = nat.gemm_filter(en, qe, thr, cap)
print(f"  candidates per query: min {int(cnt.min())} mean {float(cnt.float().mean()):.0f} max {int(cnt.max())};  ps_topk_rows_mapped {timed(lambda: nat.topk_rows_mapped(val, row, cnt, k1)):.3f} ms")
