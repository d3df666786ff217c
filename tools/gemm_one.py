"""Development aid (GPU): one launch of a cfg3-sized GEMM (for ncu).  usage: gemm_one.py q_fwd|agg_dgrad|q_wgrad [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gcn-song-embeddings_b200"))
import torch, ps_native as nat
which = sys.argv[1] if len(sys.argv) > 1 else "q_fwd"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
N, nz, din, dh, do = 1_000_000, 656_000, 256, 512, 128
feats = torch.randn(N, din, device="cuda")
zrows = torch.sort(torch.randperm(N, device="cuda")[:nz]).values.to(torch.int32)
Qw = torch.randn(dh, din, device="cuda") * 0.05; Qb = torch.randn(dh, device="cuda")
Ww = torch.randn(do, din + dh, device="cuda") * 0.05
z = torch.empty(nz, dh, device="cuda"); mask = torch.zeros(nz, dh // 32, dtype=torch.int32, device="cuda")
s_buf = torch.randn(nz, do, device="cuda")
gQ = torch.zeros(dh, din, device="cuda"); gb = torch.zeros(dh, device="cuda")
for _ in range(reps):
    if which == "q_fwd":
        nat.gemm(feats, Qw, z, nz, dh, din, p_rows=zrows, bias=Qb, act=1, mask=mask)
    elif which == "agg_dgrad":
        nat.gemm(s_buf, Ww[:, din:], z, nz, dh, do, q_kmajor=False, act=2, mask=mask)
    else:
        nat.gemm_wgrad(z, feats, gQ, dh, din, nz, x_rows=zrows, splits=74, bias_grad=gb)
torch.cuda.synchronize()
print("ok")
