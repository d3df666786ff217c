"""Development aid (GPU): the three big GEMM shapes of a cfg3 step timed alone with the L2 prefetch off / on."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gcn-song-embeddings_b200"))
import torch, ps_native as nat
torch.manual_seed(0)
N, nz, din, dh, do = 1_000_000, 656_000, 256, 512, 128
feats = torch.randn(N, din, device="cuda")
zrows = torch.sort(torch.randperm(N, device="cuda")[:nz]).values.to(torch.int32)
Qw = torch.randn(dh, din, device="cuda") * 0.05; Qb = torch.randn(dh, device="cuda")
Ww = torch.randn(do, din + dh, device="cuda") * 0.05
z = torch.empty(nz, dh, device="cuda"); mask = torch.empty(nz, dh // 32, dtype=torch.int32, device="cuda")
s_buf = torch.randn(nz, do, device="cuda")
gQ = torch.zeros(dh, din, device="cuda"); gb = torch.zeros(dh, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timed(fn, reps=5):
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts[2:])[len(ts[2:]) // 2]

calls = {
    "q_fwd_l0   [656k x 512 x 256] gather+mask": lambda: nat.gemm(feats, Qw, z, nz, dh, din, p_rows=zrows, bias=Qb, act=1, mask=mask),
    "agg_dgrad  [656k x 512 x 128] act2 mask  ": lambda: nat.gemm(s_buf, Ww[:, din:], z, nz, dh, do, q_kmajor=False, act=2, mask=mask),
    "q_wgrad_l0 [512 x 256 x 656k] gather     ": lambda: nat.gemm_wgrad(z, feats, gQ, dh, din, nz, x_rows=zrows, splits=74, bias_grad=gb),
}
import sys
dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
ref = {}
for cl in (1, 2, 3):
    nat.lib().ps_gemm_tc_cluster(cl)
    # correctness of the mode first: same shape, against mode 1 (same MMAs in the same order: bit-equal expected) and fp64
    zc = torch.empty(nz, dh, device="cuda"); mc = torch.empty(nz, dh // 32, dtype=torch.int32, device="cuda")
    nat.gemm(feats, Qw, zc, nz, dh, din, p_rows=zrows, bias=Qb, act=1, mask=mc); torch.cuda.synchronize()
    want = torch.nn.functional.leaky_relu(feats[zrows.long()][:4096].double() @ Qw.double().t() + Qb.double(), 0.01)
    err = float((zc[:4096].double() - want).abs().max() / want.abs().max())
    tail = torch.nn.functional.leaky_relu(feats[zrows.long()][-300:].double() @ Qw.double().t() + Qb.double(), 0.01)
    err_t = float((zc[-300:].double() - tail).abs().max() / tail.abs().max())
    if cl == 1:
        ref["z"], ref["m"] = zc.clone(), mc.clone()
    print("cluster mode", cl, "q_fwd max err vs fp64 (first 4096 rows / last 300 rows):", err, err_t,
          "| == mode 1:", bool(torch.equal(zc, ref["z"])) and bool(torch.equal(mc, ref["m"])), flush=True)
    gq = torch.zeros(dh, din, device="cuda"); gbb = torch.zeros(dh, device="cuda")
    nat.gemm_wgrad(zc, feats, gq, dh, din, nz, x_rows=zrows, splits=74, bias_grad=gbb); torch.cuda.synchronize()
    if cl == 1:
        ref["gq"], ref["gb"] = gq.clone(), gbb.clone()
        sub = slice(0, 200_000)
        want_g = zc[sub].double().t() @ feats[zrows[sub].long()].double()
        g_sub = torch.zeros(dh, din, device="cuda")
        nat.gemm_wgrad(zc[sub], feats, g_sub, dh, din, 200_000, x_rows=zrows[sub].contiguous(), splits=74)
        print("   wgrad (200k rows) max err vs fp64, scaled:", float((g_sub.double() - want_g).abs().max() / want_g.abs().max()))
    print("   wgrad vs mode 1: max |diff| / max |g| =", float((gq - ref["gq"]).abs().max() / ref["gq"].abs().max()),
          " bias grad:", float((gbb - ref["gb"]).abs().max() / ref["gb"].abs().max()), flush=True)
    print("cluster", cl, {k: round(timed(f), 4) for k, f in calls.items()}, flush=True)
    for k, f in calls.items():
        if "wgrad" in k:
            continue
        dbg.zero_(); nat.lib().ps_gemm_tc_trace(dbg.data_ptr()); f(); torch.cuda.synchronize(); nat.lib().ps_gemm_tc_trace(None)
        d = dbg.view(148, 8).double().mean(0).tolist()
        print("   ", k, "MMA-thread total %.0f kcyc | waits: operands %.0f%%, free TMEM %.0f%% | B-stream wait-empty %.0f%% | producer wait-empty %.0f%% | acc-warp wait-tfull %.0f%%, epilogue %.0f%% (of which the staged stores %.0f%%)"
              % (d[0] / 1e3, 100 * d[1] / d[0], 100 * d[2] / d[0], 100 * d[3] / d[0], 100 * d[4] / d[0], 100 * d[5] / d[0], 100 * d[6] / d[0], 100 * d[7] / d[0]))
lib = nat.lib()
for cl in ((1, 2) if '--ablate' in sys.argv else ()):
    lib.ps_gemm_tc_cluster(cl)
    for bits, what in ((0, "normal"), (1, "no A loads"), (2, "no B copies"), (4, "no stores"), (7, "MMA + drain only")):
        lib.ps_gemm_tc_experiment(bits)
        print("cluster", cl, "%-18s" % what, {k[:10]: round(timed(f), 4) for k, f in calls.items() if "wgrad" not in k}, flush=True)
    lib.ps_gemm_tc_experiment(0)
