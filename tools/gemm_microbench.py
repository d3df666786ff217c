"""GEMM microbenchmark for the shapes of the cfg3 train step (development tool, not a bench line).
usage: python tools/gemm_microbench.py [--backend 0|1] [--iters N]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gcn-song-embeddings_b200"))
import torch
import ps_native as nat

ap = argparse.ArgumentParser(); ap.add_argument("--backend", type=int, default=0); ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--only", default=""); ap.add_argument("--pack", type=int, default=1)
args = ap.parse_args()
nat._ensure_device(); nat.gemm_backend(args.backend); nat.gemm_tc_pack(args.pack)
torch.manual_seed(0)
N_TAB = 1_000_000
feat = torch.randn(N_TAB, 256, device="cuda")
shapes = {
    # name: (M, N, K, pk, qk, gather, accumulate)
    "q_fwd_l0   [656k x 512 x 256] gather": (655_360, 512, 256, True, True, True, False),
    "q_fwd_l0   [656k x 512 x 256] dense ": (655_360, 512, 256, True, True, False, False),
    "w_fwd_l0   [96k x 128 x 768]": (96_000, 128, 768, True, True, False, False),
    "w_dgrad_l0 [96k x 768 x 128]": (96_000, 768, 128, True, False, False, False),
    "q_wgrad_l0 [512 x 256 x 656k] gather": (512, 256, 655_360, False, False, True, True),
    "w_wgrad_l0 [128 x 768 x 96k]": (128, 768, 96_000, False, False, False, True),
}
for name, (M, N, K, pk, qk, gather, acc) in shapes.items():
    if args.only and args.only not in name:
        continue
    if acc:  # wgrad: P = dY [K rows, M], Q = X [rows, N]
        P = torch.randn(K, M, device="cuda"); Q = feat if gather else torch.randn(K, N, device="cuda")
        rows = torch.randint(0, N_TAB, (K,), device="cuda", dtype=torch.int32) if gather else None
        C = torch.zeros(M, N, device="cuda")
        run = lambda: nat.gemm(P, Q, C, M, N, K, p_kmajor=False, q_kmajor=False, q_rows=rows, accumulate=True, splits=max(1, 592 // (-(-M // 128) * -(-N // 128))))
    else:
        P = feat if gather else torch.randn(M, K, device="cuda")
        rows = torch.randint(0, N_TAB, (M,), device="cuda", dtype=torch.int32) if gather else None
        Q = torch.randn(N, K, device="cuda") if qk else torch.randn(K, N, device="cuda")
        C = torch.empty(M, N, device="cuda")
        run = lambda: nat.gemm(P, Q, C, M, N, K, p_kmajor=True, q_kmajor=qk, p_rows=rows)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    print(f"backend {args.backend} pack {args.pack}  {name:42s} {ms:8.3f} ms  {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s", flush=True)
