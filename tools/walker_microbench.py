"""Walker microbenchmark (BASELINE.json configs[4]): random-walk sampler on the cfg3 graph
(1 M tracks / 200 k playlists / 40 M edges), visit-count top-T per source.
  * reference-exact mode: n_hops = 500, alpha = 0.85 (geometric segments), T = 100
  * fixed-length mode: 100 walks of L in {3,4,5} steps per source (1e8 walks), T = 50
usage: python tools/walker_microbench.py [--sources N] [--json out.jsonl]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gcn-song-embeddings_b200"))
import torch
import ps_native as nat
import ps_synth

ap = argparse.ArgumentParser()
ap.add_argument("--sources", type=int, default=1_000_000)
ap.add_argument("--json", default=None)
args = ap.parse_args()
g = ps_synth.make_graph(1_000_000, 200_000, 40_000_000, seed=1234, device="cuda")
gh = g.device()
src = torch.arange(args.sources, device="cuda")
peak = 6548.5
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
modes = [("alpha0.85_hops500_T100", 500, 0.85, 0, 100)] + [(f"fixed{L}_walks100_T50", 100 * L, 0.85, L, 50) for L in (3, 4, 5)]
lines = []
for name, n_hops, alpha, fixed_len, T in modes:
    nat.walk_topt(gh, src[: args.sources // 8], n_hops, alpha, T, seed=1, fixed_len=fixed_len, want_i64=False, want_i32=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nat.walk_topt(gh, src, n_hops, alpha, T, seed=2, fixed_len=fixed_len, want_i64=False, want_i32=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps = args.sources * n_hops
    line = {"mode": name, "sources": args.sources, "n_hops": n_hops, "walks": args.sources * (n_hops // fixed_len) if fixed_len else None,
            "ms": round(ms, 3), "steps_per_s": round(steps / ms * 1e3, 1), "algorithmic_gbs": round(steps * 28 / ms / 1e6, 1),
            "frac_of_hbm_algorithmic": round(steps * 28 / ms / 1e6 / peak, 4)}
    lines.append(line)
    print(json.dumps(line), flush=True)
if args.json:
    with open(args.json, "w") as f:
        f.write("\n".join(json.dumps(l) for l in lines) + "\n")
