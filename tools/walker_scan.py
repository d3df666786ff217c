"""Walker timing at several source counts (development tool): PS_LIB_PATH selects a library variant."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gcn-song-embeddings_b200"))
import torch
import ps_native as nat
import ps_synth
g = ps_synth.make_graph(1_000_000, 200_000, 40_000_000, seed=1234, device="cuda")
gh = g.device()
src = torch.arange(1_000_000, device="cuda")
perm = torch.randperm(1_000_000, device="cuda")
out = []
for T, n_hops in ((100, 500), (50, 500)):
    for name, s in (("first", src), ("random", perm)):
        for n in (100_000, 200_000, 500_000, 1_000_000):
            best = 1e9
            for rep in range(3):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                nat.walk_topt(gh, s[:n], n_hops, 0.85, T, seed=2 + rep, want_i64=False, want_i32=True)
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out.append(f"T={T} {name:6s} n={n:8d}: {best:7.3f} ms  {best / n * 1e3:6.2f} us/1k sources  {n * n_hops / best / 1e6:6.2f} G steps/s")
print(os.environ.get("PS_LIB_PATH", "default"))
print("\n".join(out))
