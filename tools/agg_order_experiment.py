"""Does the ORDER of a layer's targets change aggregate_fwd's time (L2 reuse of the gathered Z rows)?  Development
experiment on the cfg3 shape: permute the rows of the layer-0 plan and time ps_aggregate_fwd.
Result (profiles/r2j_aggregate_fwd_locality.txt): no ordering changes the time by more than 3 %; a variant of the kernel
that loaded often-referenced rows with L2 evict_last and the rest with evict_first hints (createpolicy) moved the L2 hit
rate from 25 % to 30 % and the time not at all, and was not kept."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gcn-song-embeddings_b200"))
import torch
import ps_native as nat
import ps_synth
from ps_engine import NeighborTable, prepare_native

N, C, E, DIN, DH, T, B = 1_000_000, 200_000, 40_000_000, 256, 512, 50, 1024
g = ps_synth.make_graph(N, C, E, seed=1234, device="cuda")
gh = g.device()
out = nat.walk_topt(gh, torch.arange(N, device="cuda"), 500, 0.85, 100, seed=7, want_i64=False, want_i32=True)
table = NeighborTable.__new__(NeighborTable)
table.nodes, table.w, table.n, table.Tp, table.scratch = out["nodes_i32"], out["weights_f32"], N, 100, {}
torch.manual_seed(0)
batch = torch.randint(0, N, (B, 3), device="cuda")
plan, triples, counts = prepare_native(batch, 2, T, table)
lp, lp1 = plan.layers[0], plan.layers[1]
n, nz = lp.n, lp.nz
print(f"layer 0: {n} targets, {nz} z rows, {n * T} pairs; layer 1: {lp1.n} targets")
feats = torch.randn(N, DIN, device="cuda")
z = torch.randn(nz, DH, device="cuda")
cat = torch.empty(n, DIN + DH, device="cuda")
inv = torch.empty(n, device="cuda")
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")

def timed(self_rows, nbz, w, reps=5):
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nat.aggregate_fwd(feats, self_rows, DIN, z, nbz, w, DH, cat, inv)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

def run(name, perm):
    sr, nb, w = lp.self_rows[perm].contiguous(), lp.nbz[perm].contiguous(), lp.w[perm].contiguous()
    print(f"{name:40s} {timed(sr, nb, w):.3f} ms")

ident = torch.arange(n, device="cuda")
run("plan order (node id)", ident)
run("random", torch.randperm(n, device="cuda"))
run("by first (heaviest) neighbour row", torch.argsort(lp.nbz[:, 0].long(), stable=True))
run("by min neighbour row", torch.argsort(lp.nbz.min(1).values.long(), stable=True))
run("by median neighbour row", torch.argsort(lp.nbz.long().median(1).values, stable=True))
# by parent: layer-1 target whose neighbourhood lists this node first (layer-0 targets = frontier of layer 1)
pos_of = torch.full((n,), 1 << 40, dtype=torch.int64, device="cuda")
flat = lp1.nbz.reshape(-1).long()  # rows of the layer-1 input = positions in layer 0's target list
first = torch.arange(flat.numel(), device="cuda")
pos_of.scatter_reduce_(0, flat, first, reduce="amin")
run("by parent (first listing in layer 1)", torch.argsort(pos_of, stable=True))
# z rows renumbered by popularity (hot rows contiguous) does not change the order of targets; instead sort targets by
# their second-heaviest neighbour, and by a locality-sensitive key: the two heaviest neighbours
key2 = lp.nbz[:, 0].long() * nz + lp.nbz[:, 1].long()
run("by two heaviest neighbours", torch.argsort(key2, stable=True))
# how much reuse is there at all: distinct rows / pairs, and the share of pairs that hit the 30k most referenced rows
cntz = torch.bincount(lp.nbz.reshape(-1).long(), minlength=nz)
top = torch.sort(cntz, descending=True).values
for k in (10_000, 30_000, 60_000):
    print(f"pairs on the {k} most referenced rows: {float(top[:k].sum()) / (n * T):.3f}")
