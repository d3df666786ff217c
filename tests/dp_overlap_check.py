"""Run under torchrun on >= 2 GPUs (tests/test_gpu_dist.py launches it): data-parallel training with the
overlapped gradient exchange (ps_dist.GradSync: layers above layer 0 + head reduced while layer 0's backward runs)
gives the SAME parameters as the plain exchange (one allreduce after the step) and as the mean of the ranks' local
gradients applied by one Adam step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "gcn-song-embeddings_b200"), ROOT):
    sys.path.insert(0, p)
import tempfile

import torch
import torch.distributed as dist

import pinsage_training as pst
import ps_dist
import ps_synth


def build(rank, world, overlap):
    import pinsage_model as psm
    torch.manual_seed(1234)
    psm.seed_walker(0xD15C0)  # every build precomputes the same neighbourhood table
    g = ps_synth.make_graph(3000, 500, 40_000, seed=3, device="cuda")
    feats = ps_synth.features(3000, 128, seed=4, device="cuda")
    pos = ps_synth.cooccurrence_positives(g.device().indptr, g.device().indices, 3000, 20_000, seed=5)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix=f"dp{rank}_"))
    os.makedirs("runs", exist_ok=True)
    try:
        t = pst.PinSage(g, 3000, feats, pos, log=False, load_save=False)
    finally:
        os.chdir(cwd)
    t.batch_size = 96
    ps_dist.attach(t, rank, world)
    t.model.engine.want_upper_grads_event = overlap
    return t


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    results = []
    for overlap in (True, False):
        t = build(rank, world, overlap)
        gen = torch.Generator().manual_seed(100 + rank)
        batch = torch.randint(0, 3000, (96, 3), generator=gen)
        t.train_batch(batch)
        torch.cuda.synchronize()
        assert (t._grad_sync.comm is not None) == overlap
        grad = t.model.engine.flat_grad.clone()  # the summed gradient the optimiser step just consumed
        local = None
        if not overlap:  # this rank's own gradient of the same batch, for the sum check below
            t2 = build(rank, world, False)
            t2._grad_sync = None
            t2.train_batch(batch)
            torch.cuda.synchronize()
            local = t2.model.engine.flat_grad.clone()
            t2.close()
        # a few more steps keep running in this mode (events re-recorded every step, two steps in flight)
        for _ in range(3):
            loss = t.train_batch(torch.randint(0, 3000, (96, 3), generator=gen))[0]
        assert torch.isfinite(loss)
        results.append((grad, local))
        t.close()
    (g_overlap, _), (g_plain, local) = results
    scale = float(g_plain.abs().max())
    # split-K partial sums are combined with atomics, so two runs agree to rounding, not bit for bit
    assert float((g_overlap - g_plain).abs().max()) <= 2e-5 * scale, (float((g_overlap - g_plain).abs().max()), scale)
    total = local.clone()
    dist.all_reduce(total)
    assert float((total - g_plain).abs().max()) <= 2e-5 * scale  # the exchange sums the ranks' gradients
    assert float((local - g_plain).abs().max()) > 1e-3 * scale   # ... of DIFFERENT batches
    # node-range sharded inference: the fully sharded exchange path == the no-communication path == a plain embed of the shard
    t = build(rank, world, False)
    t.n = 3000
    lo, hi, e_x = ps_dist.embed_shard(t, exchange=True)
    _, _, e_c = ps_dist.embed_shard(t, exchange=False)
    ref = t.model.engine.embed(t._feats(), torch.arange(lo, hi, device="cuda"))
    assert e_x.shape == (hi - lo, 128)
    assert torch.equal(e_x, e_c), float((e_x - e_c).abs().max())
    assert torch.allclose(e_x, ref, rtol=1e-5, atol=1e-6), float((e_x - ref).abs().max())
    t.close()
    if rank == 0:
        print("dp overlap check ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
