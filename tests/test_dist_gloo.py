"""world_size-2 `gloo` tests (CPU) of the data-parallel host logic in ps_dist."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "gcn-song-embeddings_b200"))
    import ps_dist
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = ps_dist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    # gradient mean-allreduce on a flat buffer
    flat = torch.full((1000,), float(rank + 1))
    ps_dist.allreduce_mean_(flat, world)
    assert torch.allclose(flat, torch.full((1000,), 1.5))
    # parameter broadcast from rank 0
    lin = torch.nn.Linear(4, 3)
    with torch.no_grad():
        lin.weight.fill_(float(rank)); lin.bias.fill_(float(rank))
    ps_dist.broadcast_parameters(lin)
    assert float(lin.weight.abs().sum()) == 0.0
    # a replica pair stays in lock-step: same params + averaged grads -> same Adam update
    torch.manual_seed(0)
    model = torch.nn.Linear(8, 2)
    ps_dist.broadcast_parameters(model)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    torch.manual_seed(100 + rank)  # rank-seeded batches differ
    x = torch.randn(16, 8)
    model(x).pow(2).mean().backward()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    ps_dist.allreduce_mean_(flat, world)
    off = 0
    for p in model.parameters():
        p.grad.copy_(flat[off: off + p.numel()].view_as(p)); off += p.numel()
    opt.step()
    gathered = [torch.zeros_like(model.weight) for _ in range(world)]
    dist.all_gather(gathered, model.weight.data)
    assert torch.equal(gathered[0], gathered[1])
    # GradSync (the object attach() installs) without a recorded event: one plain sum over the whole flat buffer, the mean's
    # 1 / world folded into the optimiser (grad_scale) when it offers that, else divided here
    class _Obj:
        pass
    for has_scale in (True, False):
        tr, eng, optim = _Obj(), _Obj(), _Obj()
        eng.flat_grad = torch.full((64,), float(rank + 1))
        tr.model = _Obj(); tr.model.engine = eng; tr.optimizer = optim
        if has_scale:
            optim.grad_scale = 1.0
        sync = ps_dist.GradSync(tr, world)
        assert eng.want_upper_grads_event is True
        sync()
        if has_scale:
            assert optim.grad_scale == 1.0 / world and torch.equal(eng.flat_grad, torch.full((64,), 3.0))
        else:
            assert torch.equal(eng.flat_grad, torch.full((64,), 1.5))
    # max-over-ranks timing and node-range shards
    assert ps_dist.max_over_ranks(float(rank)) == float(world - 1)
    lo, hi = ps_dist.shard_range(1001, rank, world)
    sizes = torch.tensor([hi - lo]); dist.all_reduce(sizes)
    assert int(sizes) == 1001 and (lo == 0 if rank == 0 else lo == 501)
    # layer-output exchange of the sharded inference: even and ragged shards
    for n in (1000, 1001, 7):
        lo, hi = ps_dist.shard_range(n, rank, world)
        want = torch.arange(n * 3, dtype=torch.float32).view(n, 3)
        table = torch.full((n, 3), -1.0)
        table[lo:hi] = want[lo:hi]
        ps_dist.all_gather_rows_(table, lo, hi, world)
        assert torch.equal(table, want), n
        # the neighbourhood-table exchange of the sharded precompute: int32 ids
        ids = torch.full((n, 5), -1, dtype=torch.int32)
        want_ids = torch.arange(n * 5, dtype=torch.int32).view(n, 5)
        ids[lo:hi] = want_ids[lo:hi]
        ps_dist.all_gather_rows_(ids, lo, hi, world)
        assert torch.equal(ids, want_ids), n
    # identically seeded replicas still draw DIFFERENT batches once attached (ps_dist.seed_rank_streams)
    import pinsage_training as pst
    torch.manual_seed(1234)                      # the mistake a data-parallel script makes on every rank
    ps_dist.seed_rank_streams(rank, world)
    assert pst.SAMPLER_RANK == rank
    positives = torch.arange(4000).view(2000, 2)
    b, _ = pst.sample_batch(torch.arange(5000), positives, 64, None, hard_negatives=False)
    both = [torch.zeros_like(b) for _ in range(world)]
    dist.all_gather(both, b)
    assert not torch.equal(both[0], both[1])
    assert not torch.equal(both[0][:, 2], both[1][:, 2])
    ps_dist.barrier()
    out.put(rank)
    dist.destroy_process_group()


def test_data_parallel_host_logic_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(out.get(timeout=5) for _ in range(2)) == [0, 1]


def test_shard_range_covers_everything():
    import ps_dist
    for n in (0, 1, 7, 20_000_000):
        for w in (1, 2, 8):
            r = [ps_dist.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
