"""Data parallelism for the PinSage trainer: one process per GPU, graph / features /
neighbourhood table replicated in each GPU's HBM, each rank draws its own batches, and the
only exchange per step is one allreduce of the flat fp32 gradient buffer (~1.6-2.3 MB)
over NCCL (NVLink 5 / NVSwitch).  Full-graph inference shards by node range with no
communication.  The reference is single-process; this is the one
parallelism strategy the engine adds (section 8e).

The helpers are backend-agnostic (`gloo` on CPU in the tests, `nccl` on GPUs)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_*).  Returns (rank, world_size, local_rank); a no-op single-process world when
    WORLD_SIZE is unset or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"),
                                rank=rank, world_size=world)
    return rank, world, local


def bind_to_gpu_cpus(device_index: int):
    """Pin this process to the CPU cores NVML reports as local to the GPU (its NUMA node), intersected with the
    cores the container allows.  The step is a chain of host <-> device round trips (batch preparation, the per-step
    loss read); a process that lands on the far NUMA node pays for each of them.  Returns a short description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = None
        try:
            bus = torch.cuda.get_device_properties(device_index).pci_bus_id
            for i in range(pynvml.nvmlDeviceGetCount()):
                hh = pynvml.nvmlDeviceGetHandleByIndex(i)
                if int(pynvml.nvmlDeviceGetPciInfo(hh).bus) == int(bus):
                    h = hh
                    break
        except Exception:
            h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        target = local & allowed
        if not target:
            return f"not bound: none of the {len(local)} GPU-local cores is among the {len(allowed)} allowed ones"
        if target != allowed:
            os.sched_setaffinity(0, target)
        return f"bound to {len(target)} GPU-local cores (of {len(allowed)} allowed)"
    except Exception as exc:
        return f"not bound: {exc!r}"


def allreduce_mean_(flat: torch.Tensor, world_size: int):
    """In-place mean of a flat gradient buffer over all ranks."""
    if world_size > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world_size)
    return flat


class GradSync:
    """The per-step gradient exchange of one replica.  The flat gradient is summed over the ranks in two parts: everything
    but layer 0's Q.weight / Q.bias is final once layer 0's W gradient is done (ps_train_step records an event there),
    so that part is reduced on NCCL's stream while the aggregation backward and the two largest GEMMs of the step still
    compute; only conv_layers.0.Q.* (the first parameters of the flat buffer, 0.5 MB) are exchanged after the last
    kernel.  The mean's 1 / world_size is folded into the Adam kernel (FlatAdam.grad_scale)."""

    def __init__(self, trainer, world_size):
        self.engine, self.world = trainer.model.engine, world_size
        self.engine.want_upper_grads_event = True
        opt = trainer.optimizer
        self.scale_in_optimizer = hasattr(opt, "grad_scale")
        if self.scale_in_optimizer:
            opt.grad_scale = 1.0 / world_size
        self.comm = None

    def __call__(self):
        eng, flat = self.engine, self.engine.flat_grad
        ev, off = getattr(eng, "upper_grads_event", None), getattr(eng, "upper_grads_offset", 0)
        if ev is not None and 0 < off < flat.numel():
            if self.comm is None:
                self.comm = torch.cuda.Stream(priority=-1)
            self.comm.wait_event(ev)
            with torch.cuda.stream(self.comm):
                early = dist.all_reduce(flat[off:], op=dist.ReduceOp.SUM, async_op=True)
            dist.all_reduce(flat[:off], op=dist.ReduceOp.SUM)
            early.wait()  # the current stream waits for the early part
            eng.upper_grads_event = None
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if not self.scale_in_optimizer:
            flat.div_(self.world)


def broadcast_parameters(model, src=0):
    """Make every rank start from rank `src`'s parameters."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for p in model.parameters():
            dist.broadcast(p.data, src=src)


def seed_rank_streams(rank: int, world_size: int):
    """Make the batch samplers of the replicas draw DIFFERENT batches even when every rank called the same
    torch.manual_seed(S): the rank is mixed into the device sampler's Philox counter (pinsage_training.SAMPLER_RANK)
    and the torch generators behind the randperm / randint fallback paths are re-seeded with a per-rank offset.
    Without this an identically seeded N-GPU run computes N copies of one gradient."""
    import pinsage_training as pst
    pst.SAMPLER_RANK = int(rank)
    if world_size > 1 and rank > 0:
        seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * rank) & 0x7FFFFFFFFFFFFFFF
        torch.manual_seed(seed)  # seeds the CPU and every CUDA generator


def attach(trainer, rank: int, world_size: int):
    """Turn a PinSage trainer into one data-parallel replica: parameters broadcast from
    rank 0, gradient mean-allreduce before every optimiser step, rank-distinct batch streams.
    Checkpoints: only rank 0 writes state.pt (PinSage.save_model), every rank loads it after a barrier."""
    trainer.rank, trainer.world_size = rank, world_size
    seed_rank_streams(rank, world_size)
    if world_size > 1:
        broadcast_parameters(trainer.model)
        trainer._grad_sync = GradSync(trainer, world_size)
    return trainer


def shard_range(n: int, rank: int, world_size: int):
    """Contiguous node range [lo, hi) of rank `rank` (full-graph inference shards)."""
    per = -(-n // world_size)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def all_gather_rows_(table: torch.Tensor, lo: int, hi: int, world_size: int):
    """Fill the rows of `table` [n, d] that other ranks own: this rank holds rows [lo, hi) = shard_range(n, rank,
    world_size); afterwards every rank holds the full table.  One all_gather_into_tensor of equal-sized (padded)
    shards; with even shards the gather writes straight into the table."""
    n, d = table.shape
    per = -(-n // world_size)
    if per * world_size == n:
        dist.all_gather_into_tensor(table, table[lo:hi].clone())
        return table
    mine = torch.zeros((per, d), dtype=table.dtype, device=table.device)
    mine[: hi - lo] = table[lo:hi]
    full = torch.empty((per * world_size, d), dtype=table.dtype, device=table.device)
    dist.all_gather_into_tensor(full, mine)
    table.copy_(full[:n])
    return table


def precompute_neighborhoods_sharded(g, n_items, n_hops, alpha, T, path, rank: int = None, world_size: int = None, seed=None):
    """Data-parallel form of precompute_neighborhoods_topt (pinsage_model.py:109-132): rank r walks the sources
    [r N / G, (r+1) N / G) (independent sources: draws are keyed by (seed, source, step), so the table does not depend
    on the sharding), then ONE all-gather of the [N/G, T] shards (int32 ids + float32 weights: 0.8 GB in total at
    1 M items, T = 100) leaves the full device table on every rank.  Returns the reference's (weights float64 [N,T],
    nodes int64 [N,T]) tuple with the engine-native device table attached; rank 0 writes `path` (atomically) when given.
    A cached file of the right shape is loaded by every rank instead, like the reference does."""
    import os
    import pinsage_model as psm
    import ps_native
    from ps_engine import NeighborTable
    from ps_graph import as_psgraph
    rank = (dist.get_rank() if dist.is_initialized() else 0) if rank is None else rank
    world_size = (dist.get_world_size() if dist.is_initialized() else 1) if world_size is None else world_size
    if path is not None and os.path.isfile(path):
        weights, nodes = torch.load(path)
        if weights.shape[0] == n_items and weights.shape[1] == T:
            return (weights, nodes)
    if seed is None:
        seed = psm._next_seed()
    if world_size > 1:  # one key for everybody: rank 0's
        s = torch.tensor([seed & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device="cuda" if dist.get_backend() == "nccl" else "cpu")
        dist.broadcast(s, src=0)
        seed = int(s[0])
    lo, hi = shard_range(n_items, rank, world_size)
    pg = as_psgraph(g, n_items)
    out = ps_native.walk_topt(pg.device(), torch.arange(lo, hi, device="cuda"), n_hops, alpha, T, seed, want_i64=False, want_i32=True)
    table = NeighborTable.__new__(NeighborTable)
    table.n, table.Tp, table.scratch = n_items, T, {}
    if world_size > 1:
        table.nodes = torch.empty((n_items, T), dtype=torch.int32, device="cuda")
        table.w = torch.empty((n_items, T), dtype=torch.float32, device="cuda")
        table.nodes[lo:hi], table.w[lo:hi] = out["nodes_i32"], out["weights_f32"]
        all_gather_rows_(table.nodes, lo, hi, world_size)   # the single exchange step of the precompute
        all_gather_rows_(table.w, lo, hi, world_size)
    else:
        table.nodes, table.w = out["nodes_i32"], out["weights_f32"]
    # boundary format of the reference: float64 weights = count / n_hops.  Only the float32 table was exchanged; the
    # integer count is recovered from it exactly (n_hops <= 16384) and divided again in float64
    counts = torch.round(table.w.double() * n_hops).cpu()
    # IEEE division element by element (a framework scalar division may multiply by the reciprocal instead)
    weights, nodes = counts / torch.full_like(counts, float(n_hops)), table.nodes.to(torch.int64).cpu()
    if path is not None and rank == 0:
        tmp = f"{path}.tmp{os.getpid()}"
        torch.save((weights, nodes), tmp)
        os.replace(tmp, path)
    weights._ps_table = table
    return (weights, nodes)


def embed_shard(trainer, rank: int = None, world_size: int = None, chunk: int = 1 << 18, stats=None, exchange: bool = False):
    """Node-range sharded full-graph inference (BASELINE.json configs[3]): this rank's rows
    [lo, hi) of the embedding matrix, float32 on the device.  Graph, features and neighbourhood table are replicated.
    exchange=False (the contract): no communication, every rank recomputes the T-hop closure of its range.
    exchange=True: every rank transforms, projects and aggregates its own rows only; per layer the projected rows
    (leaky(Q x + b) W2^T, [N, out_dim] fp32) are all-gathered over NCCL / NVLink (Engine._embed_range_exchange): same
    result, 1 / world of every kernel's work per rank, L all-gathers.
    Returns (lo, hi, embeddings)."""
    rank = trainer.rank if rank is None else rank
    world_size = trainer.world_size if world_size is None else world_size
    lo, hi = shard_range(trainer.n, rank, world_size)
    trainer.model.eval()
    gather = None
    if exchange and world_size > 1:
        gather = lambda table, lo_, hi_: all_gather_rows_(table, lo_, hi_, world_size)
    emb = trainer.model.engine.embed_range(trainer._feats(), lo, hi, chunk=chunk, stats=stats, gather_layer=gather)
    return lo, hi, emb


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def shutdown():
    """Tear the process group down (after the last collective)."""
    if dist.is_initialized():
        dist.destroy_process_group()


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
