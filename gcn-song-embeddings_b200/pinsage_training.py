"""B200 drop-in for the reference's `pinsage_training` module: same public names,
attributes and file formats (reference: /root/reference/pinsage_training.py).

`PinSage.train_batch` runs the fused device step of ps_engine.Engine.train_step (one
shared frontier for the q / pos / neg columns, max-margin loss + backward in CUDA,
then the optimiser).  Batch construction runs on the device without the O(P) randperm
and O(N) mask of the reference (pinsage_training.py:53-77) but draws from the same law.
wandb logging is optional (only imported when log=True).
"""
from __future__ import annotations

import os
import threading
import time

import torch
import torch.nn.functional as F
from tqdm import tqdm

import pinsage_model as psm
import ps_native

BASE_RUN_DIR = "./runs"


def max_margin_loss(h_q, h_pos, h_neg, margin):
    """mean(max(q^.n^ - q^.p^ + margin, 0)) with F.normalize'd rows
    (pinsage_training.py:31-41).  CUDA inputs run the fused loss kernel (forward only here;
    autograd users get the torch expression so gradients flow)."""
    if h_q.requires_grad or h_pos.requires_grad or h_neg.requires_grad or not h_q.is_cuda:
        h_q, h_pos, h_neg = F.normalize(h_q, dim=1), F.normalize(h_pos, dim=1), F.normalize(h_neg, dim=1)
        d = (h_q * h_neg).sum(1) - (h_q * h_pos).sum(1) + margin
        return torch.stack([d, torch.zeros_like(d)], 1).max(1).values.mean()
    B = h_q.shape[0]
    emb = torch.cat([h_q, h_pos, h_neg], 0).to(torch.float32).contiguous()
    idx = torch.arange(B, device="cuda", dtype=torch.int32)
    triples = torch.stack([idx, idx + B, idx + 2 * B], 1).contiguous()
    loss = torch.zeros(1, dtype=torch.float32, device="cuda")
    ps_native.margin_loss_fwd_bwd(emb, triples, margin, 1.0, None, loss, None)
    return loss[0]


def cosine_dissimilarity(a, b):
    return 1 - F.cosine_similarity(a, b)


COSINE_TRIPLET_LOSS = torch.nn.TripletMarginWithDistanceLoss(distance_function=cosine_dissimilarity,
                                                            margin=0.0001, reduction="mean")


# ---- batch construction ----------------------------------------------------------------

def _distinct_randint(high, k, device):
    """k distinct uniform draws from [0, high) without an O(high) permutation: same law as
    randperm(high)[:k] (pinsage_training.py:58,74).  Falls back to randperm when k is a
    large fraction of high."""
    if k > high:
        k = high
    if high <= 4 * k or high < 65536:
        return torch.randperm(high, device=device)[:k]
    out = torch.empty(0, dtype=torch.int64, device=device)
    while out.numel() < k:
        cand = torch.randint(0, high, (int(1.25 * k) + 16,), device=device)
        cand = torch.cat([out, cand])
        # keep first occurrences in draw order (a uniformly random k-subset in random order)
        uniq, inv = torch.unique(cand, return_inverse=True)
        first = torch.full((uniq.numel(),), cand.numel(), dtype=torch.int64, device=device)
        first.scatter_reduce_(0, inv, torch.arange(cand.numel(), device=device), reduce="amin")
        out = cand[torch.sort(first).values]
    return out[:k]


def sample_positives_with_rep(positives, batch_size):
    """batch_size distinct random rows of `positives` (pinsage_training.py:53-62)."""
    sample = _distinct_randint(positives.shape[0], batch_size, positives.device)
    return positives[sample, :].to(torch.int64)


def sample_easy_negatives(all_ids, pos_batch):
    """One random node per pair, distinct, none of them in the positive batch
    (pinsage_training.py:64-77)."""
    n, B = all_ids.shape[0], pos_batch.shape[0]
    pos_nodeset = pos_batch.flatten().unique()
    if n - pos_nodeset.numel() < 8 * B or n < 65536:  # small id space: the reference's mask + randperm
        mask = torch.ones((n,), dtype=torch.bool, device=all_ids.device)
        mask[pos_nodeset] = False
        possible = all_ids[mask].to(torch.int64)
        negatives = possible[torch.randperm(possible.numel(), device=all_ids.device)[:B]]
    else:  # rejection sampling: uniform without replacement from the same set
        negatives = torch.empty(0, dtype=torch.int64, device=all_ids.device)
        while negatives.numel() < B:
            cand = _distinct_randint(n, 2 * B, all_ids.device)
            cand = cand[~torch.isin(cand, pos_nodeset)]
            negatives = torch.cat([negatives, cand[~torch.isin(cand, negatives)]])
        negatives = all_ids[negatives[:B]].to(torch.int64)
    batch = torch.cat((pos_batch, negatives.unsqueeze(1)), dim=1)
    nodeset = batch.flatten().unique().to(torch.int64)
    return batch, nodeset


def sample_hard_negatives(all_ids, pos_batch, nbhds, min_rank, max_rank, reference_compat=True):
    """One hard negative per pair: the query's PPR neighbour at a random rank in
    [min_rank, max_rank) (pinsage_training.py:79-87).  The reference gathers rows 0..B-1 of
    the neighbourhood table instead of the queries' rows (:84); reference_compat=True
    reproduces that, False uses the queries' rows."""
    queries = pos_batch[:, 0]
    rnd_ranks = torch.randint(min_rank, max_rank, (queries.shape[0],), device=pos_batch.device)
    rows = torch.arange(queries.shape[0], device=pos_batch.device) if reference_compat else queries
    if pos_batch.is_cuda:
        # the engine's device-resident int32 table (uploaded once), not the host int64 [N, 100] tuple: only B ids move
        from ps_engine import NeighborTable
        hard_neg = NeighborTable.of(nbhds).nodes[rows, rnd_ranks].to(torch.int64)
    else:
        hard_neg = nbhds[1][rows.cpu(), rnd_ranks.cpu()].to(torch.int64)
    batch = torch.cat((pos_batch, hard_neg.unsqueeze(1)), dim=1)
    nodeset = batch.flatten().unique().to(torch.int64)
    return batch, nodeset


_sampler_state = {"seed": None, "step": 0}
_sampler_lock = threading.Lock()
SAMPLER_RANK = 0  # data-parallel rank mixed into the device sampler's Philox counter (set by ps_dist.attach)


def sample_batch_device(all_ids, positives, batch_size):
    """Easy-negative batch int64 [B,3] drawn by ONE kernel (ps_sample_batch) with no host round trip: same law as
    sample_positives_with_rep + sample_easy_negatives.  The Philox key is derived once from torch's seed
    (torch.manual_seed controls it), the counter advances per call.  None if the sizes are outside the kernel's
    limits (the torch path below handles those)."""
    if os.environ.get("PS_NATIVE_SAMPLER", "1") == "0":
        return None
    if not (positives.is_cuda and all_ids.is_cuda and positives.dtype == torch.int64 and positives.is_contiguous()
            and ps_native.sample_batch_supported(positives.shape[0], all_ids.shape[0], batch_size)):
        return None
    st = _sampler_state
    with _sampler_lock:  # several preparation threads draw batches
        if st["seed"] != torch.initial_seed():
            st["seed"], st["step"] = torch.initial_seed(), 0
        st["step"] += 1
        # the rank occupies the high bits of the counter word: replicas that were seeded identically
        # (torch.manual_seed(S) on every rank) still draw different batches
        seed, step = st["seed"], st["step"] | (int(SAMPLER_RANK) << 40)
    return ps_native.sample_batch(positives, all_ids, all_ids.shape[0], batch_size, seed, step)


def sample_batch(all_ids, positives, batch_size, nbhds, hard_negatives=True, hn_min=10, hn_max=100):
    """(batch int64 [B,3], nodeset) (pinsage_training.py:89-97)."""
    if not hard_negatives:
        batch = sample_batch_device(all_ids, positives, batch_size)
        if batch is not None:
            return batch, batch.flatten().unique()
    pos_batch = sample_positives_with_rep(positives, batch_size)
    if hard_negatives:
        return sample_hard_negatives(all_ids, pos_batch, nbhds, hn_min, hn_max)
    return sample_easy_negatives(all_ids, pos_batch)


def batch_variance(h):
    """pinsage_training.py:99-103."""
    mean = torch.mean(h, dim=0)
    var = torch.sum(torch.pow(h - mean, 2)) / (h.shape[0] - 1)
    return torch.prod(var)


# ---- trainer -----------------------------------------------------------------------------

class PinSage():
    """The PinSage trainer (pinsage_training.py:108-295): same constructor, hyper-parameter
    attributes, methods and checkpoint format.  Features, graph, neighbourhood table and
    positives are resident in HBM after construction; `train_batch` accepts host or device
    batches."""

    def __init__(self, g, n_items, features, positives, log=True, load_save=True):
        self.run_name = "pinsage_randomft_intersect"
        self.precomp_path = getattr(g, "nbhds_path", None)

        self.g = g
        self.n = n_items
        self.all_ids = torch.arange(0, n_items, 1, dtype=torch.int64, device="cuda")
        self.features = features
        self.positives = positives

        self.n_layers = 2
        self.in_dim = features.shape[1]
        self.hidden_dim = 512
        self.out_dim = 128
        self.dimensions = (self.in_dim, self.hidden_dim, self.out_dim)
        self.n_hops = 500
        self.alpha = 0.85
        self.T = 3
        self.hard_negatives = False
        self.hn_min = 10
        self.hn_max = 100

        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            # data parallel: every rank walks 1/G of the sources, one all-gather of the shards (ps_dist)
            import ps_dist
            self.nbhds = ps_dist.precompute_neighborhoods_sharded(self.g, self.n, self.n_hops, self.alpha,
                                                                  psm.DEF_T_PRECOMP, self.precomp_path)
        else:
            self.nbhds = psm.precompute_neighborhoods_topt(self.g, self.n, self.n_hops, self.alpha,
                                                           psm.DEF_T_PRECOMP, self.precomp_path)
        self.model = psm.PinSageModel(self.g, self.n, self.n_layers, self.dimensions,
                                      self.n_hops, self.alpha, self.T, self.nbhds)

        self.lr = 1e-4
        self.decay = 0.95
        # torch.optim.Adam's interface and state_dict format, one fused kernel per step (ps_optim.FlatAdam)
        from ps_optim import FlatAdam
        self.optimizer = FlatAdam(self.model.parameters(), lr=self.lr, engine=self.model.engine)
        self.scheduler = torch.optim.lr_scheduler.ExponentialLR(self.optimizer, self.decay)
        self.margin = 1e-5
        self.epochs = 30
        self.batch_size = 128
        self.b_per_e = 500

        self.embeddings = None
        self.max_steps_in_flight = 2   # device queue depth in steps (None / 0 = unlimited)
        self.prep_workers = 1          # host threads preparing batches ahead of the training thread (prefetch_async)
        self.online_sampling = False   # True: run the walker inside every step (reference's online relevant_nodes_per_layer)
        self.reference_compat = True   # duplicate-node gradient factor, hard-negative row quirk
        self.diagnostics = True        # node-feature loss + batch variance, as train_batch returns them
        self.world_size, self.rank = 1, 0  # set by ps_dist.attach() for data-parallel runs
        self._grad_sync = None

        run_dir = os.path.join(BASE_RUN_DIR, self.run_name)
        os.makedirs(os.path.join(run_dir, "board"), exist_ok=True)

        self.e = 0
        self.b = 0

        self.log = log
        if self.log:
            import wandb
            self._wandb = wandb
            wandb.config = {"learning_rate": self.lr, "epochs": self.epochs, "batch_size": self.batch_size}
            wandb.init(project='gcn-song-embeddings', name=self.run_name)
            wandb.watch(self.model, log="all", log_freq=10, log_graph=True)

        self.load_save = load_save
        if self.load_save:
            self.load_model()

    # device-resident views of the inputs
    def _feats(self):
        return self.model.engine.features(self.features)

    def _positives_dev(self):
        p = getattr(self, "_pos_dev", None)
        if p is None or p.shape != self.positives.shape or getattr(self, "_pos_src", None) is not self.positives:
            self._pos_dev = self.positives.to("cuda", torch.int64)
            self._pos_src = self.positives
        return self._pos_dev

    def prefetch_async(self, batch=None, host_sampler=None):
        """prefetch() on a background host thread: returns a Future whose result train_batch() accepts.  The batch
        preparation holds ~100 small launches and 3 host syncs (the sizes of the three frontiers); with two batches
        in flight on the worker the training thread never waits for them, it only enqueues the step's kernels."""
        if getattr(self, "_prep_pool", None) is None:
            from concurrent.futures import ThreadPoolExecutor
            dev = torch.cuda.current_device()
            # prep_workers host threads, each with its own high-priority stream.  One is enough once the plan builder is
            # native (a preparation is ~25 launches + 3 host reads, ~2-4 ms); more helps a single process when
            # preparations queue behind resident persistent GEMMs, but 8 ranks x 2 workers oversubscribed the host
            # cores of an 8-GPU box (measured: 7.49 -> 8.11 ms per step), hence the default of 1.
            self._prep_pool = ThreadPoolExecutor(max_workers=max(1, int(getattr(self, "prep_workers", 1))), thread_name_prefix="ps_prepare",
                                                 initializer=lambda: torch.cuda.set_device(dev))
        if host_sampler is not None:  # the worker also draws the batch on the host (returns an int64 [B,3] host tensor)
            return self._prep_pool.submit(lambda: self.prefetch(host_sampler()))
        return self._prep_pool.submit(self.prefetch, batch)

    def close(self):
        """Stop the batch-preparation worker(s) (idempotent; a later prefetch_async() starts new ones)."""
        pool, self._prep_pool = getattr(self, "_prep_pool", None), None
        if pool is not None:
            pool.shutdown(wait=True, cancel_futures=True)

    def prefetch(self, batch=None):
        """Start preparing the NEXT batch (frontier plans, backward transposes: index work that does not depend
        on the weights) on a side stream while the current step's kernels run.  With batch=None a batch is drawn
        on the device by sample_batch.  Returns a handle to pass to train_batch."""
        sampler = None
        if batch is None:
            positives = self._positives_dev()
            def sampler():
                if not self.hard_negatives:
                    b = sample_batch_device(self.all_ids, positives, self.batch_size)  # one launch, no host sync
                    if b is not None:
                        return b
                return sample_batch(self.all_ids, positives, self.batch_size, self.nbhds,
                                    hard_negatives=self.hard_negatives, hn_min=self.hn_min, hn_max=self.hn_max)[0]
        else:
            batch = torch.as_tensor(batch)
        if not self.reference_compat:
            self.model.T = self.T  # the reference never re-reads T after construction (grid_search.py:46-47)
        if getattr(self, "online_sampling", False) != getattr(self, "_online_active", False):
            # online: neighbourhoods are re-sampled by the walker inside every step instead of read from the table
            self.model.nbhds = self.model.online_neighbors() if self.online_sampling else self.nbhds
            self._online_active = bool(self.online_sampling)
        return self.model.engine.prepare(batch, sampler)

    def train_batch(self, batch):
        """One optimiser step on a batch of (q, pos, neg) triples (pinsage_training.py:181-214).  `batch` is an
        int64 [B,3] tensor (host or device) or a handle from prefetch().  Returns (loss, node_feat_loss,
        variance) as 0-dim device tensors."""
        if hasattr(batch, "result"):  # a Future from prefetch_async()
            batch = batch.result()
        prep = batch if hasattr(batch, "plan") else self.prefetch(batch)
        batch = prep.batch
        feats = self._feats()
        # At most `max_steps_in_flight` steps queued on the device.  With an unbounded training queue the host runs far
        # ahead, the caching allocator has to find room for several steps' buffers at once (cudaMalloc inside the loop:
        # measured 6.57 ms/step with occasional 7.7 ms runs), and the batch preparation's kernels (another stream) queue
        # behind more resident GEMM CTAs.  Two steps keep the device busy across the step boundary (6.50 ms/step against
        # 6.69 ms with one) without either effect.
        done = getattr(self, "_steps_in_flight", None)
        if done is None:
            from collections import deque
            done = self._steps_in_flight = deque()
        while self.max_steps_in_flight and len(done) >= self.max_steps_in_flight:
            done.popleft().synchronize()
        loss, emb, triples = self.model.engine.train_step(feats, prep, self.margin, self.reference_compat, diagnostics=bool(self.diagnostics))
        if self._grad_sync is not None:
            self._grad_sync()
        self.optimizer.step()
        ev = torch.cuda.Event()
        ev.record()
        done.append(ev)
        if self.diagnostics:  # node-feature triplet loss + batch variance (pinsage_training.py:200-212)
            diag = self.model.engine.last_diag  # computed inside the fused step call
            if diag is None:
                diag = torch.empty(2, dtype=torch.float32, device="cuda")
                ps_native.train_diagnostics(feats, batch, emb, triples, COSINE_TRIPLET_LOSS.margin, diag)
            node_feat_loss, variance = diag[0], diag[1]
        else:
            node_feat_loss = variance = torch.zeros((), device="cuda")
        return loss[0], node_feat_loss, variance

    def train(self):
        """Train the model (pinsage_training.py:216-256)."""
        print("\033[0;33mTraining PinSage...\033[0m")
        from collections import deque
        pending = deque([self.prefetch_async() for _ in range(3)])  # three batches ahead: two in preparation, one ready
        while self.e < self.epochs:
            print(f"Training epoch {self.e+1}/{self.epochs}...")
            cur_lr = self.optimizer.param_groups[0]["lr"]
            t1 = time.time()
            pbar = tqdm(total=self.b_per_e)
            pbar.update(1)
            while self.b < self.b_per_e:
                loss, node_feat_loss, variance = self.train_batch(pending.popleft())  # launches this step's kernels ...
                pending.append(self.prefetch_async())                                 # ... while the worker prepares the next batches
                pbar.update(1)
                if self.b % 50 == 0:  # reading the loss synchronises the device: not every step
                    pbar.set_description(f"Loss = {float(loss)}, bathes done")
                if self.log:
                    self._wandb.log({'Train Loss': loss, 'Node Features Loss': node_feat_loss,
                                     'Batch Variance': variance, 'Learning Rate': cur_lr})
                if self.load_save:
                    self.save_model()
                self.b += 1
            print(f"{time.time() - t1}s elapsed.")
            pbar.close()
            self.b = 0
            self.e += 1
            self.scheduler.step()
        for fut in pending:  # batches prepared ahead but not needed any more
            fut.cancel()

    def embed(self, ids=None, bsize=None):
        """Node embeddings, optionally only for `ids` / in `bsize` batches
        (pinsage_training.py:258-275; like the reference, bsize ignores ids and embeds rows
        0..n-1).  Returned on the host, as the reference's callers expect."""
        if ids is None:
            ids = self.all_ids
        self.model.eval()
        if not self.reference_compat:
            self.model.T = self.T
        feats = self._feats()
        n = len(ids)
        if not bsize:
            out = self.model.engine.embed(feats, torch.as_tensor(ids).to("cuda", torch.int64))
        else:
            out = torch.zeros((n, self.out_dim), device="cuda")
            for i in range(0, n, bsize):
                rng = torch.arange(i, min(i + bsize, n), device="cuda")
                out[rng, :] = self.model.engine.embed(feats, rng)
        self.embeddings = out.cpu()
        return self.embeddings

    def load_model(self):
        load_path = os.path.join(BASE_RUN_DIR, self.run_name, "state.pt")
        if self.world_size > 1:  # data parallel: nobody reads while rank 0 may still be replacing the file
            import ps_dist
            ps_dist.barrier()
        if os.path.isfile(load_path):
            prog = torch.load(load_path, map_location="cuda")
            self.e = prog["epochs_done"]
            self.b = prog["batches_done"]
            self.model.load_state_dict(prog["model_state"])
            self.optimizer.load_state_dict(prog["optimizer_state"])
            print(f"Loaded existing model from {load_path}.")

    def save_model(self):
        """state.pt in the reference's format (pinsage_training.py:288-295).  Data parallel: the replicas are
        identical, so only rank 0 writes; the file is written beside its destination and renamed over it, so a reader
        (or a crash) never sees a torn checkpoint."""
        if self.world_size > 1 and self.rank != 0:
            return
        prog = {"epochs_done": self.e, "batches_done": self.b,
                "model_state": self.model.state_dict(), "optimizer_state": self.optimizer.state_dict()}
        path = os.path.join(BASE_RUN_DIR, self.run_name, "state.pt")
        tmp = f"{path}.tmp{os.getpid()}"
        torch.save(prog, tmp)
        os.replace(tmp, path)


def _matrix_path(emb_dir):
    """The single-tensor copy of an embedding directory lives BESIDE it (`<...>/emb` -> `<...>/emb.all.pt`): the
    reference's loaders count the files inside the directory (eval.py:100-103), so nothing may be added there."""
    return emb_dir.rstrip("/\\") + ".all.pt"


def save_embedding_matrix(emb_dir, track_ids, emb):
    """One torch.save of {"track_ids": [...], "emb": float32 [N, out]} beside `emb_dir` (atomic): the fast path of
    load_embeddings / EmbLoader.  The reference's format is one small file per track (pinsage_training.py:297-327):
    10^6 torch.save / torch.load calls at cfg3 scale, minutes of file-system work around a 0.04 s embedding pass."""
    path = _matrix_path(emb_dir)
    tmp = f"{path}.tmp{os.getpid()}"
    torch.save({"track_ids": list(track_ids), "emb": emb.detach().to("cpu", torch.float32).contiguous()}, tmp)
    os.replace(tmp, path)
    return path


def load_embedding_matrix(emb_dir, track_ids):
    """float32 [N, out] from the single-tensor copy if it exists and lists exactly `track_ids`, else None."""
    path = _matrix_path(emb_dir)
    if not os.path.isfile(path):
        return None
    blob = torch.load(path)
    if list(blob.get("track_ids", [])) != list(track_ids):
        return None
    return blob["emb"]


def save_embeddings(trainer, dataset, base_run_dir=BASE_RUN_DIR, override_run_name=None, per_track=True):
    """Embed all tracks and save them (pinsage_training.py:297-327): one `<track_id>.pt` (1-D float32 [out]) per
    track under `<base>/<run>/emb/`, skipping existing files, as the reference does -- plus ONE tensor with all of them
    beside that directory (save_embedding_matrix).  The embeddings come from a single layer-wise pass over the graph
    (Engine.embed_range) instead of N/256 frontier batches.  per_track=False writes only the single tensor."""
    track_ids = list(dataset.tracks)
    n = len(track_ids)
    run_name = override_run_name if override_run_name else trainer.run_name
    emb_dir = os.path.join(base_run_dir, run_name, "emb")
    os.makedirs(emb_dir, exist_ok=True)
    trainer.model.eval()
    with torch.no_grad():
        emb = trainer.model.engine.embed_range(trainer._feats(), 0, n).cpu()
    trainer.embeddings = emb
    save_embedding_matrix(emb_dir, track_ids, emb)
    if not per_track:
        return
    pbar = tqdm(total=n, desc="Saving embeddings")
    for i in range(n):
        save_path = os.path.join(emb_dir, track_ids[i] + ".pt")
        if not os.path.isfile(save_path):
            torch.save(emb[i, :].clone().detach(), save_path)
        if i % 256 == 255:
            pbar.update(256)
    pbar.close()


def load_embeddings(trainer, dataset, base_run_dir=BASE_RUN_DIR):
    """The saved embeddings as one [N, out] tensor (pinsage_training.py:330-339): from the single-tensor copy when it is
    there and current, else by stacking the per-track files like the reference."""
    emb_dir = os.path.join(base_run_dir, trainer.run_name, "emb")
    emb = load_embedding_matrix(emb_dir, list(dataset.tracks))
    if emb is not None:
        return emb
    return torch.stack([torch.load(os.path.join(emb_dir, t + ".pt")) for t in dataset.tracks], dim=0)


def train_and_save(dataset, track_ids, trainer):
    """pinsage_training.py:443-450 (without the eyeball kNN print)."""
    trainer.train()
    save_embeddings(trainer, dataset)
    return load_embeddings(trainer, dataset)
