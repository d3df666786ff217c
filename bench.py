#!/usr/bin/env python
"""Benchmark of the PinSage hot path (BASELINE.json metric: train nodes/sec,
sample+fwd+bwd; walk steps/sec).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload cfg3|cfg1|cfg2|micro]
                    [--sampling precomp|online] [--mode infer --workload cfg4|cfg4q [--exchange]]

Our arm (default): the drop-in trainer `pinsage_training.PinSage` on the synthetic
1 M tracks / 200 k playlists / 40 M edges graph of BASELINE.json configs[2] (256-d
features, 2 layers, T=50, batch 1024 per GPU, precomputed neighbourhoods as the reference
does by default).  One step = sample a batch + PinSage.train_batch (shared-frontier
forward, max-margin loss, backward, Adam); batches are prepared three ahead by the trainer's
worker thread (PinSage.prefetch_async), as PinSage.train() does.  `value` is timed with the
batch sampled on the device (everything resident in HBM); `e2e` goes through the same public
call with HOST batches: host sampling -> pinned buffer -> H2D -> step -> D2H of the loss,
every step.  `roofline`: CUDA events around every ABI call of the timed region, the kernel
with the largest share reported against its bound (`traffic` from the committed ncu capture).
Multi-GPU: one process per GPU (torchrun), data parallel, one NCCL allreduce of the flat
gradient per step; weak scaling (batch 1024 per GPU).

Reference arm (--impl reference): the CPU oracle port of the reference's train step
(oracle/oracle.py, full-table clones and dense autograd included) on the box's host cores,
same graph generator / config, bounded sample (small batch) per step.

--mode infer: node-range sharded full-graph embedding (BASELINE.json configs[3]); not the
headline line.
"""
from __future__ import annotations

import argparse
import json
import os

# More hardware work queues than the default 8 before the CUDA context exists: the batch-preparation stream must not
# share a queue with the training stream (measured: when the two alias, every host sync of the preparation waits for
# a whole queued train step and the step time doubles).  ps_native does the same for library users.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "gcn-song-embeddings_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "pinsage_train_nodes_per_sec"
UNIT = "nodes/s"

WORKLOADS = {
    # BASELINE.json configs[2]
    "cfg3": dict(n_tracks=1_000_000, n_cols=200_000, n_edges=40_000_000, din=256, n_layers=2, T=50, batch=1024,
                 n_pos=10_000_000, ref_batch=128),
    # BASELINE.json configs[0] stand-in (dataset_micro is not in the reference checkout): its size, the reference's defaults
    "cfg1": dict(n_tracks=4_324, n_cols=1_500, n_edges=60_000, din=512, n_layers=2, T=3, batch=128, n_pos=5_000, ref_batch=128),
    # BASELINE.json configs[1] stand-in at dataset_final_intersect scale
    "cfg2": dict(n_tracks=50_000, n_cols=10_000, n_edges=500_000, din=512, n_layers=2, T=3, batch=128, n_pos=200_000, ref_batch=128),
    # BASELINE.json configs[3]: full-graph embedding inference sharded by node range (run with --mode infer)
    "cfg4": dict(n_tracks=20_000_000, n_cols=4_000_000, n_edges=1_000_000_000, din=512, n_layers=3, T=50, batch=0, n_pos=0, ref_batch=0),
    # the same inference path at 1/4 of the size (development)
    "cfg4q": dict(n_tracks=5_000_000, n_cols=1_000_000, n_edges=250_000_000, din=512, n_layers=3, T=50, batch=0, n_pos=0, ref_batch=0),
    # small stand-in for quick checks (not a bench line)
    "micro": dict(n_tracks=20_000, n_cols=4_000, n_edges=400_000, din=256, n_layers=2, T=50, batch=256,
                  n_pos=200_000, ref_batch=8),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "source": "fallback"}  # B200_PROFILING.md


class ClockSampler:
    """SM clock / throttle-reason samples during the timed region, through NVML on a background thread
    (nvidia_ml_py).  A polling `nvidia-smi -lms` child process was measured to stall kernel launches for
    milliseconds on a fresh box (driver locks), doubling the step time of the first run; NVML queries from
    inside the process do not.  Falls back to single nvidia-smi queries when NVML is unavailable."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index, period_s=0.1):
        import threading
        self.sm, self.smax, self.reasons, self.err = [], [], set(), None
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, args=(gpu_index, period_s), daemon=True)
        self._thread.start()

    def _run(self, gpu_index, period_s):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES-safe: address the device by the PCI bus id torch reports
            bus = torch.cuda.get_device_properties(gpu_index).pci_bus_id if hasattr(torch.cuda.get_device_properties(gpu_index), "pci_bus_id") else None
            h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hh = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if int(pynvml.nvmlDeviceGetPciInfo(hh).bus) == int(bus):
                        h = hh
                        break
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            smax = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self._stop.is_set():
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                self.smax.append(float(smax))
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
                self._stop.wait(period_s)
        except Exception as exc:  # NVML missing: one nvidia-smi query, outside any polling loop
            self.err = repr(exc)
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i", str(gpu_index)],
                                     capture_output=True, text=True, timeout=20).stdout.strip().split(",")
                self.sm.append(float(out[0])); self.smax.append(float(out[1]))
            except Exception:
                pass

    def stop(self):
        self._stop.set()
        self._thread.join(timeout=5)
        out = {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.smax) if self.smax else None,
               "samples": len(self.sm), "reasons": sorted(self.reasons), "source": "nvml" if self.err is None else "nvidia-smi (single query)"}
        if not self.sm:
            out["reasons"] = ["clock query unavailable: " + str(self.err)]
        return out


def measured_traffic(tag):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the committed `ncu --set full`
    capture of this workload (profiles/traffic.json names the capture); None if the kernel was not captured."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.isfile(path):
        return None
    with open(path) as f:
        return json.load(f).get("dram_bytes_per_launch", {}).get(tag)


def walk_roofline(n_sources, n_hops, ms, peaks):
    """The walker against its bounds: algorithmic bytes (28 B per step), the DRAM bytes one launch
    really moves (ncu capture named in profiles/traffic.json, scaled by the number of steps), and the DRAM
    RANDOM-ACCESS ceiling measured with tools/random_access_bench.cu: every step ends in a 4-byte read at a random
    place of a CSR that is far larger than L2, which costs a full 64-byte burst, and the memory system serves a fixed
    number of such bursts per second whatever the kernel does."""
    steps = n_sources * n_hops
    out = {"algorithmic_gbs": round(steps * 28 / ms / 1e6, 2), "frac_of_hbm": round(steps * 28 / ms / 1e6 / peaks["hbm_gbs"], 4)}
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(path):
        with open(path) as f:
            t = json.load(f)
        per_step = t.get("walk_dram_bytes_per_step")
        rate = t.get("random_access_gacc_per_s_320MB_table")
        if per_step:
            out["traffic"] = round(per_step * steps)
            out["dram_gbs"] = round(per_step * steps / ms / 1e6, 1)
            out["frac_of_hbm_dram_measured"] = round(per_step * steps / ms / 1e6 / peaks["hbm_gbs"], 4)
            if rate:
                bound_ms = per_step * steps / 64.0 / (rate * 1e9) * 1e3
                out["random_access_bound"] = {"bursts_per_step": round(per_step / 64.0, 3), "gacc_per_s": rate, "bound_ms": round(bound_ms, 2),
                                              "frac_of_bound": round(bound_ms / ms, 3), "source": t.get("random_access_source")}
    return out


def roofline_from_profile(summary, steps, peaks, dims=None):
    """Pick the kernel with the largest share of the step and report it against its bound.  dims = (din, dh, dout, T) of
    the workload: lets the aggregation kernels be reported in SURVEY.md section 8d's algorithmic unit."""
    if not summary:
        return None
    total = sum(v["ms"] for v in summary.values())
    tag, top = max(summary.items(), key=lambda kv: kv[1]["ms"])
    per_launch_ms = top["ms"] / top["launches"]
    is_gemm = tag.startswith("gemm")
    extra = {}
    if is_gemm:
        achieved = top["flops"] / top["launches"] / (per_launch_ms * 1e-3) / 1e12
        peak, unit, bound = peaks["tflops"], "TFLOP/s", "tensor"
        # fp32 parity needs the 3xTF32 split: every algorithmic FLOP is 3 tf32 tensor-core FLOPs, and dense tf32 runs at
        # half the bf16 rate, so the fp32-equivalent ceiling of the tensor pipe is peak / 6
        extra = {"note": "fp32-in/fp32-out GEMM computed as an error-compensated 3xTF32 product on tcgen05: `frac` divides the "
                         "algorithmic fp32 FLOPs by the bf16 peak; the tensor pipe executes 3 tf32 MMAs per algorithmic product",
                 "issued_tf32_tflops": round(3 * achieved, 2), "fp32_equivalent_ceiling": round(peak / 6, 1),
                 "frac_of_fp32_equivalent_ceiling": round(achieved / (peak / 6), 4)}
        if peaks.get("tf32_tflops_sustained"):  # measured in this run (cuBLAS TF32 GEMM): the pipe's own dense rate
            extra["tf32_dense_measured_tflops"] = peaks["tf32_tflops_sustained"]
            extra["tensor_pipe_frac_vs_measured_tf32"] = round(3 * achieved / peaks["tf32_tflops_sustained"], 4)
    else:
        achieved = top["bytes"] / top["launches"] / (per_launch_ms * 1e-3) / 1e9
        peak, unit, bound = peaks["hbm_gbs"], "GB/s", "hbm"
        traffic = measured_traffic(tag)
        extra = {}
        if tag.startswith("aggregate_fwd") and dims is not None:
            # SURVEY.md section 8d's unit per target: (T+1) gathered INPUT rows of Din floats + T (index, weight) pairs + the Do-wide
            # output row.  The kernel itself gathers the TRANSFORMED rows (Dh wide, Q applied once per distinct row instead of once
            # per gathered row), so the bytes it addresses per target are larger: both figures are reported.
            din, dh, dout, T = dims
            din_l = din if tag.endswith("_l0") else dout
            per_target_kernel = T * dh * 4 + din_l * 4 + T * 8 + 4 + (din_l + dh) * 4 + 4
            per_target_survey = (T + 1) * din_l * 4 + T * 8 + dout * 4
            targets = top["bytes"] / top["launches"] / per_target_kernel
            extra = {"kernel_addressed_gbs": round(achieved, 1), "kernel_addressed_bytes_per_target": per_target_kernel,
                     "algorithmic_bytes_per_target": per_target_survey, "targets_per_launch": round(targets)}
            top = dict(top, bytes=targets * per_target_survey * top["launches"])
            achieved = top["bytes"] / top["launches"] / (per_launch_ms * 1e-3) / 1e9
        extra["note"] = ("`achieved` = SURVEY.md section 8d's algorithmic bytes per target x targets per launch over the launch time; the kernel "
                         "addresses more (Dh-wide transformed rows, `kernel_addressed_gbs`), part of which L2 serves: `traffic` / "
                         "`frac_dram_measured` are the bytes the DRAM pins really move (ncu) over the same time")
        if traffic:
            extra["dram_gbs"] = round(traffic / (per_launch_ms * 1e-3) / 1e9, 1)
            extra["frac_dram_measured"] = round(traffic / (per_launch_ms * 1e-3) / 1e9 / peak, 4)
    shares = {k: round(v["ms"] / total, 4) for k, v in sorted(summary.items(), key=lambda kv: -kv[1]["ms"])[:8]}
    return {"kernel": tag, "bound": bound, "achieved": round(achieved, 2), "peak": peak, "unit": unit,
            "frac": round(achieved / peak, 4), "traffic": measured_traffic(tag), "peak_source": peaks["source"],
            "algorithmic_per_launch": (top["flops"] if is_gemm else top["bytes"]) / top["launches"],
            **extra, "ms_per_launch": round(per_launch_ms, 4), "launches_per_step": top["launches"] / steps,
            "share_of_kernel_time": round(top["ms"] / total, 4), "kernel_time_shares": shares,
            "all": {k: {"ms_per_step": round(v["ms"] / steps, 4),
                        "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["ms"] > 0 else None,
                        "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None}
                    for k, v in summary.items()}}


# ------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference's train step)
# ------------------------------------------------------------------------------------------

def cpu_train_baseline(features_cpu, nbhds_cpu, dims, n_layers, T, batches, warmup, margin=1e-5, budget_s=150.0):
    """Oracle port of the reference's train step on the host cores.  Stops early once `budget_s` seconds of timed
    steps have run (the arm must finish within minutes); returns (nodes/s, s/step, steps timed)."""
    from oracle import oracle
    params = oracle.make_params(n_layers, dims, np.random.RandomState(0))
    tr = oracle.OracleTrainer(params, features_cpu, nbhds_cpu, T=T, n_layers=n_layers, margin=margin)
    times = []
    for i, b in enumerate(batches):
        t0 = time.perf_counter()
        tr.train_batch(b)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            if sum(times) > budget_s:
                break
    B = batches[0].shape[0]
    return 3 * B / (sum(times) / len(times)), sum(times) / len(times), len(times)


def cpu_walk_baseline(indptr, indices, n_tracks, n_hops=500, alpha=0.85, T=100, per_core=8192):
    """The oracle's restatement of do_random_walks + sample_neighborhood_topt (pinsage_model.py:32-53, 88-107;
    numpy, vectorised over sources, steps sequential) on the box's host cores: one core on `per_core` sources, then
    every core at once on its own disjoint source range (one process per core, oracle/walk_baseline.py, started
    together through a file barrier).  NOTE the restatement is ~500x faster per core than the reference's own
    pure-Python loop (1.85e4 steps/s/core): it is the stronger baseline."""
    import shutil
    cores = len(os.sched_getaffinity(0))
    per_core = max(64, min(per_core, n_tracks // max(cores, 1)))
    tmp = tempfile.mkdtemp(prefix="pswalk_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    path = os.path.join(tmp, "graph.npy")
    script = os.path.join(ROOT, "oracle", "walk_baseline.py")
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")

    def run(n_proc):
        sync = tempfile.mkdtemp(prefix="sync_", dir=tmp)
        procs = [subprocess.Popen([sys.executable, script, path, str(c * per_core), str((c + 1) * per_core), str(n_hops), str(alpha), str(T), sync],
                                  stdout=subprocess.PIPE, text=True, env=env) for c in range(n_proc)]
        t0 = time.time()
        while len([f for f in os.listdir(sync) if f.startswith("ready.")]) < n_proc and time.time() - t0 < 240:
            if any(p.poll() is not None for p in procs):
                break
            time.sleep(0.01)
        open(os.path.join(sync, "go"), "w").close()
        outs = [p.communicate(timeout=600)[0] for p in procs]
        return max(float(o.strip().splitlines()[-1]) for o in outs)

    try:
        np.save(path, np.concatenate([np.array([indptr.shape[0] - 1], dtype=np.int64), indptr.astype(np.int64), indices.astype(np.int64)]))
        single = run(1)
        allc = run(cores)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return {"single_core_steps_per_s": round(per_core * n_hops / single, 1), "all_cores_steps_per_s": round(cores * per_core * n_hops / allc, 1),
            "cores": cores, "kind": "port",
            "sample": f"{per_core} sources x {n_hops} hops + top-{T} per core (oracle.do_random_walks_philox + topt_from_trace, numpy); "
                      f"the reference's own Python loop does 1.85e4 steps/s/core (BASELINE.md section 2)"}


def needed_nbhds_cpu(indptr, indices, n_tracks, batches, T, n_layers, n_hops=500, alpha=0.85, Tp=100, seed=7):
    """Neighbourhood rows for exactly the nodes the sample batches touch, from the oracle's
    Philox walker (the reference arm has no GPU-side precompute to lean on)."""
    from oracle import oracle
    w = np.zeros((n_tracks, Tp), dtype=np.float64)
    nodes = np.zeros((n_tracks, Tp), dtype=np.int64)
    have = np.zeros(n_tracks, dtype=bool)
    cur = np.unique(np.concatenate([b.reshape(-1) for b in batches]))
    for _ in range(n_layers):
        todo = cur[~have[cur]]
        if todo.size:
            trace = oracle.do_random_walks_philox(indptr, indices, todo, n_hops, alpha, seed)
            ww, nn = oracle.topt_from_trace(trace, todo, Tp)
            w[todo], nodes[todo], have[todo] = ww, nn, True
        cur = np.unique(np.concatenate([nodes[cur, :T].reshape(-1), cur]))
    return torch.from_numpy(w), torch.from_numpy(nodes)


def run_reference(args, wl):
    """--impl reference: the oracle port on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import ps_synth
    torch.manual_seed(0)
    try:  # torchrun exports OMP_NUM_THREADS=1; the CPU arm is entitled to every host core
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    t_setup = time.perf_counter()
    indptr, indices, n_e = ps_synth.bipartite_csr(wl["n_tracks"], wl["n_cols"], wl["n_edges"], seed=1234, device="cpu")
    feats = ps_synth.features(wl["n_tracks"], wl["din"], seed=1, device="cpu")
    positives = ps_synth.cooccurrence_positives(indptr, indices, wl["n_tracks"], min(wl["n_pos"], 1_000_000), seed=2)
    B = wl["ref_batch"]
    rng = np.random.RandomState(3)
    n_steps = args.warmup + args.steps
    batches = []
    for _ in range(n_steps):
        pairs = positives[torch.from_numpy(rng.choice(positives.shape[0], B, replace=False))].numpy()
        neg = rng.randint(0, wl["n_tracks"], size=(B, 1))
        batches.append(np.concatenate([pairs, neg], 1).astype(np.int64))
    nbhds = needed_nbhds_cpu(indptr.numpy(), indices.numpy(), wl["n_tracks"], batches, wl["T"], wl["n_layers"])
    setup_s = time.perf_counter() - t_setup
    value, s_per_step, timed = cpu_train_baseline(feats, nbhds, (wl["din"], 512, 128), wl["n_layers"], wl["T"], batches, args.warmup)
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(s_per_step * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, wl, 1, batch=B, parallelism=f"cpu{cores}"),
            "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{timed} timed steps of {B} triples (3*{B} nodes) on the full {wl['n_tracks']}-track graph, T={wl['T']}"
                                       + ("" if B == wl["batch"] else f" -- a bounded sample: the GPU arm's batch is {wl['batch']}; a CPU step at that batch "
                                          "needs tens of GB of dense autograd state and (at cfg3) minutes per step")},
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "steps_timed": timed, "setup_s": round(setup_s, 1)}
    print(json.dumps(line), flush=True)
    return 0


SAMPLING_DESC = ["precomputed neighbourhoods (n_hops=500, alpha=0.85, T_precomp=100), easy negatives"]


def workload_config(name, wl, n_gpus, batch=None, parallelism=None):
    """`batch` = triples per step and GPU that the arm REALLY runs (the reference arm's bounded sample uses
    wl['ref_batch'], not wl['batch']: its line says so here, so the two arms only compare as same-config when they are)."""
    batch = wl["batch"] if batch is None else batch
    return {"workload": f"{name}: synthetic bipartite {wl['n_tracks']} tracks / {wl['n_cols']} playlists / {wl['n_edges']} edges, "
                        f"{wl['din']}-d features, {wl['n_layers']} layers, T={wl['T']}, hidden 512, out 128, batch {batch}/GPU",
            "sampling": SAMPLING_DESC[0],
            "global_batch": batch * n_gpus, "parallelism": parallelism or f"dp{n_gpus}",
            "l2_policy": ("inputs larger than L2 (features 1.0 GB, transformed rows up to 2 GB per step vs 126 MB L2)" if name == "cfg3"
                          else "working set smaller than L2 at this size: a 256 MB buffer is written between steps" if wl["n_tracks"] * wl["din"] * 4 < 120e6
                          else "features larger than L2")}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------

def _drain(pending):
    for f in pending:
        if not f.cancel():
            try:
                f.result()
            except Exception:
                pass
    pending.clear()


def train_leg(trainer, steps, warmup, world, clock_dev=None, flush=None):
    """K timed steps of trainer.train_batch through the same prefetch pipeline as the headline leg (device-drawn
    batches), CUDA events + barrier on both sides, max over ranks.  Used for the `extra` sub-lines."""
    import ps_dist
    from collections import deque
    pend = deque([trainer.prefetch_async() for _ in range(3)])

    def step():
        if flush is not None:
            flush.fill_(0.0)
        out = trainer.train_batch(pend.popleft())
        pend.append(trainer.prefetch_async())
        return out

    for _ in range(warmup):
        step()
    torch.cuda.synchronize(); ps_dist.barrier()
    cs = ClockSampler(clock_dev) if clock_dev is not None else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()[0]
    e1.record()
    torch.cuda.synchronize(); ps_dist.barrier()
    ms = ps_dist.max_over_ranks(e0.elapsed_time(e1) / steps)
    _drain(pend)
    return {"ms_per_step": round(ms, 4), "value": round(3 * trainer.batch_size * world / (ms * 1e-3), 1), "unit": UNIT, "steps": steps,
            "warmup": warmup, "final_loss": float(loss), "clocks": cs.stop() if cs else None}


def small_config_leg(name, local, steps=100, warmup=20):
    """The reference's own default configuration (T=3, batch 128, 512-d: BASELINE.json configs[0] / configs[1] stand-ins)
    through the drop-in trainer on one GPU; a 256 MB buffer is written between steps (the working set fits in L2)."""
    import ps_synth
    import pinsage_training as pst
    wl = WORKLOADS[name]
    g = ps_synth.make_graph(wl["n_tracks"], wl["n_cols"], wl["n_edges"], seed=1234, device="cuda")
    feats = ps_synth.features(wl["n_tracks"], wl["din"], seed=1, device="cuda")
    positives = ps_synth.cooccurrence_positives(g.device().indptr, g.device().indices, wl["n_tracks"], wl["n_pos"], seed=2)
    cwd, tmp = os.getcwd(), tempfile.mkdtemp(prefix="psbench_")
    os.chdir(tmp); os.makedirs("runs", exist_ok=True)
    try:
        trainer = pst.PinSage(g, wl["n_tracks"], feats, positives, log=False, load_save=False)
    finally:
        os.chdir(cwd)
    trainer.T = wl["T"]; trainer.model.T = wl["T"]; trainer.batch_size = wl["batch"]
    flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    out = train_leg(trainer, steps, warmup, 1, clock_dev=local, flush=flush)
    trainer.close()
    out["config"] = workload_config(name, wl, 1)
    return out


def measure_tf32_peak():
    """Dense TF32 GEMM rate of this GPU (cuBLAS through torch.matmul on fp32 8192^3 with allow_tf32): the tensor-pipe
    denominator of the 3xTF32 GEMMs, which MEASURED_PEAKS.json (bf16 only) does not hold."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
        for _ in range(3):
            a @ b
        best, evs = 1e9, []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); evs.append((e0, e1))
        torch.cuda.synchronize()
        best = min(e0.elapsed_time(e1) for e0, e1 in evs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 150
        for _ in range(reps):
            a @ b
        e1.record(); torch.cuda.synchronize()
        return {"tf32_tflops_burst": round(2 * n ** 3 / (best * 1e-3) / 1e12, 1),
                "tf32_tflops_sustained": round(2 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12, 1),
                "how": "torch.matmul fp32 8192^3, allow_tf32: best of 10 (burst), 150 back to back (sustained)"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def run_ours(args, wl):
    import ps_dist
    import ps_native
    import ps_synth
    import pinsage_model as psm
    import pinsage_training as pst

    rank, world, local = ps_dist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device("cuda", local)
    peaks = load_peaks()
    # The GPU arm's host work is tiny tensor ops on two Python threads; an OpenMP team per op only adds contention
    # (measured: the e2e leg went bimodal, 7.3 / 12 ms per step).  The CPU baseline gets all cores back below.
    host_threads = torch.get_num_threads()
    torch.set_num_threads(1)
    allowed_cpus = os.sched_getaffinity(0)
    numa_note = ps_dist.bind_to_gpu_cpus(local)
    if args.tc_waves:
        ps_native.lib().ps_gemm_tc_waves(args.tc_waves)
    if args.gemm_reserve_sms:
        ps_native.lib().ps_gemm_tc_reserve_sms(args.gemm_reserve_sms)
    torch.manual_seed(1000 + rank)
    N, C, din, T, B, L = wl["n_tracks"], wl["n_cols"], wl["din"], wl["T"], wl["batch"], wl["n_layers"]

    # ---- setup (untimed): graph, features, positives resident in HBM ----
    g = ps_synth.make_graph(N, C, wl["n_edges"], seed=1234, device="cuda")
    gh = g.device()
    feats = ps_synth.features(N, din, seed=1, device="cuda")
    positives = ps_synth.cooccurrence_positives(gh.indptr, gh.indices, N, wl["n_pos"], seed=2)

    # ---- walker microbench (BASELINE.json "walk steps/sec"): all N sources x 500 hops, T=100 ----
    all_src = torch.arange(N, device="cuda")
    ps_native.walk_topt(gh, all_src, 500, 0.85, 100, seed=1, want_i64=False, want_i32=True)  # warm-up (clocks, caches)
    torch.cuda.synchronize()
    walk_runs = []
    for rep in range(3):  # median of 3 launches, each over all N sources (output 0.8 GB, graph 0.33 GB: larger than L2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ps_native.walk_topt(gh, all_src, 500, 0.85, 100, seed=2 + rep, want_i64=False, want_i32=True)
        e1.record(); torch.cuda.synchronize()
        walk_runs.append(e0.elapsed_time(e1))
    walk_ms = statistics.median(walk_runs)
    walk_steps_per_s = N * 500 / (walk_ms * 1e-3)

    # ---- the trainer, through the public drop-in API ----
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="psbench_")
    os.chdir(tmp)
    os.makedirs("runs", exist_ok=True)
    try:
        trainer = pst.PinSage(g, N, feats, positives, log=False, load_save=False)
    finally:
        os.chdir(cwd)
    trainer.T = T; trainer.model.T = T
    trainer.batch_size = B
    trainer.online_sampling = args.sampling == "online"
    trainer.prep_workers = args.prep_workers
    if args.steps_in_flight >= 0:
        trainer.max_steps_in_flight = args.steps_in_flight
    ps_dist.attach(trainer, rank, world)
    nbhds_cpu = trainer.nbhds

    from collections import deque
    # batches i+1 and i+2 are sampled + planned by a worker thread on a side stream while step i runs (PinSage.train does the same)
    pending = deque([trainer.prefetch_async() for _ in range(3)])

    host = {"train": 0.0, "prefetch": 0.0}

    flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda") if N * din * 4 < 120e6 else None

    def device_step():
        if flush is not None:
            flush.fill_(0.0)  # small workloads: evict L2 between steps (untagged, inside the timed region)
        h0 = time.perf_counter()
        out = trainer.train_batch(pending.popleft())
        h1 = time.perf_counter()
        pending.append(trainer.prefetch_async())
        host["train"] += h1 - h0; host["prefetch"] += time.perf_counter() - h1
        return out

    for _ in range(args.setup_steps):  # untimed: lets the caching allocator reach its steady state on both streams
        device_step()                  # (a cudaMalloc inside a step synchronises the device and stalls the pipeline)
    for _ in range(args.warmup):
        device_step()
    torch.cuda.synchronize(); ps_dist.barrier()
    dev_allocs0 = torch.cuda.memory_stats().get("num_device_alloc", 0)

    # ---- timed region: K steps, device-resident inputs ----
    sampler = ClockSampler(local) if rank == 0 else None
    ps_native.profiler = ps_native.Profiler()  # calls composed from Python (the online walker)
    ps_native.native_profile(True)             # the launches inside ps_train_step (CUDA events on the launching stream)
    launches0 = ps_native.launch_count
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); ps_dist.barrier()
    w0 = time.perf_counter()
    host["train"] = host["prefetch"] = 0.0
    t0.record()
    for _ in range(args.steps):
        torch.cuda.nvtx.range_push("ps_step")  # lets `ncu --nvtx --nvtx-include "ps_step/"` isolate the timed steps
        loss, _, _ = device_step()
        torch.cuda.nvtx.range_pop()
    t1.record()
    torch.cuda.synchronize(); ps_dist.barrier()
    wall = time.perf_counter() - w0
    dev_ms = t0.elapsed_time(t1)
    launches = ps_native.launch_count - launches0
    dev_allocs = torch.cuda.memory_stats().get("num_device_alloc", 0) - dev_allocs0
    host_ms = {k: round(v * 1e3 / args.steps, 3) for k, v in host.items()}
    clocks = sampler.stop() if sampler else None
    prof = ps_native.profiler.summary(); ps_native.profiler = None
    prof.update(ps_native.native_profile_summary()); ps_native.native_profile(False)
    step_ms = ps_dist.max_over_ranks(max(dev_ms, 0.0) / args.steps)
    wall_ms = ps_dist.max_over_ranks(wall * 1e3 / args.steps)
    # device events bracket the region; the wall clock is reported beside it (host-side frontier construction
    # synchronises, so the two agree)
    value = 3 * B * world / (step_ms * 1e-3)
    final_loss = float(loss)

    # ---- e2e: the same public call with HOST batches (pinned -> H2D -> step -> D2H of the loss) ----
    pos_cpu = positives.cpu()
    ids_cpu = torch.arange(N)
    import itertools
    import threading
    bufs = [torch.empty((B, 3), dtype=torch.int64).pin_memory() for _ in range(8)]  # at most 4 batches are alive at once
    buf_ids, buf_lock = itertools.count(), threading.Lock()
    pending_e2e = deque()

    pos_np = pos_cpu.numpy()
    host_rng = np.random.default_rng(4242 + rank)

    def host_sample_numpy():
        """The law of pst.sample_batch with easy negatives (B distinct positive rows; B distinct negatives outside the
        pairs), as ~10 numpy calls: the torch-CPU version costs the preparation thread 1-3 ms of small-op dispatch."""
        rows = np.empty(0, dtype=np.int64)
        while rows.size < B:
            cand = np.concatenate([rows, host_rng.integers(0, pos_np.shape[0], size=B + B // 4 + 16)])
            _, first = np.unique(cand, return_index=True)
            rows = cand[np.sort(first)]
        pairs = pos_np[rows[:B]]
        pos_nodes = np.unique(pairs)
        neg = np.empty(0, dtype=np.int64)
        while neg.size < B:
            cand = host_rng.integers(0, N, size=2 * B)
            cand = np.concatenate([neg, cand[~np.isin(cand, pos_nodes)]])
            _, first = np.unique(cand, return_index=True)
            neg = cand[np.sort(first)]
        return np.concatenate([pairs, neg[:B, None]], axis=1)

    def host_batch():  # runs on the preparation worker: host sampling into a pinned buffer (H2D happens in prefetch)
        with buf_lock:
            buf = bufs[next(buf_ids) % len(bufs)]
        buf.copy_(torch.from_numpy(host_sample_numpy()))
        return buf

    def e2e_step():
        while len(pending_e2e) < 3:
            pending_e2e.append(trainer.prefetch_async(host_sampler=host_batch))
        out = trainer.train_batch(pending_e2e.popleft())
        pending_e2e.append(trainer.prefetch_async(host_sampler=host_batch))
        return float(out[0])  # D2H read of the step's result

    for _ in range(args.setup_steps // 2 + args.warmup):  # untimed: allocator steady state for the host-batch path too
        e2e_step()
    torch.cuda.synchronize(); ps_dist.barrier()
    w0 = time.perf_counter()
    e2e_each = []
    for _ in range(args.steps):
        s0 = time.perf_counter()
        e2e_step()
        e2e_each.append((time.perf_counter() - s0) * 1e3)
    torch.cuda.synchronize(); ps_dist.barrier()
    e2e_ms = ps_dist.max_over_ranks((time.perf_counter() - w0) * 1e3 / args.steps)
    e2e_value = 3 * B * world / (e2e_ms * 1e-3)

    if args.torch_profile:  # optional: where the non-kernel time of a step goes (not a bench number); every rank steps
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as tp:
            for _ in range(3):
                device_step()
            torch.cuda.synchronize()
        if rank == 0:
            with open(args.torch_profile, "w") as f:
                f.write(tp.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70))
            tp.export_chrome_trace(args.torch_profile + ".trace.json")

    # ---- extra sub-lines (not the headline): strong scaling, online sampling, the cfg5 walker modes, the reference's
    #      own default configs, full-graph inference on this graph, the TF32 peak ----
    extra = {}
    _drain(pending); _drain(pending_e2e)
    if world > 1 and not args.no_extras:  # fixed GLOBAL batch of wl['batch'] triples (strong scaling)
        trainer.batch_size = max(1, B // world)
        leg = train_leg(trainer, args.steps, max(3, args.warmup), world, clock_dev=local if rank == 0 else None)
        leg["scaling"] = "strong"; leg["global_batch"] = trainer.batch_size * world
        extra["strong"] = leg
        trainer.batch_size = B
    if world == 1 and not args.no_extras:
        if args.sampling != "online":
            trainer.online_sampling = True
            leg = train_leg(trainer, min(args.steps, 10), 3, 1, clock_dev=local)
            leg["sampling"] = "online: ps_walk_topt (n_hops=500, alpha=0.85) on every layer's frontier inside every step"
            extra["online_sampling"] = leg
            trainer.online_sampling = False
            trainer.prefetch(); torch.cuda.synchronize()  # back to the table
        modes = {}
        for flen in (3, 4, 5):  # BASELINE.json configs[4]: 1e8 walks of length 3-5 (deterministic restart), top-50
            runs = []
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ps_native.walk_topt(gh, all_src, 100 * flen, 0.85, 50, seed=20 + rep, fixed_len=flen, want_i64=False, want_i32=True)
                e1.record(); torch.cuda.synchronize()
                if rep:
                    runs.append(e0.elapsed_time(e1))
            ms = statistics.median(runs)
            modes[f"len{flen}"] = {"walks": N * 100, "steps": N * 100 * flen, "T": 50, "ms": round(ms, 3),
                                   "steps_per_s": round(N * 100 * flen / (ms * 1e-3), 1),
                                   "frac_of_hbm_algorithmic": round(N * 100 * flen * 28 / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)}
        extra["walker_fixed_length"] = modes
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        runs = []
        for rep in range(3):  # full-graph inference of this graph (configs[3]'s path at cfg3 size): embed all N rows
            e0.record()
            emb_all = trainer.model.engine.embed_range(feats, 0, N)
            e1.record(); torch.cuda.synchronize()
            runs.append(e0.elapsed_time(e1))
            del emb_all
        extra["inference_full_graph"] = {"nodes": N, "ms": round(min(runs[1:]), 3), "nodes_per_s": round(N / (min(runs[1:]) * 1e-3), 1),
                                         "how": "Engine.embed_range(0, N): layer-wise, 2 layers, T=50 (bench.py --mode infer runs configs[3] itself)"}
        peaks.update(measure_tf32_peak())
    if world == 1 and not args.no_extras:
        for name in ("cfg1", "cfg2"):
            if name != args.workload:
                extra[name] = small_config_leg(name, local)

    trainer.close()

    line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(step_ms, 3), "wall_ms_per_step": round(wall_ms, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args.workload, wl, world),
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "ms_per_step": round(e2e_ms, 3),
                    "h2d_bytes_per_step": B * 3 * 8, "d2h_bytes_per_step": 4,
                    "ms_per_step_median": round(statistics.median(e2e_each), 3), "ms_per_step_max": round(max(e2e_each), 3)},
            "gpu_launches": launches, "clocks": clocks,
            "walk": {"metric": "walk_steps_per_sec", "value": round(walk_steps_per_s, 1), "unit": "steps/s",
                     "sources": N, "n_hops": 500, "alpha": 0.85, "T": 100, "ms": round(walk_ms, 3),
                     **walk_roofline(N, 500, walk_ms, peaks)},
            "final_loss": final_loss, "host_ms_per_step": host_ms, "cudaMallocs_in_timed_region": dev_allocs,
            "host_placement": numa_note}
    import ps_engine
    if ps_engine._PREP_TIMING is not None and ps_engine._PREP_TIMING["n"]:
        t = ps_engine._PREP_TIMING
        line["prep_timing_ms"] = {k: round(v * 1e3 / t["n"], 3) for k, v in t.items() if k != "n"}
    line["extra"] = extra
    if rank == 0:
        line["roofline"] = roofline_from_profile(prof, args.steps, peaks, dims=(din, 512, 128, T))
        line["peaks"] = peaks
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, allowed_cpus)
            torch.set_num_threads(host_threads)
            Bc = wl["ref_batch"]
            cpu_batches = [pst.sample_batch(ids_cpu, pos_cpu, Bc, nbhds_cpu, hard_negatives=False)[0].numpy() for _ in range(3)]
            cv, cs, ct = cpu_train_baseline(feats.cpu(), nbhds_cpu, (din, 512, 128), L, T, cpu_batches, warmup=1, budget_s=20.0)
            line["cpu_baseline"] = {"value": round(cv, 3), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{ct} timed steps (1 warm-up) of {Bc} triples (3*{Bc} nodes) on the same {N}-track graph, T={T}; {cs:.2f} s/step"}
            line["walk"]["cpu_baseline"] = cpu_walk_baseline(g.indptr.numpy(), g.indices.numpy(), N)
        print(json.dumps(line), flush=True)
    ps_dist.barrier()
    ps_dist.shutdown()
    return 0


def run_inference(args, wl):
    """--mode infer: full-graph embedding inference sharded by node range (BASELINE.json configs[3]); every rank
    embeds its rows [lo, hi) with no communication (ps_dist.embed_shard).  nodes/s = N / max-over-ranks time of one
    pass; the one-time neighbourhood precompute (walker over all N) is timed separately.  Not the headline bench line."""
    import ps_dist
    import ps_native
    import ps_synth
    import pinsage_model as psm
    rank, world, local = ps_dist.init_from_env()
    N, C, din, T, L = wl["n_tracks"], wl["n_cols"], wl["din"], wl["T"], wl["n_layers"]
    t0 = time.perf_counter()
    g = ps_synth.make_graph(N, C, wl["n_edges"], seed=1234, device="cuda")
    gh = g.device()
    feats = ps_synth.features(N, din, seed=1, device="cuda")
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ps_native.walk_topt(gh, torch.arange(N, device="cuda"), 500, 0.85, T, seed=7, want_i64=False, want_i32=True)
    e1.record(); torch.cuda.synchronize()
    walk_ms = e0.elapsed_time(e1)
    from ps_engine import NeighborTable
    table = NeighborTable.__new__(NeighborTable)
    table.nodes, table.w, table.n, table.Tp, table.scratch = out["nodes_i32"], out["weights_f32"], N, T, {}
    dims = (din, 512, 128)
    torch.manual_seed(0)  # the same random-init weights (xavier, bias 0.3: the model's own initialiser) on every rank
    model = psm.PinSageModel(g, N, L, dims, 500, 0.85, T, table)

    class _T:  # the two attributes embed_shard reads
        pass
    tr = _T(); tr.rank, tr.world_size, tr.n, tr.model = rank, world, N, model
    tr._feats = lambda: feats
    times, stats = [], {}
    for it in range(args.warmup + args.steps):
        torch.cuda.synchronize(); ps_dist.barrier()
        e0.record()
        lo, hi, emb = ps_dist.embed_shard(tr, chunk=1 << 18, stats=stats, exchange=args.exchange)
        e1.record(); torch.cuda.synchronize()
        if it >= args.warmup:
            times.append(e0.elapsed_time(e1))
        checksum = float(emb.double().sum())
        del emb
    ms = ps_dist.max_over_ranks(sum(times) / len(times))
    # one more (untimed) pass with CUDA events around every ABI call: where the pass goes, and the HBM-bound part against its roofline
    peaks = load_peaks()
    ps_native.profiler = ps_native.Profiler()
    ps_dist.embed_shard(tr, chunk=1 << 18, exchange=args.exchange)
    prof = ps_native.profiler.summary()
    ps_native.profiler = None
    kernels = {k: {"ms": round(v["ms"], 2), "launches": v["launches"], "tflops": round(v["flops"] / v["ms"] / 1e9, 1) if v["ms"] else 0.0,
                   "gbs": round(v["bytes"] / v["ms"] / 1e6, 1) if v["ms"] else 0.0} for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    agg = {k: v for k, v in prof.items() if k.startswith("aggregate_fwd")}
    roofline = None
    if agg:
        a_ms, a_bytes = sum(v["ms"] for v in agg.values()), sum(v["bytes"] for v in agg.values())
        roofline = {"kernel": "aggregate_fwd (all layers)", "bound": "hbm", "achieved": round(a_bytes / a_ms / 1e6, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": round(a_bytes / a_ms / 1e6 / peaks["hbm_gbs"], 4), "traffic": None, "peak_source": peaks["source"],
                    "ms_per_pass": round(a_ms, 2), "share_of_kernel_time": round(a_ms / max(sum(v["ms"] for v in prof.values()), 1e-9), 3),
                    "note": "algorithmic bytes: T gathered projection rows (out_dim wide) + the self row + indices / weights + the concatenated row written, per target"}
    if rank == 0:
        print(json.dumps({"metric": "pinsage_embed_nodes_per_sec", "value": round(N / (ms * 1e-3), 1), "unit": "nodes/s", "n_gpus": world,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_pass": round(ms, 2), "higher_is_better": True, "scaling": "strong",
                          "dtype": "f32", "data": "synthetic",
                          "config": {"workload": f"{args.workload}: {N} tracks / {C} playlists / {wl['n_edges']} edges, {din}-d features, {L} layers, T={T}, "
                                                 f"node-range shards over {world} GPU(s), " + ("layer outputs all-gathered (NCCL)" if args.exchange and world > 1 else "no communication"), "rank0_rows": hi - lo, "rank0_closure": stats},
                          "walk": {"steps_per_s": round(N * 500 / (walk_ms * 1e-3), 1), "ms": round(walk_ms, 2), "sources": N, "n_hops": 500, "T": T},
                          "setup_s": round(setup_s, 1), "hbm_gb_allocated": round(torch.cuda.max_memory_allocated() / 1e9, 1),
                          "checksum_rank0": checksum, "roofline": roofline, "kernels_rank0": kernels,
                          "cpu_baseline": None if N > 2_000_000 else "see the train mode's line",
                          "cpu_baseline_note": "the reference's PinSage.embed clones three [N, Din] tables per call (pinsage_training.py:258-275 -> "
                                               "pinsage_model.py:21-30): 3 x 41 GB at this size, not runnable on the box's host; the CPU port is timed "
                                               "on the train step (cfg3) instead"}), flush=True)
    ps_dist.barrier()
    ps_dist.shutdown()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="train", choices=["train", "infer"], help="infer = node-range sharded full-graph embedding (cfg4 / cfg4q)")
    ap.add_argument("--prep-workers", type=int, default=1, help="host threads that prepare batches ahead of the training thread")
    ap.add_argument("--sampling", default="precomp", choices=["precomp", "online"],
                    help="precomp = table lookup of neighbourhoods precomputed once (the reference's default); online = the walker runs inside every step")
    ap.add_argument("--setup-steps", type=int, default=8, help="extra untimed steps before the W warm-up steps (allocator steady state)")
    ap.add_argument("--exchange", action="store_true", help="infer mode: all-gather layer outputs instead of recomputing the closure")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--steps-in-flight", type=int, default=-1, help="override PinSage.max_steps_in_flight (0 = unlimited; development)")
    ap.add_argument("--no-extras", action="store_true", help="skip the `extra` sub-lines (strong scaling, online sampling, walker modes, cfg1 / cfg2, TF32 peak)")
    ap.add_argument("--tc-waves", type=int, default=0, help="override ps_gemm_tc_waves (development)")
    ap.add_argument("--gemm-reserve-sms", type=int, default=0, help="ps_gemm_tc_reserve_sms: SMs the persistent GEMMs leave free (development)")
    ap.add_argument("--torch-profile", default=None, help="write a torch.profiler kernel table of 3 extra steps here")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.sampling == "online":
        SAMPLING_DESC[0] = "online: the walker (n_hops=500, alpha=0.85, top-T) runs inside every step on the frontier of each layer, easy negatives"
    if args.mode == "infer":
        return run_inference(args, wl)
    if args.impl == "reference":
        return run_reference(args, wl)
    if args.warmup < 3:
        print("note: fewer than 3 warm-up steps requested; the contract asks for W >= 3", file=sys.stderr)
    return run_ours(args, wl)


if __name__ == "__main__":
    sys.exit(main())
