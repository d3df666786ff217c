"""ctypes binding of libpinsage_b200.so (C ABI declared in include/pinsage_b200.h).

There is deliberately NO fallback: if the shared library is missing or a CUDA device is
not available, every entry point raises.  PyTorch is used only for device memory and
streams; no torch type crosses the ABI (raw device pointers, sizes and a cudaStream_t).
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_uint64, c_void_p

# The engine overlaps batch preparation (side stream) with the train step (main stream); with the default 8 hardware
# work queues the two streams can alias to one queue, which serialises them.  Only effective before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import torch  # noqa: E402

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PS_LIB_PATH") or os.path.join(_HERE, "libpinsage_b200.so")  # PS_LIB_PATH: development builds

_lib = None
_tls = threading.local()  # the library's CUDA runtime keeps a current device per host thread


class NativeError(RuntimeError):
    pass


_SIGNATURES = {
    "ps_version": ([], c_int),
    "ps_last_error": ([], c_char_p),
    "ps_set_device": ([c_int], c_int),
    "ps_gemm_backend": ([c_int], c_int),
    "ps_gemm_ex": ([c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64,
                    c_int64, c_int64, c_int64, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p], c_int),
    "ps_gemm_mask_supported": ([c_int64, c_int64, c_int64], c_int),
    "ps_gemm_wgrad": ([c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p], c_int),
    "ps_gemm_tc_pack": ([c_int], c_int),
    "ps_gemm_tc_waves": ([c_int], c_int),
    "ps_gemm_tc_reserve_sms": ([c_int], c_int),
    "ps_gemm_tc_prefetch": ([c_int], c_int),
    "ps_gemm_tc_cluster": ([c_int], c_int),
    "ps_gemm_tc_trace": ([c_void_p], c_int),
    "ps_gemm_tc_experiment": ([c_int], c_int),
    "ps_graph_create": ([c_void_p, c_void_p, c_int64, c_int64, c_int64, ctypes.POINTER(c_void_p), c_void_p], c_int),
    "ps_graph_destroy": ([c_void_p], c_int),
    "ps_graph_use_indptr32": ([c_void_p, c_int], c_int),
    "ps_walk_algo": ([c_int], c_int),
    "ps_walk_topt": ([c_void_p, c_void_p, c_int64, c_int, c_double, c_int, c_int, c_uint64,
                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    "ps_trace_topt": ([c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    "ps_gemm": ([c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64,
                 c_int64, c_int64, c_int64, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p], c_int),
    "ps_aggregate_fwd": ([c_void_p, c_int64, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int,
                          c_int64, c_void_p, c_int64, c_void_p, c_void_p], c_int),
    "ps_aggregate_bwd": ([c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p,
                          c_void_p, c_int, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int, c_void_p], c_int),
    "ps_norm_leaky_bwd": ([c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_void_p], c_int),
    "ps_l2norm_rows": ([c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p], c_int),
    "ps_leaky_bwd": ([c_void_p, c_void_p, c_int64, c_void_p], c_int),
    "ps_colsum": ([c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p], c_int),
    "ps_scatter_add_rows": ([c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p], c_int),
    "ps_count_triples": ([c_void_p, c_int64, c_int64, c_void_p, c_void_p], c_int),
    "ps_margin_loss_fwd_bwd": ([c_void_p, c_int64, c_void_p, c_int64, c_int, c_float, c_float, c_void_p, c_int64,
                                c_void_p, c_void_p, c_int64, c_void_p], c_int),
    "ps_plan_layer": ([c_void_p, c_int64, c_int, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    "ps_plan_transpose": ([c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p], c_int),
    "ps_sample_batch_workspace": ([c_int], c_int64),
    "ps_sample_batch": ([c_void_p, c_int64, c_void_p, c_int64, c_int, c_uint64, c_uint64, c_void_p, c_void_p, c_int64,
                         c_void_p, c_void_p], c_int),
    "ps_topk_rows": ([c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p], c_int),
    "ps_train_diagnostics": ([c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_float, c_void_p, c_void_p], c_int),
    "ps_adam_step": ([c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float, c_int64,
                      c_float, c_void_p], c_int),
}

PS_MAX_LAYERS = 8


class LayerPlanC(ctypes.Structure):
    _fields_ = [("n", c_int64), ("nz", c_int64), ("self_rows", c_void_p), ("nbz", c_void_p), ("w", c_void_p), ("zrows", c_void_p),
                ("seg_off", c_void_p), ("pair_q", c_void_p), ("chunk_off", c_void_p), ("chunk_row", c_void_p)]


class LayerParamsC(ctypes.Structure):
    _fields_ = [(k, c_void_p) for k in ("Qw", "Qb", "Ww", "Wb", "gQw", "gQb", "gWw", "gWb")]


class StepArgsC(ctypes.Structure):
    """ps_step_args of include/pinsage_b200.h."""
    _fields_ = [("n_layers", ctypes.c_int32), ("T", ctypes.c_int32), ("in_dim", ctypes.c_int32), ("hidden_dim", ctypes.c_int32),
                ("out_dim", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("feats", c_void_p), ("ld_feats", c_int64),
                ("layers", LayerPlanC * PS_MAX_LAYERS), ("params", LayerParamsC * PS_MAX_LAYERS),
                ("G1w", c_void_p), ("G1b", c_void_p), ("G2w", c_void_p), ("gG1w", c_void_p), ("gG1b", c_void_p), ("gG2w", c_void_p),
                ("triples", c_void_p), ("B", c_int64), ("dup_counts", c_void_p),
                ("margin", c_float), ("feat_margin", c_float),
                ("flat_grad", c_void_p), ("n_params", c_int64),
                ("workspace", c_void_p), ("workspace_bytes", c_int64),
                ("loss_out", c_void_p), ("batch", c_void_p), ("diag_out", c_void_p), ("emb_out", ctypes.POINTER(c_void_p)),
                ("upper_grads_event", c_void_p)]


class PlanDescLayerC(ctypes.Structure):
    _fields_ = [(k, c_int64) for k in ("n", "nz", "off_nodes", "off_self_rows", "off_nbz", "off_w", "off_zrows",
                                       "off_seg_off", "off_pair_q", "off_chunk_off", "off_chunk_row")]


class PlanDescC(ctypes.Structure):
    """ps_plan_desc of include/pinsage_b200.h."""
    _fields_ = [("U", c_int64), ("off_top", c_int64), ("off_triples", c_int64), ("off_counts", c_int64),
                ("layers", PlanDescLayerC * PS_MAX_LAYERS), ("bytes_used", c_int64), ("bytes_needed", c_int64)]


_SIGNATURES.update({
    "ps_prepare_plan": ([c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_int64,
                         ctypes.POINTER(PlanDescC), c_void_p], c_int),
    "ps_prepare_plan_online": ([c_void_p, c_int64, c_void_p, c_int64, c_int, c_double, c_uint64, c_int, c_int, c_int, c_void_p, c_int64,
                                ctypes.POINTER(PlanDescC), c_void_p], c_int),
    "ps_train_step_workspace": ([ctypes.POINTER(StepArgsC)], c_int64),
    "ps_train_step": ([ctypes.POINTER(StepArgsC), c_void_p], c_int),
    "ps_profile_enable": ([c_int], c_int),
    "ps_profile_dump": ([c_char_p, c_int64], c_int64),
    "ps_gemm_filter": ([c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p], c_int),
    "ps_topk_rows_mapped": ([c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p], c_int),
    "ps_csr_build": ([c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p], c_int),
    "ps_standardize": ([c_void_p, c_int64, c_int64, c_int, c_double, c_void_p, c_void_p, c_void_p], c_int),
})

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load_library(path: str = LIB_PATH):
    """dlopen the library and declare every prototype (no device needed)."""
    if not os.path.isfile(path):
        raise NativeError(
            f"{path} is missing: build it with `python __graft_entry__.py build` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.argtypes = argtypes
        fn.restype = restype
    return lib


def lib():
    global _lib
    if _lib is None:
        _lib = load_library()
        if _lib.ps_version() != 1:
            raise NativeError("libpinsage_b200.so ABI version mismatch")
    return _lib


def _ensure_device():
    """The library links its own CUDA runtime; point it at torch's current device."""
    if not torch.cuda.is_available():
        raise NativeError("no CUDA device: the PinSage engine has no CPU fallback")
    dev = torch.cuda.current_device()
    if getattr(_tls, "device", None) != dev:
        check(lib().ps_set_device(dev))
        _tls.device = dev
    return dev


def check(rc: int):
    if rc != 0:
        raise NativeError(f"libpinsage_b200 error {rc}: {lib().ps_last_error().decode()}")


def _p(t, dtype=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeError("expected a CUDA tensor")
    if dtype is not None and t.dtype != dtype:
        raise NativeError(f"expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous() and t.dim() > 1 and t.stride(-1) != 1:
        raise NativeError("innermost dimension must be contiguous")
    return c_void_p(t.data_ptr())


launch_count = 0  # kernel-launching ABI calls made so far (bench.py reports the per-step delta)


class Profiler:
    """CUDA-event timing of individual ABI calls on the launching stream (bench.py's
    roofline leg).  Recording an event does not synchronise; summary() does."""

    def __init__(self):
        self.records = []

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for tag, e0, e1, flops, nbytes in self.records:
            d = out.setdefault(tag, {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0})
            d["ms"] += e0.elapsed_time(e1)
            d["launches"] += 1
            d["flops"] += flops
            d["bytes"] += nbytes
        return out


profiler = None  # set to a Profiler() to time every ABI call


def native_profile(on: bool):
    """Device timing of the launches inside ps_train_step (csrc/step.cu): enable / disable + reset."""
    check(lib().ps_profile_enable(int(bool(on))))


def native_profile_summary():
    """{tag: {ms, launches, flops, bytes}} of the tagged launches since native_profile(True) (synchronises the device)."""
    need = lib().ps_profile_dump(None, 0)
    if need < 0:
        raise NativeError("ps_profile_dump failed")
    buf = ctypes.create_string_buffer(int(need) + 16)
    n = lib().ps_profile_dump(buf, len(buf))
    out = {}
    for line in buf.value[: max(n, 0)].decode().splitlines():
        tag, ms, launches, flops, nbytes = line.split()
        out[tag] = {"ms": float(ms), "launches": int(launches), "flops": float(flops), "bytes": float(nbytes)}
    return out


PS_ERR_NOSPACE, PS_ERR_RANGE = -5, -6
_arena_hint = {}  # (B, T, n_layers, n_ids) -> bytes that were enough last time


def prepare_plan(batch, table_nodes, table_w, T, n_layers, need_backward=True, online=None):
    """ps_prepare_plan on the current stream: (arena uint8 tensor, PlanDescC).  The arena is sized from the previous call
    with the same shape and grown on demand (the sizes of the frontiers are only known on the device).
    online = (GraphHandle, n_items, n_hops, alpha, seed): ps_prepare_plan_online, the walker instead of the table."""
    global launch_count
    _ensure_device()
    B = batch.shape[0]
    n_ids = int(online[1]) if online is not None else table_nodes.shape[0]
    key = (B, T, n_layers, n_ids)
    size = _arena_hint.get(key, max(1 << 20, 64 * 3 * B * (T + 1) * T))
    desc = PlanDescC()
    while True:
        size = ((size + 256 + (1 << 24) - 1) >> 24 << 24) - 256  # 16 MB buckets: the caching allocator reuses the same blocks
        arena = torch.empty(size + 256, dtype=torch.uint8, device="cuda")
        base = (arena.data_ptr() + 255) & ~255
        launch_count += 6 + 14 * n_layers
        stream = c_void_p(torch.cuda.current_stream().cuda_stream)
        if online is not None:
            graph, _, n_hops, alpha, seed = online
            rc = lib().ps_prepare_plan_online(_p(batch, torch.int64), B, graph._h, n_ids, int(n_hops), float(alpha), int(seed),
                                              int(T), int(n_layers), int(need_backward), c_void_p(base), size, ctypes.byref(desc), stream)
        else:
            rc = lib().ps_prepare_plan(_p(batch, torch.int64), B, _p(table_nodes, torch.int32), _p(table_w, torch.float32), n_ids,
                                       table_nodes.shape[1], int(T), int(n_layers), int(need_backward), c_void_p(base), size,
                                       ctypes.byref(desc), stream)
        if rc == PS_ERR_NOSPACE:
            size = max(int(desc.bytes_needed), 2 * size)
            continue
        if rc == PS_ERR_RANGE:
            raise IndexError("node id out of range")  # the reference raises IndexError on OOB ids too
        check(rc)
        _arena_hint[key] = max(_arena_hint.get(key, 0), int(desc.bytes_used * 1.25) + 4096)
        return arena, base - arena.data_ptr(), desc


def train_step(args: "StepArgsC", launches: int):
    """ps_train_step on the current stream; `launches` = kernel launches it issues (bench.py's launch count)."""
    global launch_count
    launch_count += launches
    check(lib().ps_train_step(ctypes.byref(args), c_void_p(torch.cuda.current_stream().cuda_stream)))


class _Timed:
    __slots__ = ("tag", "flops", "nbytes", "e0")

    def __init__(self, tag, flops=0.0, nbytes=0.0):
        self.tag, self.flops, self.nbytes = tag, flops, nbytes

    def __enter__(self):
        if profiler is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if profiler is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            profiler.records.append((self.tag, self.e0, e1, self.flops, self.nbytes))
        return False


def _stream():
    global launch_count
    launch_count += 1
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _ld(t):
    return t.stride(0) if t.dim() == 2 else t.numel()


# ------------------------------------------------------------------------------------
# graph + walker
# ------------------------------------------------------------------------------------

class GraphHandle:
    """Owns the device CSR tensors and the opaque ps_graph_t*."""

    def __init__(self, indptr: torch.Tensor, indices: torch.Tensor, n_tracks: int, n_cols: int):
        _ensure_device()
        self.indptr = indptr.to(device="cuda", dtype=torch.int64).contiguous()
        self.indices = indices.to(device="cuda", dtype=torch.int32).contiguous()
        self.n_tracks, self.n_cols = int(n_tracks), int(n_cols)
        if self.indptr.numel() != n_tracks + n_cols + 1:
            raise NativeError("indptr must have n_tracks + n_cols + 1 entries")
        handle = c_void_p()
        check(lib().ps_graph_create(_p(self.indptr), _p(self.indices), self.n_tracks, self.n_cols,
                                    self.indices.numel(), ctypes.byref(handle), _stream()))
        self._h = handle

    def use_indptr32(self, on: bool) -> bool:
        """False forces the walker's 8-byte row-offset path (what graphs with >= 2^32 CSR entries take)."""
        return bool(lib().ps_graph_use_indptr32(self._h, int(bool(on))))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.ps_graph_destroy(h)


def walk_algo(mode: int) -> int:
    """0 = sort-based walker kernel where it applies (default), 1 = hash-table kernel always; returns the previous mode."""
    return lib().ps_walk_algo(int(mode))


def walk_topt(graph: GraphHandle, sources: torch.Tensor, n_hops: int, alpha: float, T: int, seed: int,
              fixed_len: int = 0, want_i64: bool = True, want_i32: bool = False, want_trace: bool = False):
    """ps_walk_topt.  Returns dict with the requested outputs."""
    _ensure_device()
    src = sources.to(device="cuda", dtype=torch.int64).contiguous()
    n = src.numel()
    out = {}
    if want_i64:
        out["nodes"] = torch.empty((n, T), dtype=torch.int64, device="cuda")
        out["weights"] = torch.empty((n, T), dtype=torch.float64, device="cuda")
    if want_i32:
        out["nodes_i32"] = torch.empty((n, T), dtype=torch.int32, device="cuda")
        out["weights_f32"] = torch.empty((n, T), dtype=torch.float32, device="cuda")
    if want_trace:
        out["trace"] = torch.empty((n, n_hops), dtype=torch.int32, device="cuda")
    check(lib().ps_walk_topt(graph._h, _p(src), n, int(n_hops), float(alpha), int(fixed_len), int(T),
                             int(seed) & 0xFFFFFFFFFFFFFFFF,
                             _p(out.get("nodes")), _p(out.get("weights")), _p(out.get("nodes_i32")),
                             _p(out.get("weights_f32")), _p(out.get("trace")), _stream()))
    return out


def trace_topt(trace: torch.Tensor, sources: torch.Tensor, T: int):
    """ps_trace_topt on a caller-supplied int64 trace [n, n_hops]."""
    _ensure_device()
    tr = trace.to(device="cuda", dtype=torch.int64).contiguous()
    src = sources.to(device="cuda", dtype=torch.int64).contiguous()
    n, n_hops = tr.shape
    nodes = torch.empty((n, T), dtype=torch.int64, device="cuda")
    w = torch.empty((n, T), dtype=torch.float64, device="cuda")
    check(lib().ps_trace_topt(_p(tr), _p(src), n, n_hops, int(T), _p(nodes), _p(w), None, None, _stream()))
    return w, nodes


# ------------------------------------------------------------------------------------
# dense + row-wise kernels
# ------------------------------------------------------------------------------------

def gemm(P, Q, C, M, N, K, *, p_kmajor=True, q_kmajor=True, p_rows=None, q_rows=None, bias=None,
         act=0, l2norm=False, norm_out=None, accumulate=False, splits=1, mask=None, tag="gemm"):
    """C[i,j] (+)= act(sum_r P(i,r) Q(j,r) + bias[j]); see ps_gemm / ps_gemm_ex in the header.  mask: int32
    [M, N/32] sign mask of the activation (written with act=1, read with act=2)."""
    with _Timed(tag, 2.0 * M * N * K, 4.0 * (M * K + N * K + M * N)):
        _gemm(P, Q, C, M, N, K, p_kmajor, q_kmajor, p_rows, q_rows, bias, act, l2norm, norm_out, accumulate, splits, mask)


def gemm_mask_supported(M, N, K) -> bool:
    return bool(lib().ps_gemm_mask_supported(int(M), int(N), int(K)))


def _gemm(P, Q, C, M, N, K, p_kmajor, q_kmajor, p_rows, q_rows, bias, act, l2norm, norm_out, accumulate, splits, mask=None):
    check(lib().ps_gemm_ex(_p(P, torch.float32), _ld(P), int(p_kmajor), _p(p_rows, torch.int32),
                           _p(Q, torch.float32), _ld(Q), int(q_kmajor), _p(q_rows, torch.int32),
                           _p(C, torch.float32), _ld(C), int(M), int(N), int(K), _p(bias, torch.float32),
                           int(act), int(l2norm), _p(norm_out, torch.float32), int(accumulate), int(splits),
                           _p(mask, torch.int32), _ld(mask) if mask is not None else 0, _stream()))


def gemm_wgrad(dY, X, dW, M, N, K, *, x_rows=None, splits=1, bias_grad=None, tag="gemm_wgrad"):
    """ps_gemm_wgrad: dW[M,N] += dY[:K,:M]^T X[rows][:K,:N] and bias_grad[M] += colsum(dY) in one call."""
    with _Timed(tag, 2.0 * M * N * K, 4.0 * (M * K + N * K + M * N)):
        check(lib().ps_gemm_wgrad(_p(dY, torch.float32), _ld(dY), _p(X, torch.float32), _ld(X), _p(x_rows, torch.int32),
                                  _p(dW, torch.float32), _ld(dW), int(M), int(N), int(K), int(splits),
                                  _p(bias_grad, torch.float32), _stream()))


def gemm_backend(mode: int) -> int:
    """0 = tcgen05 3xTF32 (default), 1 = CUDA-core fp32; returns the previous mode."""
    return lib().ps_gemm_backend(int(mode))


def gemm_tc_pack(on: int) -> int:
    """1 = pre-packed weight operands + bulk copies on the tensor-core path (default), 0 = off; returns the previous setting."""
    return lib().ps_gemm_tc_pack(int(on))


def aggregate_fwd(hin, self_rows, din, z, nbz, nbw, dh, cat, inv_wsum, tag="aggregate_fwd"):
    n, T = nbz.shape
    # algorithmic bytes: T gathered dh-wide rows + the self row, indices + weights, the [din+dh] output row
    with _Timed(tag, 2.0 * n * T * dh, float(n) * (T * dh * 4 + din * 4 + T * 8 + 4 + (din + dh) * 4 + 4)):
        _aggregate_fwd(hin, self_rows, din, z, nbz, nbw, dh, cat, inv_wsum, n, T)


def _aggregate_fwd(hin, self_rows, din, z, nbz, nbw, dh, cat, inv_wsum, n, T):
    check(lib().ps_aggregate_fwd(_p(hin, torch.float32), _ld(hin), _p(self_rows, torch.int32), int(din),
                                 _p(z, torch.float32), _ld(z), _p(nbz, torch.int32), _p(nbw, torch.float32),
                                 int(T), int(dh), int(n), _p(cat, torch.float32), _ld(cat),
                                 _p(inv_wsum, torch.float32), _stream()))


AGG_BWD_CHUNK = 64  # pairs per warp-sized work unit of the aggregation backward


def aggregate_bwd_chunks(seg_off, chunk_pairs=AGG_BWD_CHUNK):
    """chunk_off int32 [nz+1] for ps_aggregate_bwd (index arithmetic on the device, no sync)."""
    n_chunks = (seg_off[1:] - seg_off[:-1] + (chunk_pairs - 1)) // chunk_pairs
    chunk_off = torch.zeros_like(seg_off)
    chunk_off[1:] = torch.cumsum(n_chunks, 0)
    return chunk_off


def aggregate_bwd_chunk_rows(chunk_off, pairs, nz, chunk_pairs=AGG_BWD_CHUNK):
    """chunk_row int32 [max_chunks]: the z-row owning every chunk (device index arithmetic, no sync)."""
    max_chunks = pairs // chunk_pairs + nz
    q = torch.arange(max_chunks, dtype=torch.int32, device=chunk_off.device)
    return (torch.searchsorted(chunk_off, q, right=True) - 1).to(torch.int32)


def aggregate_bwd(dcat, col_off, dh, seg_off, pair_q, nbw, inv_wsum, T, z, chunk_off=None, chunk_pairs=AGG_BWD_CHUNK,
                  chunk_row=None, apply_leaky=True, tag="aggregate_bwd"):
    pairs, nz = pair_q.numel(), z.shape[0]
    if chunk_off is None:
        chunk_off = aggregate_bwd_chunks(seg_off, chunk_pairs)
    max_chunks = pairs // chunk_pairs + nz  # upper bound of chunk_off[-1], known without a device read
    ws = torch.empty((max(max_chunks, 1), dh), dtype=torch.float32, device=z.device)
    # algorithmic bytes: one dh-wide dcat row per (target, slot) pair + pair index/weight/inv_wsum, Z read + written
    with _Timed(tag, 2.0 * pairs * dh, float(pairs) * (dh * 4 + 12) + float(nz) * ((2 if apply_leaky else 1) * dh * 4 + 8)):
        check(lib().ps_aggregate_bwd(_p(dcat, torch.float32), _ld(dcat), int(col_off), int(dh),
                                     _p(seg_off, torch.int32), _p(chunk_off, torch.int32), int(chunk_pairs), int(max_chunks),
                                     _p(pair_q, torch.int32), _p(nbw, torch.float32), _p(inv_wsum, torch.float32), int(T),
                                     _p(z, torch.float32), _ld(z), int(nz), _p(ws, torch.float32),
                                     _p(chunk_row, torch.int32), int(apply_leaky), _stream()))


def norm_leaky_bwd(h, norm, dh, dpre):
    n, d = h.shape
    check(lib().ps_norm_leaky_bwd(_p(h, torch.float32), _ld(h), _p(norm, torch.float32), _p(dh, torch.float32), _ld(dh),
                                  _p(dpre, torch.float32), _ld(dpre), int(n), int(d), _stream()))


def l2norm_rows(x, norm_out):
    n, d = x.shape
    check(lib().ps_l2norm_rows(_p(x, torch.float32), _ld(x), int(n), int(d), _p(norm_out, torch.float32), _stream()))


def leaky_bwd(y, dy):
    assert y.is_contiguous() and dy.is_contiguous()
    check(lib().ps_leaky_bwd(_p(y, torch.float32), _p(dy, torch.float32), int(y.numel()), _stream()))


def colsum(x, out):
    n, d = x.shape
    check(lib().ps_colsum(_p(x, torch.float32), _ld(x), int(n), int(d), _p(out, torch.float32), _stream()))


def scatter_add_rows(src, rows, dst, d):
    check(lib().ps_scatter_add_rows(_p(src, torch.float32), _ld(src), _p(rows, torch.int32), _p(dst, torch.float32),
                                    _ld(dst), int(src.shape[0]), int(d), _stream()))


def count_triples(triples, U, dup_counts):
    check(lib().ps_count_triples(_p(triples, torch.int32), int(triples.shape[0]), int(U), _p(dup_counts, torch.int32), _stream()))


def margin_loss_fwd_bwd(emb, triples, margin, grad_scale, dup_counts, loss_out, demb):
    U, d = emb.shape
    check(lib().ps_margin_loss_fwd_bwd(_p(emb, torch.float32), _ld(emb), _p(triples, torch.int32), int(triples.shape[0]),
                                       int(d), float(margin), float(grad_scale), _p(dup_counts, torch.int32), int(U),
                                       _p(loss_out, torch.float32), _p(demb, torch.float32),
                                       _ld(demb) if demb is not None else 0, _stream()))


def plan_layer(nb, cur, with_self, n_ids):
    """ps_plan_layer: (uniq, nbz int32 [n,T], self_rows int32 [n] or None, size) of a layer's next frontier.  uniq is int64
    when with_self (it becomes the next layer's targets) and int32 otherwise (gather index of the Q transform).
    One host read (the size)."""
    _ensure_device()
    n, T = nb.shape
    cap = max(1, min(int(n_ids), n * T + (n if with_self else 0)))
    u64 = torch.empty(cap, dtype=torch.int64, device="cuda") if with_self else None
    u32 = None if with_self else torch.empty(cap, dtype=torch.int32, device="cuda")
    nbz = torch.empty((n, T), dtype=torch.int32, device="cuda")
    self_rows = torch.empty(n, dtype=torch.int32, device="cuda") if with_self else None
    count = torch.empty(1, dtype=torch.int32, device="cuda")
    check(lib().ps_plan_layer(_p(nb, torch.int32), int(n), int(T), _p(cur, torch.int64), int(with_self), int(n_ids),
                              _p(u64), _p(u32), _p(nbz), _p(self_rows), _p(count), _stream()))
    size = int(count)  # the layer's one host sync
    return (u64 if with_self else u32)[:size], nbz, self_rows, size


def plan_transpose(nbz, nz, chunk_pairs=AGG_BWD_CHUNK):
    """ps_plan_transpose: (pair_q, seg_off, chunk_off, chunk_row) for ps_aggregate_bwd; no host sync."""
    pairs = nbz.numel()
    max_chunks = pairs // chunk_pairs + nz
    pair_q = torch.empty(pairs, dtype=torch.int32, device="cuda")
    seg_off = torch.empty(nz + 1, dtype=torch.int32, device="cuda")
    chunk_off = torch.empty(nz + 1, dtype=torch.int32, device="cuda")
    chunk_row = torch.empty(max(max_chunks, 1), dtype=torch.int32, device="cuda")
    check(lib().ps_plan_transpose(_p(nbz, torch.int32), int(pairs), int(nz), int(chunk_pairs), _p(pair_q), _p(seg_off),
                                  _p(chunk_off), _p(chunk_row), int(max_chunks), _stream()))
    return pair_q, seg_off, chunk_off, chunk_row


def sample_batch(positives, all_ids, n_items, B, seed, step, out=None):
    """ps_sample_batch: int64 [B,3] (q, pos, neg) drawn on the device in one launch (no host sync)."""
    _ensure_device()
    if out is None:
        out = torch.empty((B, 3), dtype=torch.int64, device="cuda")
    ws = torch.empty(lib().ps_sample_batch_workspace(int(B)), dtype=torch.uint8, device="cuda")
    check(lib().ps_sample_batch(_p(positives, torch.int64), int(positives.shape[0]), _p(all_ids, torch.int64), int(n_items), int(B),
                                int(seed) & 0xFFFFFFFFFFFFFFFF, int(step) & 0xFFFFFFFFFFFFFFFF, _p(out, torch.int64),
                                _p(ws), ws.numel(), None, _stream()))
    return out


def sample_batch_supported(P, n_items, B):
    return 0 < B <= 2600 and 16 * B <= P < 0xFFFFFFFF and 16 * B <= n_items < (1 << 31)


def topk_rows(x, k):
    """ps_topk_rows: (values float32 [n, k], indices int64 [n, k]) of the k largest entries of every row of x."""
    _ensure_device()
    n, m = x.shape
    val = torch.empty((n, k), dtype=torch.float32, device="cuda")
    idx = torch.empty((n, k), dtype=torch.int64, device="cuda")
    check(lib().ps_topk_rows(_p(x, torch.float32), _ld(x), int(n), int(m), int(k), _p(val), _p(idx), _stream()))
    return val, idx


def gemm_filter(table, queries, thr, cap):
    """ps_gemm_filter: candidate lists (cnt int32 [nq], val float32 [nq, cap], row int32 [nq, cap]) of the table rows
    whose dot product with query j reaches thr[j].  Raises NativeError(-4) for shapes outside the packed path."""
    _ensure_device()
    nq, d = queries.shape
    cnt = torch.zeros(nq, dtype=torch.int32, device="cuda")
    val = torch.empty((nq, cap), dtype=torch.float32, device="cuda")
    row = torch.empty((nq, cap), dtype=torch.int32, device="cuda")
    with _Timed("gemm_knn_filter", 2.0 * table.shape[0] * nq * d, 4.0 * (table.numel() + queries.numel())):
        check(lib().ps_gemm_filter(_p(table, torch.float32), _ld(table), _p(queries, torch.float32), _ld(queries), int(table.shape[0]),
                                   int(nq), int(d), _p(thr, torch.float32), _p(cnt), _p(val), _p(row), int(cap), _stream()))
    return cnt, val, row


def topk_rows_mapped(val, col_ids, counts, k):
    """ps_topk_rows_mapped: exact top-k of per-row candidate lists -> (values [n, k], original column ids int64 [n, k])."""
    _ensure_device()
    n, cap = val.shape
    out_v = torch.empty((n, k), dtype=torch.float32, device="cuda")
    out_i = torch.empty((n, k), dtype=torch.int64, device="cuda")
    check(lib().ps_topk_rows_mapped(_p(val, torch.float32), _p(col_ids, torch.int32), _p(counts, torch.int32), int(cap), int(n), int(k),
                                    _p(out_v), _p(out_i), _stream()))
    return out_v, out_i


def csr_build(src, dst, n_nodes):
    """ps_csr_build: (indptr int64 [n_nodes + 1], indices int32 [E]) on the device from a directed edge list; edges of a
    source keep their listed order.  Raises IndexError on endpoints outside [0, n_nodes)."""
    _ensure_device()
    src = torch.as_tensor(src).to("cuda", torch.int64).contiguous()
    dst = torch.as_tensor(dst).to("cuda", torch.int64).contiguous()
    if src.shape != dst.shape or src.dim() != 1:
        raise ValueError("src / dst must be 1-D and of equal length")
    indptr = torch.empty(int(n_nodes) + 1, dtype=torch.int64, device="cuda")
    indices = torch.empty(src.numel(), dtype=torch.int32, device="cuda")
    rc = lib().ps_csr_build(_p(src), _p(dst), int(src.numel()), int(n_nodes), _p(indptr), _p(indices), _stream())
    if rc == PS_ERR_RANGE:
        raise IndexError("edge endpoint out of range")
    check(rc)
    return indptr, indices


def standardize_(x, eps=1e-12):
    """ps_standardize in place on a float32 [n, d] device tensor; returns (mean [d], std + eps [d])."""
    _ensure_device()
    if x.dim() != 2:
        raise ValueError("expected a [n, d] matrix")
    mean = torch.empty(x.shape[1], dtype=torch.float32, device="cuda")
    std = torch.empty(x.shape[1], dtype=torch.float32, device="cuda")
    check(lib().ps_standardize(_p(x, torch.float32), _ld(x), int(x.shape[0]), int(x.shape[1]), float(eps), _p(mean), _p(std), _stream()))
    return mean, std


def train_diagnostics(feats, batch, emb, triples, feat_margin, out2):
    """ps_train_diagnostics: out2 = [feature triplet loss, batch variance of the query embeddings]."""
    check(lib().ps_train_diagnostics(_p(feats, torch.float32), _ld(feats), int(feats.shape[1]), _p(batch, torch.int64), int(batch.shape[0]),
                                     _p(emb, torch.float32), _ld(emb), int(emb.shape[1]), _p(triples, torch.int32), float(feat_margin),
                                     _p(out2, torch.float32), _stream()))


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, grad_scale=1.0):
    check(lib().ps_adam_step(_p(param, torch.float32), _p(grad, torch.float32), _p(exp_avg, torch.float32),
                             _p(exp_avg_sq, torch.float32), int(param.numel()), float(lr), float(beta1), float(beta2),
                             float(eps), int(step), float(grad_scale), _stream()))
