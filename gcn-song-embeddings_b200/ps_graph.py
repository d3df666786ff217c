"""CSR-backed bipartite track-collection graph: the adjacency object the engine uses in
place of the reference's DGLGraph.

The reference's hot path touches its graph only through `successors`,
`number_of_nodes` and two ad-hoc attributes `nbhds_path` / `base_dir`
(pinsage_model.py:41,44,93; pinsage_training.py:117; spotify_graph.py:52-55).  PSGraph
offers those (plus the few degree/edge accessors eval.py and baselines.py use) on top of
a CSR that is uploaded once to HBM and handed to libpinsage_b200 as raw pointers.
Node ids: tracks are [0, n_tracks), collections [n_tracks, n_tracks + n_cols)
(spotify_graph.py:43-46,58).  Multi-edges are kept, as DGL keeps them.
"""
from __future__ import annotations

import numpy as np
import torch

import ps_native


class PSGraph:
    def __init__(self, indptr, indices, n_tracks: int, n_cols: int, nbhds_path=None, base_dir=None):
        self.indptr = torch.as_tensor(indptr, dtype=torch.int64).cpu().contiguous()
        self.indices = torch.as_tensor(indices).to(torch.int32).cpu().contiguous()
        self.n_tracks, self.n_cols = int(n_tracks), int(n_cols)
        self.nbhds_path = nbhds_path
        self.base_dir = base_dir
        self._handle = None
        if self.indptr.numel() != self.n_tracks + self.n_cols + 1:
            raise ValueError("indptr must have n_tracks + n_cols + 1 entries")

    # ---- construction -------------------------------------------------------------
    @classmethod
    def from_edges(cls, src, dst, n_tracks, n_cols, **kw):
        """Directed edge list (both directions listed, as in graph.json) -> CSR.  Edges of a
        source keep their listed order (stable sort), matching DGL's insertion order.
        With a CUDA device the CSR is built in HBM by ps_csr_build (csrc/ingest.cu: one stable radix sort by
        source + row offsets) and the handle adopts those tensors; host-only tooling (the CPU tests, dataset
        inspection on a machine without a GPU) gets the same arrays from a numpy stable argsort."""
        n = n_tracks + n_cols
        if torch.cuda.is_available():
            indptr, indices = ps_native.csr_build(torch.as_tensor(src), torch.as_tensor(dst), n)
            g = cls(indptr.cpu(), indices.cpu(), n_tracks, n_cols, **kw)
            g._dev_csr = (indptr, indices)  # device() adopts them instead of uploading again
            return g
        src = np.asarray(src, dtype=np.int64)
        dst = np.asarray(dst, dtype=np.int64)
        if src.size and (src.min() < 0 or src.max() >= n or dst.min() < 0 or dst.max() >= n):
            raise IndexError("edge endpoint out of range")
        order = np.argsort(src, kind="stable")
        indptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.bincount(src, minlength=n), out=indptr[1:])
        return cls(indptr, dst[order].astype(np.int32), n_tracks, n_cols, **kw)

    @classmethod
    def from_dgl_like(cls, g, n_tracks):
        """Adopt any object with DGL's `edges()` / `number_of_nodes()` (e.g. a real DGLGraph)."""
        src, dst = g.edges()
        n = g.number_of_nodes()
        out = cls.from_edges(np.asarray(src), np.asarray(dst), n_tracks, n - n_tracks,
                             nbhds_path=getattr(g, "nbhds_path", None), base_dir=getattr(g, "base_dir", None))
        return out

    # ---- DGL-flavoured accessors ---------------------------------------------------
    def number_of_nodes(self):
        return self.n_tracks + self.n_cols

    def __len__(self):
        return self.number_of_nodes()

    def successors(self, v):
        v = int(v)
        return self.indices[self.indptr[v]:self.indptr[v + 1]].to(torch.int64)

    def predecessors(self, v):  # the graph is symmetric (both directions are listed)
        return self.successors(v)

    def out_degrees(self, v=None):
        deg = self.indptr[1:] - self.indptr[:-1]
        return deg if v is None else deg[v]

    def in_degrees(self, v=None):
        deg = torch.bincount(self.indices.to(torch.int64), minlength=self.number_of_nodes())
        return deg if v is None else deg[v]

    def edges(self):
        deg = self.indptr[1:] - self.indptr[:-1]
        src = torch.repeat_interleave(torch.arange(self.number_of_nodes()), deg)
        return src, self.indices.to(torch.int64)

    # ---- device side ----------------------------------------------------------------
    def device(self) -> "ps_native.GraphHandle":
        """Upload once; ps_graph_create validates that every node has a successor."""
        if self._handle is None:
            indptr, indices = getattr(self, "_dev_csr", None) or (self.indptr, self.indices)
            self._handle = ps_native.GraphHandle(indptr, indices, self.n_tracks, self.n_cols)
            self._dev_csr = None
        return self._handle


def as_psgraph(g, n_items) -> PSGraph:
    """Accept a PSGraph or any DGL-like graph (converted once and cached on the object)."""
    if isinstance(g, PSGraph):
        return g
    cached = getattr(g, "_ps_graph", None)
    if cached is None:
        cached = PSGraph.from_dgl_like(g, n_items)
        try:
            g._ps_graph = cached
        except Exception:
            pass
    return cached
