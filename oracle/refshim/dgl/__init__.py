"""Minimal stand-in for the `dgl` package, used ONLY by oracle/make_golden.py in the
dev container to import the unmodified reference (`/root/reference/pinsage_model.py`,
`pinsage_training.py`) and generate golden vectors.  Test infrastructure, never shipped
on the product path.  The reference's hot path uses DGL purely as an adjacency lookup
(`successors`, `number_of_nodes`; reference pinsage_model.py:41,44,93 and
spotify_graph.py:48-63), so a CSR-backed object is sufficient."""
import numpy as np
import torch


class DGLGraph:
    def __init__(self):
        self._n = 0
        self._src = np.zeros(0, dtype=np.int64)
        self._dst = np.zeros(0, dtype=np.int64)
        self._csr = None

    def add_nodes(self, n):
        self._n += int(n)
        self._csr = None

    def add_edges(self, u, v):
        self._src = np.concatenate([self._src, np.asarray(u, dtype=np.int64)])
        self._dst = np.concatenate([self._dst, np.asarray(v, dtype=np.int64)])
        self._csr = None

    def _build(self):
        if self._csr is None:
            order = np.argsort(self._src, kind="stable")  # keep insertion order per source
            dst = self._dst[order]
            counts = np.bincount(self._src, minlength=self._n)
            indptr = np.zeros(self._n + 1, dtype=np.int64)
            np.cumsum(counts, out=indptr[1:])
            self._csr = (indptr, torch.from_numpy(dst.copy()))
        return self._csr

    def successors(self, v):
        indptr, dst = self._build()
        v = int(v)
        return dst[indptr[v]:indptr[v + 1]]

    def predecessors(self, v):
        v = int(v)
        return torch.from_numpy(self._src[self._dst == v].copy())

    def number_of_nodes(self):
        return self._n

    def __len__(self):
        return self._n

    def out_degrees(self, v=None):
        indptr, _ = self._build()
        deg = torch.from_numpy(np.diff(indptr))
        return deg if v is None else deg[v]

    def in_degrees(self, v=None):
        deg = torch.from_numpy(np.bincount(self._dst, minlength=self._n))
        return deg if v is None else deg[v]

    def edges(self):
        return torch.from_numpy(self._src.copy()), torch.from_numpy(self._dst.copy())

    def csr(self):
        indptr, dst = self._build()
        return indptr.copy(), dst.numpy().copy()
