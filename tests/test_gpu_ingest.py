"""Ingest on the device: ps_csr_build / ps_standardize against numpy / fp64 restatements
and against the reference loader's own output (tests/golden/dataset.npz, written by oracle/make_golden.py from
the unmodified reference's SpotifyGraph.to_dgl_graph, spotify_graph.py:41-85)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _csr_numpy(src, dst, n):
    order = np.argsort(src, kind="stable")
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=indptr[1:])
    return indptr, dst[order].astype(np.int32)


@pytest.mark.parametrize("n,e,seed", [(7, 0, 0), (50, 400, 1), (5000, 200_000, 2), (1_200_000, 3_000_000, 3)])
def test_csr_build_matches_stable_sort(n, e, seed):
    """Bit-exact: row offsets and the listed order of every row's successors (duplicates and empty rows included)."""
    import ps_native as nat
    rs = np.random.RandomState(seed)
    src = rs.randint(0, n, size=e).astype(np.int64)
    dst = rs.randint(0, n, size=e).astype(np.int64)
    if e:
        src[: e // 10] = src[0]  # one heavy row with repeated (src, dst) pairs
        dst[: e // 20] = dst[0]
    indptr, indices = nat.csr_build(torch.from_numpy(src), torch.from_numpy(dst), n)
    ri, rx = _csr_numpy(src, dst, n)
    assert np.array_equal(indptr.cpu().numpy(), ri)
    assert np.array_equal(indices.cpu().numpy(), rx)


def test_csr_build_rejects_out_of_range():
    import ps_native as nat
    src = torch.tensor([0, 1, 2], dtype=torch.int64)
    with pytest.raises(IndexError):
        nat.csr_build(src, torch.tensor([1, 2, 3], dtype=torch.int64), 3)
    with pytest.raises(IndexError):
        nat.csr_build(torch.tensor([0, -1, 2], dtype=torch.int64), src, 3)


@pytest.mark.parametrize("n,d", [(2, 4), (1000, 33), (200_000, 256)])
def test_standardize_matches_fp64(n, d):
    import ps_native as nat
    gen = torch.Generator().manual_seed(n + d)
    x = torch.randn((n, d), generator=gen) * torch.linspace(0.1, 30.0, d) + torch.linspace(-50.0, 50.0, d)
    x[:, 0] = 3.25  # a constant column: std = 0 -> division by eps, like the reference's (x - mean) / (0 + 1e-12)
    xd = x.cuda()
    mean, std = nat.standardize_(xd, 1e-12)
    x64 = x.double()
    ref_mean, ref_std = x64.mean(0), x64.std(0, unbiased=True)
    assert torch.allclose(mean.cpu().double(), ref_mean, rtol=1e-6, atol=1e-6)
    assert torch.allclose(std.cpu().double()[1:], ref_std[1:], rtol=1e-6)
    ref = ((x - ref_mean.float()) / (ref_std.float() + 1e-12))
    got = xd.cpu()
    assert torch.allclose(got[:, 1:], ref[:, 1:], rtol=1e-5, atol=1e-6)
    assert torch.equal(got[:, 0], torch.zeros(n))  # (3.25 - 3.25) / 1e-12


def test_spotify_graph_device_ingest_matches_reference_loader(golden, tmp_path):
    """The drop-in loader on a GPU box (CSR through ps_csr_build, features through ps_standardize) == the
    reference's loader on the same files."""
    from spotify_graph import SpotifyGraph
    from test_abi_and_host import _rewrite_dataset
    g = golden("dataset")
    d = str(tmp_path)
    _rewrite_dataset(g, d)
    ds = SpotifyGraph(d, os.path.join(d, "features_openl3"))
    graph, track_ids, col_ids, features = ds.to_dgl_graph()
    assert getattr(graph, "_dev_csr", None) is not None, "the CSR was not built on the device"
    assert np.array_equal(graph.indptr.numpy(), g["indptr"]) and np.array_equal(graph.indices.numpy(), g["indices"])
    assert np.allclose(features.numpy(), g["features"], rtol=1e-5, atol=1e-6)
    h = graph.device()  # adopts the device tensors; ps_graph_create validates degrees
    assert h.indices.numel() == g["indices"].size
