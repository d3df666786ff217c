"""GPU tests of the evaluation side: device cosine kNN against the
reference's knn_from_emb golden, and the dashboard train -> save embeddings -> kNN -> HR/MRR table flow on a
small dataset written in the reference's on-disk schema."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_knn_from_emb_matches_reference(golden):
    """ps_knn.knn_from_emb == baselines.knn_from_emb (baselines.py:91-103) on the reference-generated golden."""
    import ps_knn
    g = golden("metrics_knn")
    emb = torch.from_numpy(g["knn_emb"])
    w, n = ps_knn.knn_from_emb(emb, torch.arange(300), 10)
    assert not w.is_cuda and n.dtype == torch.int64
    assert np.array_equal(n.numpy(), g["knn_n"])
    assert np.allclose(w.numpy(), g["knn_w"], rtol=1e-5, atol=1e-6)
    # query tiles smaller than the query set, device-resident inputs
    w2, n2 = ps_knn.knn_from_emb(emb.cuda(), torch.arange(300).cuda(), 10, q_tile=128)
    assert w2.is_cuda and np.array_equal(n2.cpu().numpy(), g["knn_n"])


def test_knn_large_self_first():
    """Size-independent properties at a size the oracle would not finish: every query's own row is dropped
    (rank 0), similarities are sorted descending and lie in [-1, 1]."""
    import ps_knn
    torch.manual_seed(0)
    emb = torch.randn(50_000, 128, device="cuda")
    q = torch.randint(0, 50_000, (2048,), device="cuda")
    w, n = ps_knn.knn_from_emb(emb, q, 100)
    assert n.shape == (2048, 100) and not (n == q[:, None]).any()
    assert (w[:, :-1] >= w[:, 1:]).all() and w.max() <= 1 + 1e-5 and w.min() >= -1 - 1e-5
    ref = torch.nn.functional.cosine_similarity(emb[q[:64], None, :], emb[n[:64]], dim=2)
    assert torch.allclose(ref, w[:64], rtol=1e-4, atol=1e-5)


def test_dashboard_train_eval_flow(tmp_path, monkeypatch):
    """dashboard.train_pinsage / eval_baselines (dashboard.py:48-172, PinSage rows) end to end on the device."""
    import ps_synth
    import dashboard
    import pinsage_training as pt
    d = str(tmp_path / "dataset")
    ps_synth.write_dataset(d, 400, 60, 4000, 128, 3000, seed=11)
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(pt, "BASE_RUN_DIR", str(tmp_path / "runs"))
    trainer = dashboard.train_pinsage(d, "features_openl3", "positives_lfm.json", run_name="t",
                                      epochs=1, b_per_e=20, batch_size=64, load_save=False)
    emb_dir = tmp_path / "runs" / "t" / "emb"
    files = sorted(os.listdir(emb_dir))
    assert len(files) == 400
    e0 = torch.load(emb_dir / files[0])
    assert e0.shape == (trainer.out_dim,) and e0.dtype == torch.float32
    table = dashboard.eval_baselines(d, "features_openl3", "positives_lfm.json", run_name="t",
                                     save_dir=str(tmp_path / "eval_cache"), k=50)
    row = table.loc["PinsageBase"]
    assert 0.0 <= row["hr (k=10)"] <= row["hr (k=100)"] <= 1.0 and 0.0 < row["mrr"] <= 1.0
    # the cached kNN lists are the reference's 5-tuple
    knn_w, knn_n, *_times = torch.load(tmp_path / "eval_cache" / "knn" / "PinsageBase.pt")
    assert knn_n.shape == (400, 50)
