"""GPU parity tests of the dense / row-wise kernels through the C ABI, each against a
plain fp32/fp64 torch statement of the same op."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-4


@pytest.fixture(scope="module")
def nat():
    import ps_native
    return ps_native


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def leaky(x):
    return torch.nn.functional.leaky_relu(x, 0.01)


@pytest.fixture(params=["tcgen05", "simt"])
def backend(request, nat):
    """Run a GEMM test on the tcgen05 3xTF32 path (default dispatch) and on the CUDA-core path."""
    old = nat.gemm_backend(0 if request.param == "tcgen05" else 1)
    yield request.param
    nat.gemm_backend(old)


@pytest.mark.parametrize("M,N,K", [(1, 4, 4), (130, 36, 20), (257, 128, 64), (1000, 512, 256), (4096, 132, 772),
                                   (128, 128, 32), (128, 256, 32), (300, 64, 40), (5000, 768, 128)])
@pytest.mark.parametrize("pk,qk", [(True, True), (True, False), (False, True), (False, False)])
def test_gemm_layouts(nat, backend, M, N, K, pk, qk):
    torch.manual_seed(M * 7 + N)
    if not pk and M % 4:
        M = (M + 3) // 4 * 4
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
    P = A.contiguous() if pk else A.t().contiguous()
    Q = B.contiguous() if qk else B.t().contiguous()
    C = torch.empty(M, N, device="cuda")
    nat.gemm(P, Q, C, M, N, K, p_kmajor=pk, q_kmajor=qk)
    want = (A.double() @ B.double().t())
    assert rel(C, want) < 1e-5


@pytest.mark.parametrize("M,N,K", [(3000, 512, 128), (70, 64, 32), (4096, 96, 128)])
def test_gemm_act2_masks_by_existing_output(nat, backend, M, N, K):
    """act = 2: C holds leaky_relu outputs y and receives (P Q^T) * leaky'(y); Q given as a column block of a wider
    MN-major matrix (how the engine applies W[:, din:] in the aggregation backward)."""
    torch.manual_seed(M + N)
    S = torch.randn(M, K, device="cuda")
    W = torch.randn(K, 256 + N, device="cuda")            # [do, din + dh]
    y = leaky(torch.randn(M, N, device="cuda"))
    C = y.clone()
    nat.gemm(S, W[:, 256:], C, M, N, K, q_kmajor=False, act=2)
    want = (S.double() @ W[:, 256:].double()) * torch.where(y > 0, 1.0, 0.01).double()
    assert rel(C, want) < 1e-5


def test_gemm_sign_mask_roundtrip(nat):
    """ps_gemm_ex: act=1 records sign(output) as one bit per element; act=2 with that mask stores sum * leaky'
    without reading C (what the engine's forward / aggregation-backward pair does)."""
    torch.manual_seed(5)
    for M, N, K in ((2000, 512, 256), (130, 64, 32), (1500, 96, 64)):
        assert nat.gemm_mask_supported(M, N, K)
        X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
        y = torch.empty(M, N, device="cuda"); mask = torch.zeros(M, N // 32, dtype=torch.int32, device="cuda")
        nat.gemm(X, W, y, M, N, K, bias=b, act=1, mask=mask)
        bits = ((mask.view(M, N // 32, 1) >> torch.arange(32, device="cuda", dtype=torch.int32)) & 1).reshape(M, N).bool()
        assert torch.equal(bits, y > 0)
        S = torch.randn(M, 64, device="cuda"); V = torch.randn(64, N, device="cuda")
        out = torch.full((M, N), float("nan"), device="cuda")   # never read
        nat.gemm(S, V, out, M, N, 64, q_kmajor=False, act=2, mask=mask)
        want = (S.double() @ V.double()) * torch.where(y > 0, 1.0, 0.01).double()
        assert rel(out, want) < 1e-5
    assert not nat.gemm_mask_supported(100, 48, 64)


def test_gemm_gather_bias_act_norm(nat, backend):
    torch.manual_seed(1)
    table = torch.randn(5000, 256, device="cuda")
    rows = torch.randint(0, 5000, (777,), device="cuda", dtype=torch.int32)
    W = torch.randn(128, 256, device="cuda") * 0.1; b = torch.randn(128, device="cuda")
    out = torch.empty(777, 128, device="cuda"); norm = torch.empty(777, device="cuda")
    nat.gemm(table, W, out, 777, 128, 256, p_rows=rows, bias=b, act=1, l2norm=True, norm_out=norm)
    y = leaky(table[rows.long()].double() @ W.double().t() + b.double())
    assert rel(norm, y.norm(dim=1)) < 1e-5
    assert rel(out, y / y.norm(dim=1, keepdim=True)) < 1e-5
    # narrow output (N < tile) with the norm epilogue
    W2 = torch.randn(32, 256, device="cuda") * 0.1
    out2 = torch.empty(777, 32, device="cuda")
    nat.gemm(table, W2, out2, 777, 32, 256, p_rows=rows, act=1, l2norm=True)
    y2 = leaky(table[rows.long()].double() @ W2.double().t())
    assert rel(out2, y2 / y2.norm(dim=1, keepdim=True)) < 1e-5


@pytest.mark.parametrize("splits", [1, 7, 64])
def test_gemm_wgrad_accumulate(nat, backend, splits):
    """dW[n,k] += sum_m dY[m,n] X[rows[m],k]  (both operands MN-major, gather on the contraction index)."""
    torch.manual_seed(2)
    m = 10_000
    dY = torch.randn(m, 96, device="cuda"); X = torch.randn(3000, 64, device="cuda")
    rows = torch.randint(0, 3000, (m,), device="cuda", dtype=torch.int32)
    dW = torch.ones(96, 64, device="cuda")
    nat.gemm(dY, X, dW, 96, 64, m, p_kmajor=False, q_kmajor=False, q_rows=rows, accumulate=True, splits=splits)
    want = 1.0 + dY.double().t() @ X[rows.long()].double()
    assert rel(dW, want) < 2e-5
    # the bench-sized weight gradient: [512 x 256] over ~100k gathered rows
    m = 100_000
    dY = torch.randn(m, 512, device="cuda"); X = torch.randn(30_000, 256, device="cuda")
    rows = torch.randint(0, 30_000, (m,), device="cuda", dtype=torch.int32)
    dW = torch.zeros(512, 256, device="cuda")
    nat.gemm(dY, X, dW, 512, 256, m, p_kmajor=False, q_kmajor=False, q_rows=rows, accumulate=True, splits=max(splits, 2))
    assert rel(dW, dY.double().t() @ X[rows.long()].double()) < 2e-5


def test_gemm_tcgen05_is_fp32_accurate(nat):
    """The 3xTF32 split keeps near-fp32 accuracy (a plain TF32 product sits near 5e-4).  What is left is the
    tensor core's truncating TMEM accumulator (~2e-8 per MMA in the chain), bounded by the dispatcher's
    chain-length cap."""
    torch.manual_seed(11)
    A = torch.randn(2048, 1024, device="cuda"); B = torch.randn(512, 1024, device="cuda")
    want = A.double() @ B.double().t()
    errs = {}
    for name, mode in (("tcgen05", 0), ("simt", 1)):
        old = nat.gemm_backend(mode)
        C = torch.empty(2048, 512, device="cuda")
        nat.gemm(A, B, C, 2048, 512, 1024)
        nat.gemm_backend(old)
        errs[name] = rel(C, want)
    assert errs["simt"] < 2e-6 and errs["tcgen05"] < 2e-5, errs


def test_gemm_rejects_bad_shapes(nat):
    A = torch.randn(8, 6, device="cuda"); B = torch.randn(4, 6, device="cuda"); C = torch.empty(8, 4, device="cuda")
    with pytest.raises(nat.NativeError):
        nat.gemm(A, B, C, 8, 4, 6)  # K, ld not multiples of 4


@pytest.mark.parametrize("n,T,din,dh", [(1, 1, 4, 4), (333, 3, 64, 96), (1000, 50, 256, 512), (500, 10, 128, 1024), (64, 7, 32, 132)])
def test_aggregate_fwd_bwd(nat, n, T, din, dh):
    torch.manual_seed(n + T)
    n_in, nz = 2 * n + 5, 3 * n + 7
    hin = torch.randn(n_in, din, device="cuda")
    z = leaky(torch.randn(nz, dh, device="cuda"))
    self_rows = torch.randint(0, n_in, (n,), device="cuda", dtype=torch.int32)
    nbz = torch.randint(0, nz, (n, T), device="cuda", dtype=torch.int32)
    w = torch.randint(1, 40, (n, T), device="cuda").float() / 500
    cat = torch.empty(n, din + dh, device="cuda"); inv = torch.empty(n, device="cuda")
    nat.aggregate_fwd(hin, self_rows, din, z, nbz, w, dh, cat, inv)
    agg = (w.double()[:, :, None] * z[nbz.long()].double()).sum(1) / w.double().sum(1, keepdim=True)
    assert torch.equal(cat[:, :din], hin[self_rows.long()])
    assert rel(cat[:, din:], agg) < 1e-6
    assert rel(inv, 1 / w.double().sum(1)) < 1e-6
    # backward: dZ[u] = leaky'(z[u]) * sum_{(i,t): nbz=u} w/wsum * dcat[i, din:]
    dcat = torch.randn(n, din + dh, device="cuda")
    flat = nbz.reshape(-1)
    _, order = torch.sort(flat)
    seg = torch.zeros(nz + 1, dtype=torch.int32, device="cuda")
    seg[1:] = torch.cumsum(torch.bincount(flat, minlength=nz), 0)
    zz = z.clone()
    nat.aggregate_bwd(dcat, din, dh, seg, order.to(torch.int32), w, inv, T, zz)
    coef = (w.double() / w.double().sum(1, keepdim=True)).reshape(-1)
    contrib = coef[:, None] * dcat[:, din:].double().repeat_interleave(T, 0)
    want = torch.zeros(nz, dh, device="cuda", dtype=torch.float64).index_add_(0, flat.long(), contrib)
    want = want * torch.where(z > 0, 1.0, 0.01).double()
    assert rel(zz, want) < 1e-5


def test_aggregate_bwd_skewed_segments(nat):
    """Load-balanced backward: one z-row that owns thousands of pairs (split over many chunks), rows that own
    none, and everything in between; the result must not depend on the chunk size."""
    torch.manual_seed(9)
    n, T, din, dh, nz = 4000, 20, 64, 256, 3000
    z = leaky(torch.randn(nz, dh, device="cuda"))
    hot = torch.rand(n, T, device="cuda") < 0.3
    nbz = torch.where(hot, torch.zeros(n, T, device="cuda", dtype=torch.int64), torch.randint(1, nz // 2, (n, T), device="cuda")).to(torch.int32)
    w = torch.randint(1, 40, (n, T), device="cuda").float() / 500
    inv = (1 / w.sum(1)).contiguous()
    dcat = torch.randn(n, din + dh, device="cuda")
    flat = nbz.reshape(-1)
    _, order = torch.sort(flat)
    seg = torch.zeros(nz + 1, dtype=torch.int32, device="cuda")
    seg[1:] = torch.cumsum(torch.bincount(flat, minlength=nz), 0)
    coef = (w.double() * inv.double()[:, None]).reshape(-1)
    want = torch.zeros(nz, dh, device="cuda", dtype=torch.float64).index_add_(
        0, flat.long(), coef[:, None] * dcat[:, din:].double().repeat_interleave(T, 0))
    want = want * torch.where(z > 0, 1.0, 0.01).double()
    assert int(seg[1]) > 20000 and bool((want[nz // 2:] == 0).all())
    outs = []
    for chunk in (64, 7, 100000):
        zz = z.clone()
        nat.aggregate_bwd(dcat, din, dh, seg, order.to(torch.int32), w, inv, T, zz, chunk_pairs=chunk)
        assert rel(zz, want) < 1e-5
        assert bool((zz[nz // 2:] == 0).all())
        outs.append(zz)
    again = z.clone()
    nat.aggregate_bwd(dcat, din, dh, seg, order.to(torch.int32), w, inv, T, again, chunk_pairs=64)
    assert torch.equal(again, outs[0])  # deterministic (no atomics)
    # apply_leaky = False: the plain segmented sum into an output-only buffer (what the engine runs on d_pre)
    plain = torch.full((nz, dh), 7.0, device="cuda")
    nat.aggregate_bwd(dcat, din, dh, seg, order.to(torch.int32), w, inv, T, plain, chunk_pairs=64, apply_leaky=False)
    want_plain = torch.zeros(nz, dh, device="cuda", dtype=torch.float64).index_add_(
        0, flat.long(), coef[:, None] * dcat[:, din:].double().repeat_interleave(T, 0))
    assert rel(plain, want_plain) < 1e-5 and bool((plain[nz // 2:] == 0).all())
    # precomputed chunk -> row map (what the engine passes) == the in-kernel search, bit for bit
    mapped = z.clone()
    coff = nat.aggregate_bwd_chunks(seg, 64)
    crow = nat.aggregate_bwd_chunk_rows(coff, flat.numel(), nz, 64)
    nat.aggregate_bwd(dcat, din, dh, seg, order.to(torch.int32), w, inv, T, mapped, chunk_off=coff, chunk_pairs=64, chunk_row=crow)
    assert torch.equal(mapped, outs[0])


def test_rowwise_kernels(nat):
    torch.manual_seed(3)
    n, d = 1234, 128
    pre = torch.randn(n, d, device="cuda", dtype=torch.float64, requires_grad=True)
    y = leaky(pre); h = y / y.norm(dim=1, keepdim=True)
    g = torch.randn(n, d, device="cuda", dtype=torch.float64)
    (h * g).sum().backward()
    dpre = torch.empty(n, d, device="cuda")
    nat.norm_leaky_bwd(h.detach().float(), y.detach().norm(dim=1).float(), g.float(), dpre)
    assert rel(dpre, pre.grad) < 1e-5
    # leaky_bwd
    yy = leaky(torch.randn(n, d, device="cuda")); dy = torch.randn(n, d, device="cuda"); ref = dy * torch.where(yy > 0, 1.0, 0.01)
    nat.leaky_bwd(yy, dy)
    assert torch.allclose(dy, ref)
    # colsum (accumulating)
    x = torch.randn(5001, 772, device="cuda"); out = torch.ones(772, device="cuda")
    nat.colsum(x, out)
    assert rel(out, 1 + x.double().sum(0)) < 1e-5
    # scatter_add_rows into a strided source view
    src = torch.randn(300, 200, device="cuda"); dst = torch.randn(1000, 64, device="cuda"); ref = dst.clone()
    rows = torch.randperm(1000, device="cuda")[:300].to(torch.int32)
    ref[rows.long()] += src[:, :64]
    nat.scatter_add_rows(src, rows, dst, 64)
    assert torch.allclose(dst, ref)
    # l2norm_rows
    x = torch.randn(77, 512, device="cuda"); ref = x / x.norm(dim=1, keepdim=True); nrm = torch.empty(77, device="cuda")
    nat.l2norm_rows(x, nrm)
    assert rel(x, ref) < 1e-6


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_margin_loss_vs_reference(nat, golden, tag):
    """K10 against the reference's max_margin_loss values and autograd gradients."""
    g = golden("loss")
    q, p, n = (torch.tensor(g[f"{tag}_{nm}"], device="cuda") for nm in "qpn")
    B, d = q.shape
    emb = torch.cat([q, p, n], 0).contiguous()
    idx = torch.arange(B, device="cuda", dtype=torch.int32)
    triples = torch.stack([idx, idx + B, idx + 2 * B], 1).contiguous()
    loss = torch.zeros(1, device="cuda"); demb = torch.zeros_like(emb)
    nat.margin_loss_fwd_bwd(emb, triples, float(g[f"{tag}_margin"]), 1.0, None, loss, demb)
    assert abs(float(loss) - float(g[f"{tag}_loss"])) <= 1e-5 * max(1.0, abs(float(g[f"{tag}_loss"])))
    want = torch.tensor(np.concatenate([g[f"{tag}_dq"], g[f"{tag}_dp"], g[f"{tag}_dn"]]), device="cuda")
    assert rel(demb, want) < RTOL


def test_margin_loss_shared_rows_and_dup_factor(nat):
    """Triples that index shared embedding rows; duplicate counts multiply the gradient."""
    torch.manual_seed(5)
    U, d, B = 40, 64, 64
    emb = torch.randn(U, d, device="cuda", dtype=torch.float64, requires_grad=True)
    triples = torch.randint(0, U, (B, 3), device="cuda")
    counts = torch.empty(3, U, dtype=torch.int32, device="cuda")
    nat.count_triples(triples.to(torch.int32), U, counts)
    for j in range(3):
        assert torch.equal(counts[j].long(), torch.bincount(triples[:, j], minlength=U))
    norm = torch.nn.functional.normalize
    hq, hp, hn = (norm(emb[triples[:, j]], dim=1) for j in range(3))
    dsum = (hq * hn).sum(1) - (hq * hp).sum(1) + 0.1
    lossv = torch.clamp(dsum, min=0).mean()
    # reference-compat gradient: each occurrence scaled by its column count
    gq, gp, gn = torch.autograd.grad(lossv, [hq, hp, hn], retain_graph=True)
    rows = []
    for j, (hx, gx) in enumerate(((hq, gq), (hp, gp), (hn, gn))):
        k = counts[j].double()[triples[:, j]][:, None]
        rows.append(torch.autograd.grad(hx, emb, gx * k, retain_graph=True)[0])
    want = sum(rows)
    loss = torch.zeros(1, device="cuda"); demb = torch.zeros(U, d, device="cuda")
    nat.margin_loss_fwd_bwd(emb.detach().float(), triples.to(torch.int32), 0.1, 1.0, counts, loss, demb)
    assert abs(float(loss) - float(lossv)) < 1e-5
    assert rel(demb, want) < RTOL


def test_adam_matches_torch(nat):
    torch.manual_seed(6)
    p = torch.randn(10_001, device="cuda"); ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=1e-3)
    m = torch.zeros_like(p); v = torch.zeros_like(p)
    for step in range(1, 6):
        g = torch.randn_like(p)
        ref.grad = g.clone(); opt.step()
        nat.adam_step(p, g, m, v, 1e-3, 0.9, 0.999, 1e-8, step)
    assert torch.allclose(p, ref.data, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("B,P,n_items", [(128, 5000, 4324), (1024, 200_000, 50_000), (2600, 60_000, 41_600)])
def test_sample_batch_bit_exact_vs_oracle(B, P, n_items):
    """K12: the one-launch device batch sampler equals its CPU restatement bit for bit and has the properties the
    reference's sample_batch guarantees (distinct positive rows, distinct negatives outside the pairs)."""
    import ps_native as nat
    from oracle import oracle
    rng = np.random.RandomState(B)
    positives = rng.randint(0, n_items, size=(P, 2)).astype(np.int64)
    pos_d = torch.from_numpy(positives).cuda()
    ids = torch.arange(n_items, device="cuda")
    for step in (1, 2, 77):
        got = nat.sample_batch(pos_d, ids, n_items, B, seed=0xABCDEF12345, step=step).cpu().numpy()
        want = oracle.sample_batch_philox(positives, n_items, B, 0xABCDEF12345, step)
        assert np.array_equal(got, want)
        oracle.check_batch_properties(got, positives, n_items)
    a = nat.sample_batch(pos_d, ids, n_items, B, seed=1, step=1).cpu().numpy()
    b = nat.sample_batch(pos_d, None, n_items, B, seed=1, step=2).cpu().numpy()
    assert not np.array_equal(a, b)


def test_sample_batch_uniform():
    """Negatives and positive rows are uniform: chi-square over 64 buckets on 200 batches."""
    import ps_native as nat
    n_items, P, B = 65_536, 131_072, 1024
    positives = torch.stack([torch.arange(P) % n_items, (torch.arange(P) * 7 + 3) % n_items], 1).cuda()
    ids = torch.arange(n_items, device="cuda")
    neg = np.zeros(64); row = np.zeros(64)
    for step in range(200):
        b = nat.sample_batch(positives, ids, n_items, B, seed=5, step=step).cpu().numpy()
        neg += np.bincount(b[:, 2] * 64 // n_items, minlength=64)
        row += np.bincount((b[:, 0] % 64), minlength=64)  # q = row index mod n_items: buckets of the row index
    for h in (neg, row):
        exp = h.sum() / 64
        assert ((h - exp) ** 2 / exp).sum() < 130  # chi2(63) 99.99th percentile ~ 115


@pytest.mark.parametrize("M,N,K,splits,gather", [(512, 256, 20_000, 74, True), (128, 768, 13_375, 53, False), (128, 128, 741, 3, False),
                                                 (96, 64, 10_000, 7, True), (36, 20, 300, 1, False)])
def test_gemm_wgrad_with_bias_gradient(nat, backend, M, N, K, splits, gather):
    """ps_gemm_wgrad: dW += dY^T X[rows] and db += colsum(dY) in one call (AddmmBackward of nn.Linear,
    pinsage_model.py:201,208,259); the bias gradient comes out of the GEMM's own operand tiles on the tcgen05 path."""
    torch.manual_seed(M + K)
    dY = torch.randn(K, M, device="cuda")
    X = torch.randn(3000 if gather else K, N, device="cuda")
    rows = torch.randint(0, 3000, (K,), device="cuda", dtype=torch.int32) if gather else None
    dW = torch.full((M, N), 0.5, device="cuda"); db = torch.full((M,), -1.0, device="cuda")
    nat.gemm_wgrad(dY, X, dW, M, N, K, x_rows=rows, splits=splits, bias_grad=db)
    Xg = X[rows.long()] if gather else X
    assert rel(dW, 0.5 + dY.double().t() @ Xg.double()) < 2e-5
    assert rel(db, -1.0 + dY.double().sum(0)) < 2e-5
    dW2 = torch.zeros((M, N), device="cuda")
    nat.gemm_wgrad(dY, X, dW2, M, N, K, x_rows=rows, splits=splits)  # bias gradient not wanted
    assert rel(dW2, dY.double().t() @ Xg.double()) < 2e-5


def test_train_diagnostics_match_torch(nat):
    """ps_train_diagnostics == the reference's per-step diagnostics (pinsage_training.py:200-212): the cosine triplet
    loss of the raw batch features and batch_variance of the query embeddings."""
    import pinsage_training as pst
    torch.manual_seed(3)
    for B, din, do in ((128, 512, 128), (1024, 256, 128), (7, 64, 32)):
        feats = torch.randn(5000, din, device="cuda")
        batch = torch.randint(0, 5000, (B, 3), device="cuda")
        emb = torch.randn(3 * B, do, device="cuda")
        triples = torch.randint(0, 3 * B, (B, 3), device="cuda", dtype=torch.int32)
        out = torch.empty(2, device="cuda")
        nat.train_diagnostics(feats, batch, emb, triples, pst.COSINE_TRIPLET_LOSS.margin, out)
        norm = torch.nn.functional.normalize
        want_l = pst.COSINE_TRIPLET_LOSS(norm(feats[batch[:, 0]], dim=1), norm(feats[batch[:, 1]], dim=1), norm(feats[batch[:, 2]], dim=1))
        want_v = pst.batch_variance(emb[triples[:, 0].long()])
        assert abs(float(out[0]) - float(want_l)) < 1e-5 * max(1.0, abs(float(want_l)))
        assert abs(float(out[1]) - float(want_v)) < 1e-4 * abs(float(want_v))


def test_flat_adam_equals_torch_adam():
    """ps_optim.FlatAdam (one fused kernel per step) == torch.optim.Adam step for step, with torch's state_dict format
    in both directions (the reference checkpoints `optimizer.state_dict()`, pinsage_training.py:288-295) and the
    scheduler driving param_groups[0]['lr']."""
    from ps_optim import FlatAdam
    torch.manual_seed(0)
    a = torch.nn.Sequential(torch.nn.Linear(64, 96), torch.nn.Linear(96, 8, bias=False)).cuda()
    b = torch.nn.Sequential(torch.nn.Linear(64, 96), torch.nn.Linear(96, 8, bias=False)).cuda()
    b.load_state_dict(a.state_dict())
    oa, ob = torch.optim.Adam(a.parameters(), lr=1e-3), FlatAdam(b.parameters(), lr=1e-3)
    sa, sb = torch.optim.lr_scheduler.ExponentialLR(oa, 0.9), torch.optim.lr_scheduler.ExponentialLR(ob, 0.9)
    for step in range(6):
        x = torch.randn(32, 64, device="cuda")
        for m, o in ((a, oa), (b, ob)):
            o.zero_grad(); m(x).pow(2).mean().backward(); o.step()
        if step % 2:
            sa.step(); sb.step()
        if step == 2:  # checkpoint round trip through torch's format, both ways
            sd_b, sd_a = ob.state_dict(), oa.state_dict()
            assert set(sd_b) == set(sd_a) and set(sd_b["state"][0]) == set(sd_a["state"][0])
            ob2 = FlatAdam(b.parameters(), lr=1e-3); ob2.load_state_dict(sd_b)
            oa2 = torch.optim.Adam(a.parameters(), lr=1e-3); oa2.load_state_dict(sd_b)   # torch reads ours
            ob2.load_state_dict(sd_a)                                                    # we read torch's
            assert ob2._t == 3 and rel(ob2._flat_m, ob._flat_m) < 1e-6
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert rel(pb, pa) < 1e-6
    assert oa.param_groups[0]["lr"] == ob.param_groups[0]["lr"]


@pytest.mark.parametrize("M,N,K", [(40_001 // 4 * 4, 512, 256), (38_000, 128, 768), (37_900, 512, 128), (75_000, 256, 64)])
def test_gemm_tile_pairs_with_multicast_weights(nat, M, N, K):
    """Tall packed-weight GEMMs run as 2-CTA clusters on tile PAIRS (odd tile counts leave a phantom second tile), either
    sharing the weight stream by TMA multicast (cluster mode 1) or on one tcgen05.mma.cta_group::2 with TMA tensor-map stores
    (modes 2 / 3, the default): forward with row gather + bias + leaky + sign mask, the act=2 backward with that mask, the
    l2norm epilogue -- against fp64 products, and bit-equal across all modes incl. the one-CTA-per-tile kernel."""
    torch.manual_seed(M + N)
    table = torch.randn(M + 5000, K, device="cuda")
    rows = torch.randint(0, M + 5000, (M,), device="cuda", dtype=torch.int32)
    W = torch.randn(N, K, device="cuda") * 0.1; b = torch.randn(N, device="cuda")
    outs = {}
    for cl in (1, 0, 2, 3):
        old = nat.lib().ps_gemm_tc_cluster(cl)
        y = torch.empty(M, N, device="cuda"); mask = torch.zeros(M, N // 32, dtype=torch.int32, device="cuda")
        nat.gemm(table, W, y, M, N, K, p_rows=rows, bias=b, act=1, mask=mask)
        S = torch.randn(M, 64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)); V = torch.randn(64, N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
        d = torch.full((M, N), float("nan"), device="cuda")
        nat.gemm(S, V, d, M, N, 64, q_kmajor=False, act=2, mask=mask)
        out = None
        if N <= 256:
            out = torch.empty(M, N, device="cuda"); norm = torch.empty(M, device="cuda")
            nat.gemm(table, W, out, M, N, K, p_rows=rows, bias=b, act=1, l2norm=True, norm_out=norm)
        outs[cl] = (y, mask, d, out)
        nat.lib().ps_gemm_tc_cluster(old)
    y, mask, d, out = outs[1]
    want = leaky(table[rows.long()].double() @ W.double().t() + b.double())
    assert rel(y, want) < 1e-5
    bits = ((mask.view(M, N // 32, 1) >> torch.arange(32, device="cuda", dtype=torch.int32)) & 1).reshape(M, N).bool()
    assert torch.equal(bits, y > 0)
    assert rel(d, (S.double() @ V.double()) * torch.where(y > 0, 1.0, 0.01).double()) < 1e-5
    if out is not None:
        assert rel(out, want / want.norm(dim=1, keepdim=True)) < 1e-5
    for other in (0, 2, 3):
        for a_, b_ in zip(outs[1], outs[other]):
            assert (a_ is None and b_ is None) or torch.equal(a_, b_)   # same arithmetic per output element


@pytest.mark.parametrize("M,N,K,gather", [(512, 256, 300_000, True), (384, 256, 150_001 // 4 * 4, False), (1024, 512, 90_000, True)])
def test_gemm_wgrad_on_tile_pairs(nat, M, N, K, gather):
    """Weight-gradient GEMMs (activation x activation, split-K, bias gradient from the operand tiles) on tile pairs with
    cta_group::2 (cluster mode 3, the default) against fp64 and against the one-CTA kernels (mode 1): same products, the
    split-K partials meet in atomics, so equal to rounding."""
    torch.manual_seed(K)
    dY = torch.randn(K, M, device="cuda") * 0.1
    X = torch.randn(K + 1000, N, device="cuda")
    rows = torch.randint(0, K + 1000, (K,), device="cuda", dtype=torch.int32) if gather else None
    Xg = X[rows.long()] if gather else X[:K]
    want = dY.double().t() @ Xg.double()
    want_b = dY.double().sum(0)
    got = {}
    for cl in (1, 3):
        old = nat.lib().ps_gemm_tc_cluster(cl)
        g = torch.zeros(M, N, device="cuda"); gb = torch.zeros(M, device="cuda")
        nat.gemm_wgrad(dY, X, g, M, N, K, x_rows=rows, splits=-(-148 * 4 // (-(-M // 128) * -(-N // 128))), bias_grad=gb)
        nat.lib().ps_gemm_tc_cluster(old)
        assert rel(g, want) < 2e-6 and rel(gb, want_b) < 2e-6, (cl, rel(g, want), rel(gb, want_b))
        got[cl] = (g, gb)
    scale = float(want.abs().max())
    assert float((got[1][0] - got[3][0]).abs().max()) < 1e-5 * scale
