"""GPU parity of the widened rows: PQ/OPQ encode, PQ training, EMA update, the device beam
search and the --eval_all_documents streaming top-k — against reference golden vectors and the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import oracle
from gpu_util import assert_topk_equivalent, ctx, dev

pytestmark = pytest.mark.gpu
MODES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "modes")


def g(name):
    return np.load(os.path.join(MODES, name))


@pytest.fixture(scope="module")
def data():
    import datasets

    return datasets.case_docs("small64"), datasets.make_queries(64)


# ---- PQ / OPQ encode ------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric", ["l2", "ip"])
def test_pq_encode_matches_reference_golden(data, metric):
    X, _ = data
    cb = g("pq_codebook.npy")
    codes = ctx().pq_encode(dev(X), dev(cb), metric=metric).cpu().numpy()
    ref = g(f"pq_codes_{metric}.npy")
    assert codes.dtype == np.int32 and codes.shape == ref.shape
    ties, real = oracle.classify_pq_mismatches(X, cb, codes, ref, metric)
    assert real == 0 and ties <= 2, (ties, real)


@pytest.mark.parametrize("n,d,M,bits", [(5000, 768, 32, 8), (3333, 768, 4, 5), (1000, 128, 8, 6), (37, 96, 6, 3), (0, 64, 4, 4)])
def test_pq_encode_matches_oracle_shapes(n, d, M, bits):
    rs = np.random.RandomState(n + d)
    K = 2 ** bits
    X = rs.standard_normal((n, d)).astype(np.float32)
    cb = rs.standard_normal((M, K, d // M)).astype(np.float32)
    for metric in ("l2", "ip"):
        codes = ctx().pq_encode(dev(X), dev(cb), metric=metric).cpu().numpy()
        ref = oracle.pq_encode(X, cb, metric) if n else np.zeros((0, M), np.int32)
        ties, real = oracle.classify_pq_mismatches(X, cb, codes, ref, metric)
        assert real == 0 and ties <= max(2, n * M // 5000), (metric, ties, real)


@pytest.mark.parametrize("n,d,M,bits,metric", [(40000, 768, 4, 5, "l2"), (40000, 768, 4, 5, "ip"), (20011, 256, 2, 6, "l2"),
                                               (9000, 128, 2, 5, "l2")])
def test_pq_encode_tensor_route_equals_subvector_kernel(monkeypatch, n, d, M, bits, metric):
    """M*K <= 128 and n >= 4096: mevi_pq_encode runs the RQ tensor kernel on the block-padded codebook (default) with the
    sub-vector kernel as the arbiter of flagged rows.  Codes must equal the oracle's up to fp32 ties (float64 arbiter
    per sub-vector) and the sub-vector kernel's own codes almost everywhere."""
    rs = np.random.RandomState(n + d + M)
    K = 2 ** bits
    X = rs.standard_normal((n, d)).astype(np.float32)
    cb = (rs.standard_normal((M, K, d // M)) * 0.7).astype(np.float32)
    c = ctx()
    l0 = c.launches
    codes_t = c.pq_encode(dev(X), dev(cb), metric=metric).cpu().numpy()
    assert c.launches - l0 >= 8  # the tensor route's preparation kernels ran (the sub-vector path is ONE launch)
    monkeypatch.setenv("MEVI_PQ_TENSOR", "0")
    l0 = c.launches
    codes_s = c.pq_encode(dev(X), dev(cb), metric=metric).cpu().numpy()
    assert c.launches - l0 == 1
    monkeypatch.delenv("MEVI_PQ_TENSOR")
    ref = oracle.pq_encode(X, cb, metric)
    for got in (codes_t, codes_s):
        ties, real = oracle.classify_pq_mismatches(X, cb, got, ref, metric)
        assert real == 0 and ties <= max(2, n * M // 5000), (metric, ties, real)
    assert (codes_t != codes_s).any(1).mean() < 1e-3


@pytest.mark.parametrize("n,M,ds,metric,kind", [(30011, 24, 32, "l2", "gauss"), (30011, 24, 32, "ip", "gauss"), (4096, 32, 32, "l2", "gauss"),
                                                (50000, 24, 32, "l2", "rows"), (20000, 24, 32, "l2", "scaled"), (9000, 3, 32, "l2", "gauss"),
                                                (30011, 32, 24, "l2", "gauss"), (20000, 32, 24, "ip", "gauss"), (40000, 32, 24, "l2", "rows"),
                                                (8000, 5, 24, "l2", "scaled")])
def test_pq_encode_wide_codebook_tensor_kernel(monkeypatch, n, M, ds, metric, kind):
    """K = 256 centroids per sub-vector of width 32 or 24 (24 x 256 and the reference default 32 x 256 at d = 768):
    mevi_pq_encode runs pq_tensor_kernel (split-fp16 tcgen05 contraction per sub-vector, (row, sub-vector) pairs inside the
    error bound re-decided by the fp32 pair arbiter).  Codes must be IDENTICAL to the sub-vector kernel's (the arbiter
    repeats its arithmetic) and equal the oracle's up to fp32 ties.  `rows`: centroids are data rows (zero distances,
    duplicates of a row among the centroids); `scaled`: 1e-3-scale data with an offset."""
    rs = np.random.RandomState(n + M)
    d = ds * M
    X = rs.standard_normal((n, d)).astype(np.float32)
    if kind == "scaled":
        X = (X * 1e-3 + 0.25).astype(np.float32)
    if kind == "rows":
        pick = rs.randint(0, n, size=(M, 256))
        pick[:, 7] = pick[:, 3]  # a duplicated centroid: exact ties, lowest index must win
        cb = np.stack([X[pick[j], ds * j:ds * j + ds] for j in range(M)]).astype(np.float32)
    else:
        cb = (rs.standard_normal((M, 256, ds)) * (1e-3 if kind == "scaled" else 0.8) + (0.25 if kind == "scaled" else 0.0)).astype(np.float32)
    c = ctx()
    l0 = c.launches
    codes_t = c.pq_encode(dev(X), dev(cb), metric=metric).cpu().numpy()
    c.check()
    assert c.launches - l0 >= 8  # preparation + tensor kernel + pair arbiter (the sub-vector path is ONE launch)
    monkeypatch.setenv("MEVI_PQ_TENSOR", "0")
    l0 = c.launches
    codes_s = c.pq_encode(dev(X), dev(cb), metric=metric).cpu().numpy()
    assert c.launches - l0 == 1
    monkeypatch.delenv("MEVI_PQ_TENSOR")
    assert np.array_equal(codes_t, codes_s), int((codes_t != codes_s).sum())
    sub = slice(0, 6000)
    ref = oracle.pq_encode(X[sub], cb, metric)
    ties, real = oracle.classify_pq_mismatches(X[sub], cb, codes_t[sub], ref, metric)
    assert real == 0 and ties <= 40, (metric, ties, real)


def test_pq_and_opq_document_cluster_mirror(data):
    from mevi_b200.pq import ProductQuantization

    X, _ = data
    cb, rot = g("pq_codebook.npy"), g("opq_rotate.npy")
    pq = ProductQuantization("pq", 4, 4, "l2", 64, "kmeans", "grad")
    with torch.no_grad():
        pq.codebook.copy_(torch.tensor(cb))
    clus, mapping = pq.get_document_cluster(X, 0, 1, 128, True)
    ref = g("pq_codes_l2.npy")
    got = np.array([mapping[i] for i in range(X.shape[0])], dtype=np.int32)
    ties, real = oracle.classify_pq_mismatches(X, cb, got, ref)
    assert real == 0 and ties <= 2
    assert sum(len(v) for v in clus.values()) == X.shape[0]
    # 2-rank shard of the same call (pq.py:218-225)
    _, m1 = pq.get_document_cluster(X, 1, 2, 128, True)
    assert min(m1) == X.shape[0] // 2 and all(m1[i] == mapping[i] for i in m1)
    opq = ProductQuantization("opq", 4, 4, "l2", 64, "kmeans", "grad")
    with torch.no_grad():
        opq.codebook.copy_(torch.tensor(cb))
        opq.rotate.copy_(torch.tensor(rot))
    _, mapping = opq.get_document_cluster(X, 0, 1, 128, True)
    got = np.array([mapping[i] for i in range(X.shape[0])], dtype=np.int32)
    # the rotation is a library GEMM whose rounding differs from the CPU's: arbitrate on the float64 rotation,
    # with the tie window widened to the rotation's own rounding (1e-5 relative)
    Xr = (X.astype(np.float64) @ rot.astype(np.float64).T).astype(np.float32)
    ties, real = oracle.classify_pq_mismatches(Xr, cb, got, g("opq_codes_l2.npy"), tie_eps=1e-5)
    assert real == 0 and ties <= 4


def test_pq_training_reaches_reference_quality(data):
    from mevi_b200.pq import ProductQuantization

    X, _ = data
    pq = ProductQuantization("pq", 4, 4, "l2", 64, "kmeans", "grad")
    pq.unsupervised_update_codebook_manually(X, 41, "kmeans")
    assert tuple(pq.codebook.shape) == (4, 16, 16) and pq.last_preds.shape == (X.shape[0], 4)

    def mse(cb, codes):
        rec = np.concatenate([cb[j][codes[:, j]] for j in range(4)], axis=1)
        return float(((X - rec) ** 2).mean())

    ours = mse(pq.codebook.detach().numpy(), np.asarray(pq.last_preds))
    ref = mse(g("pq_codebook.npy"), g("pq_last_preds.npy"))
    assert ours <= 1.02 * ref, (ours, ref)
    # labels are the codes of the final codebook
    again = ctx().pq_encode(dev(X), pq.codebook.detach().cuda(), metric="l2").cpu().numpy()
    ties, real = oracle.classify_pq_mismatches(X, pq.codebook.detach().numpy(), again, np.asarray(pq.last_preds), tie_eps=1e-5)
    assert real == 0


# ---- accumulate by code / EMA ---------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,K,M", [(512, 64, 16, 3), (100000, 768, 32, 4), (7, 768, 32, 1), (3000, 128, 256, 2)])
def test_accumulate_by_code_matches_float64(n, d, K, M):
    rs = np.random.RandomState(n)
    X = rs.standard_normal((n, d)).astype(np.float32)
    codes = rs.randint(0, K, size=(n, M)).astype(np.int32)
    codes[: n // 2, 0] = 1  # skew
    c = ctx()
    Xd, cd = dev(X), dev(codes)
    for j in range(M):
        buf = c.accumulate_by_code(Xd, cd[:, j], K, assign_stride=M).cpu().numpy().astype(np.float64)
        sums = np.zeros((K, d))
        np.add.at(sums, codes[:, j], X.astype(np.float64))
        counts = np.bincount(codes[:, j], minlength=K)
        assert (buf[K * d:] == counts).all()
        scale = np.abs(X).astype(np.float64).sum(0).max() + 1e-30
        assert np.abs(buf[: K * d].reshape(K, d) - sums).max() <= 2e-6 * scale
    if K <= 64:  # shared-memory accumulators: bit-reproducible
        a = c.accumulate_by_code(Xd, cd[:, 0], K, assign_stride=M)
        b = c.accumulate_by_code(Xd, cd[:, 0], K, assign_stride=M)
        assert torch.equal(a, b)


@pytest.mark.parametrize("kind", ["rq", "pq"])
def test_ema_update_matches_reference_golden(data, kind):
    from mevi_b200.pq import ProductQuantization

    X, _ = data
    if kind == "rq":
        cb = torch.load(os.path.join(MODES, "..", "small64", "codebook.pt"), map_location="cpu", weights_only=False).detach()
        bits = 4
    else:
        cb, bits = torch.tensor(g("pq_codebook.npy")), 4
    M = cb.shape[0]
    e = ProductQuantization(kind, M, bits, "l2", 64, "kmeans", "ema")
    e.restart_unused_codes = False
    with torch.no_grad():
        e.codebook.copy_(cb)
        e.embed_ema.copy_(cb)
        e.cluster_size_ema.fill_(1.0)
    e = e.cuda()
    e.train()
    vec = torch.tensor(X[:512].copy()).cuda()
    proba, index, loss = e.forward(vec)
    assert (index.cpu().numpy() == g(f"ema_{kind}_index.npy")).all()
    np.testing.assert_allclose(vec.cpu().numpy(), g(f"ema_{kind}_vecs_after.npy"), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(e.cluster_size_ema.cpu().numpy(), g(f"ema_{kind}_size.npy"), rtol=1e-6)
    np.testing.assert_allclose(e.embed_ema.cpu().numpy(), g(f"ema_{kind}_embed.npy"), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(e.codebook.detach().cpu().numpy(), g(f"ema_{kind}_codebook.npy"), rtol=1e-5, atol=1e-6)
    # with the restart on (pq.py:404-423): unused codes are re-seeded from the batch, sizes floor at 1
    e.restart_unused_codes = True
    with torch.no_grad():
        e.cluster_size_ema.fill_(0.5)
    e.forward(torch.tensor(X[:512].copy()).cuda())
    assert torch.isfinite(e.codebook).all() and (e.cluster_size_ema >= 0.99).all()


# ---- beam search on the device --------------------------------------------------------------------------
def _beam_compare(lab, sc, lab_ref, sc_ref, rtol):
    lab, sc, lab_ref, sc_ref = (np.asarray(a) for a in (lab, sc, lab_ref, sc_ref))
    assert lab.shape == lab_ref.shape and sc.shape == sc_ref.shape
    np.testing.assert_allclose(sc, sc_ref, rtol=rtol, atol=1e-30)
    bad = 0
    for q in range(lab.shape[0]):
        diff = np.nonzero((lab[q] != lab_ref[q]).any(-1))[0]
        for p in diff:
            # a differing leaf must sit in a run of (near-)equal scores: a tie torch.topk may order either way
            lo, hi = max(p - 1, 0), min(p + 1, lab.shape[1] - 1)
            near = min(abs(sc_ref[q, p] - sc_ref[q, lo]) if lo != p else np.inf, abs(sc_ref[q, p] - sc_ref[q, hi]) if hi != p else np.inf)
            assert near <= 4 * rtol * abs(sc_ref[q, p]) + 1e-30, (q, p, sc_ref[q, lo:hi + 1])
            bad += 1
    return bad


@pytest.mark.parametrize("nb", [10, 100])
def test_beam_search_matches_reference_golden(case, nb):
    from mevi_b200.pq import ProductQuantization

    pq = ProductQuantization("rq", case.M, case.meta["bits"], "l2", case.d, "kmeans", "grad")
    with torch.no_grad():
        pq.codebook.copy_(torch.tensor(case.codebook))
    lab, sc = pq.beam_search(dev(case.Q), nb, return_proba=True)
    assert lab.dtype == torch.int64 and lab.is_cuda and tuple(lab.shape) == (case.Q.shape[0], nb, case.M)
    swapped = _beam_compare(lab.cpu().numpy(), sc.cpu().numpy(), case.load(f"beam{nb}_labels.npy"),
                            case.load(f"beam{nb}_scores.npy"), rtol=2e-3)
    assert swapped <= 0.02 * lab.shape[0] * nb


@pytest.mark.parametrize("metric,prod,nb", [("l2", False, 10), ("ip", True, 20), ("ip", False, 7), ("l2", True, 1), ("l2", True, 40)])
def test_beam_search_kernel_matches_tensor_op_formulation(gauss, metric, prod, nb):
    from mevi_b200.pq import ProductQuantization

    pq = ProductQuantization("rq", gauss.M, 5, metric, gauss.d, "kmeans", "grad", rq_topk_score="prod" if prod else "sum")
    with torch.no_grad():
        pq.codebook.copy_(torch.tensor(gauss.codebook))
    # 'ip' logits of N(0,1) rows are O(|x||c|): scale the queries down so the softmax is not one-hot
    Q = gauss.Q * (0.05 if metric == "ip" else 1.0)
    lab, sc = pq.beam_search(dev(Q), nb, return_proba=True)
    lab_ref, sc_ref = pq._beam_search_tensor_ops(torch.tensor(Q), nb, True)
    _beam_compare(lab.cpu().numpy(), sc.cpu().numpy(), lab_ref.numpy(), sc_ref.numpy(), rtol=2e-3)
    if nb == 1 and metric == "l2":  # a single beam is the greedy encode
        codes = ctx().rq_encode(dev(Q), dev(gauss.codebook), mode="exact").cpu().numpy()
        assert (lab[:, 0].cpu().numpy() == codes).mean() > 0.99


def test_beam_search_beyond_the_kernel_limits_falls_back_to_tensor_ops():
    """8-bit codebooks with 100 beams (25,600 candidates per level) exceed the kernel's shared-memory state: the drop-in
    runs the reference's tensor-op formulation on the device instead of raising (pq.py:613-713 handles any shape)."""
    from mevi_b200.pq import ProductQuantization

    rs = np.random.RandomState(11)
    d = 64
    pq = ProductQuantization("rq", 2, 8, "l2", d, "kmeans", "grad")
    with torch.no_grad():
        pq.codebook.copy_(torch.tensor(rs.standard_normal((2, 256, d)).astype(np.float32)))
    assert not pq._beam_kernel_fits(100, d) and pq._beam_kernel_fits(10, d)
    Q = rs.standard_normal((5, d)).astype(np.float32) * 0.3
    lab, sc = pq.beam_search(dev(Q), 100, return_proba=True)
    lab_ref, sc_ref = pq._beam_search_tensor_ops(torch.tensor(Q), 100, True)
    assert lab.is_cuda and tuple(lab.shape) == (5, 100, 2)
    _beam_compare(lab.cpu().numpy(), sc.cpu().numpy(), lab_ref.numpy(), sc_ref.numpy(), rtol=2e-3)


def test_beam_search_rejects_more_beams_than_leaves():
    from mevi_b200 import _lib

    with pytest.raises(_lib.MeviError):
        ctx().rq_beam_search(torch.zeros((2, 64), device="cuda:0"), torch.zeros((2, 4, 64), device="cuda:0"), 17)


# ---- --eval_all_documents ----------------------------------------------------------------------------------
@pytest.mark.parametrize("pool,batch", [(100, 700), (1000, 1024), (5000, 4096)])
def test_eval_all_documents_matches_oracle(pool, batch):
    from mevi_b200.rerank import eval_all_documents

    rs = np.random.RandomState(pool)
    D = rs.standard_normal((3000, 768)).astype(np.float32)
    Q = rs.standard_normal((9, 768)).astype(np.float32)
    s_ref, i_ref = oracle.eval_all_documents(Q, D, pool, batch_size=1024)
    for src in (D, dev(D)):
        if pool > 2048:
            with pytest.raises(Exception):
                eval_all_documents(Q, src, pool, batch_size=batch)
            continue
        s, i = eval_all_documents(Q, src, pool, batch_size=batch)
        assert i.dtype == torch.int32 and tuple(s.shape) == s_ref.shape
        D64, Q64 = D.astype(np.float64), Q.astype(np.float64)
        assert_topk_equivalent(s.cpu().numpy(), i.cpu().numpy().astype(np.int64), s_ref, i_ref.astype(np.int64), rtol=1e-5,
                               atol=2e-4, pool_scores=lambda q, doc: float(D64[doc] @ Q64[q]))
