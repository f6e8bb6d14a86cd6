"""pq / opq / EMA / eval-all-documents branches: oracle restatements pinned to reference golden vectors
(tests/golden/modes, minted by make_golden_modes.py from the unmodified MEVI/pq.py) and the host-side mirrors."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import oracle

MODES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "modes")


def g(name):
    return np.load(os.path.join(MODES, name))


@pytest.fixture(scope="module")
def data():
    import datasets

    meta = json.load(open(os.path.join(MODES, "meta.json")))
    X = datasets.case_docs("small64")
    Q = datasets.make_queries(64)
    assert datasets.sha256(X) == meta["x_sha256"] and datasets.sha256(Q) == meta["q_sha256"]
    return X, Q, meta


def test_oracle_pq_encode_matches_reference(data):
    X, _, _ = data
    cb = g("pq_codebook.npy")
    assert (oracle.pq_encode(X, cb, "l2", batch_size=128) == g("pq_codes_l2.npy")).all()
    assert (oracle.pq_encode(X, cb, "ip", batch_size=128) == g("pq_codes_ip.npy")).all()
    assert (oracle.pq_encode(X, cb, "l2", rotate=g("opq_rotate.npy"), batch_size=128) == g("opq_codes_l2.npy")).all()
    # the reference's own k-means labels agree with a re-encode except at fp32 near-ties
    lp = g("pq_last_preds.npy")
    ties, real = oracle.classify_pq_mismatches(X, cb, lp, g("pq_codes_l2.npy"), tie_eps=1e-5)
    assert real == 0


@pytest.mark.parametrize("kind", ["rq", "pq"])
def test_oracle_ema_matches_reference(data, kind):
    X, _, meta = data
    if kind == "rq":
        cb = torch.load(os.path.join(MODES, "..", "small64", "codebook.pt"), map_location="cpu", weights_only=False).detach()
    else:
        cb = torch.tensor(g("pq_codebook.npy"))
    M, K = cb.shape[0], cb.shape[1]
    # the reference hands ema_update the tensor forward_rq modified in place (pq.py:357, 317-318)
    vec_seen = g(f"ema_{kind}_vecs_after.npy")
    if kind == "pq":
        assert np.array_equal(vec_seen, X[:512])
    else:
        assert not np.array_equal(vec_seen, X[:512])
    sums, counts = oracle.ema_sums_counts(vec_seen, g(f"ema_{kind}_index.npy"), M, K, kind)
    new_cb, ee, cs = oracle.ema_apply(cb, cb, torch.ones(M, K), sums, counts)
    np.testing.assert_allclose(cs.numpy(), g(f"ema_{kind}_size.npy"), rtol=1e-6)
    np.testing.assert_allclose(ee.numpy(), g(f"ema_{kind}_embed.npy"), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(new_cb.numpy(), g(f"ema_{kind}_codebook.npy"), rtol=1e-5, atol=1e-6)


def test_mirror_constructor_modes():
    from mevi_b200.pq import ProductQuantization

    pq = ProductQuantization("pq", 4, 4, "l2", 64, "kmeans", "grad")
    assert tuple(pq.codebook.shape) == (4, 16, 16) and pq.last_dim == 16
    opq = ProductQuantization("opq", 4, 4, "l2", 64, "kmeans", "grad")
    assert tuple(opq.rotate.shape) == (64, 64) and not opq.rotate.requires_grad
    ema = ProductQuantization("rq", 3, 4, "l2", 64, "kmeans", "ema")
    assert tuple(ema.cluster_size_ema.shape) == (3, 16) and tuple(ema.embed_ema.shape) == (3, 16, 64)
    assert not ema.codebook.requires_grad and ema.decay == 0.99 and ema.eps == 1e-5 and ema.restart_unused_codes
    assert {"codebook", "cluster_size_ema", "embed_ema"} <= set(ema.state_dict())


def test_mirror_pq_forward_and_beam_match_reference(data):
    from mevi_b200.pq import ProductQuantization

    X, Q, _ = data
    pq = ProductQuantization("pq", 4, 4, "l2", 64, "kmeans", "grad")
    with torch.no_grad():
        pq.codebook.copy_(torch.tensor(g("pq_codebook.npy")))
    proba, index, loss = pq.forward(torch.tensor(X[:256].copy()))
    assert loss is None and (index.numpy() == g("pq_forward_index.npy")).all()
    np.testing.assert_allclose(proba.detach().numpy(), g("pq_forward_proba.npy"), rtol=1e-6)
    lab, sc = pq.beam_search(torch.tensor(Q), 8, return_proba=True)
    assert (lab.numpy() == g("pq_beam8_labels.npy")).all()
    np.testing.assert_allclose(sc.numpy(), g("pq_beam8_scores.npy"), rtol=1e-6)
    rec = pq.get_reconstruct_vector(index)
    ref = torch.cat([pq.codebook[j][index[:, j]] for j in range(4)], dim=-1)
    assert torch.equal(rec, ref)


def test_oracle_eval_all_documents_is_streaming_flat_topk(data):
    X, Q, _ = data
    s, i = oracle.eval_all_documents(Q, X, 50, batch_size=300)
    s2, i2 = oracle.flat_ip_topk(Q, X, 50)
    assert i.dtype == np.int32 and s.shape == (Q.shape[0], 50)
    np.testing.assert_allclose(s, s2, rtol=1e-5, atol=1e-5)
    assert (np.sort(i, 1) == np.sort(i2, 1)).mean() > 0.99


def test_pq_encode_equals_rq_encode_on_a_block_padded_codebook():
    """Design check for moving the PQ encode onto the RQ tensor kernel: zero-padding sub-vector
    centroid (j,k) to the full width makes the levels' supports orthogonal, so the residual corrections vanish
    (r_j . c = x . c) and the greedy RQ encode of pq.py:281-305 returns the PQ codes of pq.py:249-279 — up to fp32
    ties, because the padded direct-form distance adds the same constant to every centroid of a level."""
    from oracle import oracle

    rs = np.random.RandomState(3)
    n, d, M, K = 4000, 96, 4, 32
    X = rs.standard_normal((n, d)).astype(np.float32)
    cb = rs.standard_normal((M, K, d // M)).astype(np.float32)
    padded = np.zeros((M, K, d), np.float32)
    for j in range(M):
        padded[j, :, j * (d // M):(j + 1) * (d // M)] = cb[j]
    pq_codes = oracle.pq_encode(X, cb, dist_mode="l2")
    rq_codes = oracle.rq_encode(X, padded, batch_size=512)
    differ = np.nonzero((pq_codes != rq_codes).any(axis=1))[0]
    assert len(differ) <= 2e-3 * n
    ties, real = oracle.classify_pq_mismatches(X, cb, pq_codes, rq_codes)
    assert real == 0, (ties, real)
