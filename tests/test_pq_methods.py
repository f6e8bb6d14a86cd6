"""The remaining ProductQuantization methods reference callers use (reconstruction helpers, align, 'avg' init),
against golden vectors minted from the unmodified MEVI/pq.py (tests/golden/make_golden_methods.py)."""
import os
import pickle

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
METH = os.path.join(GOLD, "methods")


def g(name):
    return np.load(os.path.join(METH, name))


def _pqs():
    from mevi_b200.pq import ProductQuantization

    rq_cb = torch.load(os.path.join(GOLD, "small64", "codebook.pt"), map_location="cpu", weights_only=False).detach()
    pq_cb = torch.tensor(np.load(os.path.join(GOLD, "modes", "pq_codebook.npy")))
    d = 64
    rq = ProductQuantization("rq", rq_cb.shape[0], int(np.log2(rq_cb.shape[1])), "l2", d, "kmeans", "grad")
    pq = ProductQuantization("pq", pq_cb.shape[0], int(np.log2(pq_cb.shape[1])), "l2", d, "kmeans", "grad")
    with torch.no_grad():
        rq.codebook.copy_(rq_cb)
        pq.codebook.copy_(pq_cb)
    return rq, pq


def test_reconstruct_helpers_match_reference():
    import datasets

    X = datasets.case_docs("small64")
    rq, pq = _pqs()
    emb = torch.tensor(X[:256])
    rq_codes = torch.tensor(np.load(os.path.join(GOLD, "small64", "codes.npy"))[:256]).long()
    pq_codes = torch.tensor(np.load(os.path.join(GOLD, "modes", "pq_codes_l2.npy"))[:256]).long()
    assert rq.get_reconstruct_loss_for_embeddings(emb, rq_codes).item() == pytest.approx(float(g("rq_recon_loss.npy")), rel=1e-6)
    assert pq.get_reconstruct_loss_for_embeddings(emb, pq_codes).item() == pytest.approx(float(g("pq_recon_loss.npy")), rel=1e-6)
    for name, obj in (("rq", rq), ("pq", pq)):
        out = obj.get_reconstruct_vector_matrix_multiply(torch.tensor(g(f"{name}_soft_index.npy")))
        np.testing.assert_allclose(out.detach().numpy(), g(f"{name}_recon_mm.npy"), rtol=1e-6, atol=1e-6)


def test_align_codebook_matches_reference():
    from mevi_b200.pq import ProductQuantization

    rq, _ = _pqs()
    al = ProductQuantization("rq", rq.subvector_num, rq.subvector_bits, "l2", 64, "kmeans", "grad")
    with torch.no_grad():
        al.codebook.copy_(torch.tensor(g("align_in.npy")))
        al.align_codebook(rq.codebook.detach())
    assert np.array_equal(al.codebook.detach().numpy(), g("align_new.npy"))


def test_faiss_branches_raise_not_implemented():
    rq, _ = _pqs()
    for call in (lambda: rq.codebook_from_index(None, "x.index"), lambda: rq.build_faiss_index(np.zeros((4, 64), np.float32)),
                 lambda: rq.unsupervised_update_codebook_faiss(None, 0)):
        with pytest.raises(NotImplementedError):
            call()
    assert rq.wrapped_augment_xb("unchanged") == "unchanged"  # l2: identity, as pq.py:98-99
    xb = np.random.RandomState(0).standard_normal((5, 64)).astype(np.float32)
    aug = rq.augment_xb(xb)
    assert aug.shape == (5, 65) and np.allclose((aug ** 2).sum(1), (xb ** 2).sum(1).max())
    assert rq.augment_xq(torch.zeros(3, 64)).shape == (3, 65) and rq.augment_xq(xb).shape == (5, 65)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["rq", "pq"])
def test_avg_init_from_document_cluster_matches_reference(kind, tmp_path):
    """pq.py:488-524 through the device accumulate-by-code kernel; the reference accumulates in float64."""
    import datasets
    from mevi_b200.pq import ProductQuantization

    X = datasets.case_docs("small64")
    rq, pq = _pqs()
    src = rq if kind == "rq" else pq
    path = os.path.join(GOLD, "small64", "rqclus.pkl") if kind == "rq" else os.path.join(METH, "pqclus.pkl")
    av = ProductQuantization(kind, src.subvector_num, src.subvector_bits, "l2", 64, "avg", "grad")
    av.device_index = 0
    with torch.no_grad():
        av.codebook.zero_()
    av.init_pq_using_document_cluster(X, path, 128)
    np.testing.assert_allclose(av.codebook.detach().numpy(), g(f"avg_{kind}_codebook.npy"), rtol=2e-5, atol=2e-6)
