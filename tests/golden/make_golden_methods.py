"""Mint golden vectors for the remaining ProductQuantization methods by running the UNMODIFIED reference
(`MEVI/pq.py`).  Run once in the authoring container (needs /root/reference):

    python tests/golden/make_golden_methods.py

Inputs: the seeded `small64` documents (2000 x 64) with the reference-built codebooks already under tests/golden.
Outputs (tests/golden/methods/):
  rq_recon_loss.npy, pq_recon_loss.npy        get_reconstruct_loss_for_embeddings            [pq.py:743-766]
  rq_recon_mm.npy, pq_recon_mm.npy            get_reconstruct_vector_matrix_multiply         [pq.py:786-799]
  soft_index.npy                              its seeded [64, M, K] soft weights
  align_new.npy                               align_codebook of a seeded permutation+noise   [pq.py:600-611]
  align_in.npy                                the codebook it permuted
  avg_rq_codebook.npy, avg_pq_codebook.npy    init_pq_using_document_cluster from the golden rqclus.pkl / a pq
                                              cluster dictionary built from pq_codes_l2.npy [pq.py:488-524]
"""
from __future__ import annotations

import os
import pickle
import sys
from collections import defaultdict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import datasets  # noqa: E402
from oracle.ref_import import load_reference_pq  # noqa: E402


def main():
    refpq = load_reference_pq()
    out = os.path.join(HERE, "methods")
    os.makedirs(out, exist_ok=True)
    X = datasets.case_docs("small64")
    d = X.shape[1]

    def save(name, a):
        np.save(os.path.join(out, name), np.asarray(a))

    rq_cb = torch.load(os.path.join(HERE, "small64", "codebook.pt"), map_location="cpu", weights_only=False).detach()
    rq_codes = np.load(os.path.join(HERE, "small64", "codes.npy"))
    pq_cb = torch.tensor(np.load(os.path.join(HERE, "modes", "pq_codebook.npy")))
    pq_codes = np.load(os.path.join(HERE, "modes", "pq_codes_l2.npy"))
    M_rq, K_rq = rq_cb.shape[:2]
    M_pq, K_pq = pq_cb.shape[:2]
    bits = lambda K: int(np.log2(K))

    rq = refpq.ProductQuantization("rq", M_rq, bits(K_rq), "l2", d, "kmeans", "grad")
    pq = refpq.ProductQuantization("pq", M_pq, bits(K_pq), "l2", d, "kmeans", "grad")
    with torch.no_grad():
        rq.codebook.copy_(rq_cb)
        pq.codebook.copy_(pq_cb)
    emb = torch.tensor(X[:256])
    save("rq_recon_loss.npy", rq.get_reconstruct_loss_for_embeddings(emb, torch.tensor(rq_codes[:256]).long()).item())
    save("pq_recon_loss.npy", pq.get_reconstruct_loss_for_embeddings(emb, torch.tensor(pq_codes[:256]).long()).item())
    rs = np.random.RandomState(5)
    soft = {}
    for name, obj, Mm, Kk in (("rq", rq, M_rq, K_rq), ("pq", pq, M_pq, K_pq)):
        w = rs.random_sample((64, Mm, Kk)).astype(np.float32)
        soft[name] = w
        save(f"{name}_soft_index.npy", w)
        save(f"{name}_recon_mm.npy", obj.get_reconstruct_vector_matrix_multiply(torch.tensor(w)).detach().numpy())
    # align: a noisy, permuted copy of the rq codebook must be permuted back onto the original order
    perm = np.stack([rs.permutation(K_rq) for _ in range(M_rq)])
    noisy = np.stack([rq_cb.numpy()[j][perm[j]] for j in range(M_rq)]) + 0.01 * rs.standard_normal(rq_cb.shape).astype(np.float32)
    save("align_in.npy", noisy.astype(np.float32))
    al = refpq.ProductQuantization("rq", M_rq, bits(K_rq), "l2", d, "kmeans", "grad")
    with torch.no_grad():
        al.codebook.copy_(torch.tensor(noisy.astype(np.float32)))
        al.align_codebook(rq_cb)
    save("align_new.npy", al.codebook.detach().numpy())
    # avg init from cluster dictionaries
    av = refpq.ProductQuantization("rq", M_rq, bits(K_rq), "l2", d, "avg", "grad")
    with torch.no_grad():
        av.codebook.zero_()
        av.init_pq_using_document_cluster(X.copy(), os.path.join(HERE, "small64", "rqclus.pkl"), 128)
    save("avg_rq_codebook.npy", av.codebook.detach().numpy())
    clus = defaultdict(list)
    for i, t in enumerate(map(tuple, pq_codes.tolist())):
        clus[t].append(i)
    with open(os.path.join(out, "pqclus.pkl"), "wb") as fw:
        pickle.dump(dict(clus), fw)
    avp = refpq.ProductQuantization("pq", M_pq, bits(K_pq), "l2", d, "avg", "grad")
    with torch.no_grad():
        avp.codebook.zero_()
        avp.init_pq_using_document_cluster(X.copy(), os.path.join(out, "pqclus.pkl"), 128)
    save("avg_pq_codebook.npy", avp.codebook.detach().numpy())
    print("wrote", sorted(os.listdir(out)))


if __name__ == "__main__":
    main()
