"""Mint golden vectors for the pq / opq / EMA / pq-beam branches by running the UNMODIFIED reference
(`MEVI/pq.py`).  Run once in the authoring container (needs /root/reference):

    python tests/golden/make_golden_modes.py

Inputs are the seeded `small64` documents/queries of datasets.py (2000 x 64).  Outputs (tests/golden/modes/):
  pq_codebook.npy, pq_last_preds.npy   reference k-means build, pq_type='pq', M=4, bits=4        [pq.py:551-581]
  pq_codes_l2.npy / pq_codes_ip.npy    reference get_document_cluster -> get_pq_document_cluster [217-279]
  opq_rotate.npy, opq_codes_l2.npy     same with pq_type='opq' and a seeded orthogonal rotation   [259-261]
  pq_beam8_labels/scores.npy           reference beam_search, pq branch                            [613-713]
  pq_forward_{proba,index}.npy         reference forward_pq                                        [321-337]
  ema_{rq,pq}_{codebook,embed,size}.npy, ema_{rq,pq}_index.npy
                                       reference forward() in train mode with pq_update_method='ema' and
                                       restart_unused_codes=False (the restart draws torch.randperm)  [307-319, 371-433]
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import datasets  # noqa: E402
from oracle.ref_import import load_reference_pq  # noqa: E402

M, BITS = 4, 4


def main():
    refpq = load_reference_pq()
    out = os.path.join(HERE, "modes")
    os.makedirs(out, exist_ok=True)
    X = datasets.case_docs("small64")
    Q = datasets.make_queries(64)
    d = X.shape[1]

    def save(name, a):
        np.save(os.path.join(out, name), np.asarray(a))

    # ---- pq build + encode ---------------------------------------------------
    pq = refpq.ProductQuantization("pq", M, BITS, "l2", d, "kmeans", "grad")
    pq.unsupervised_update_codebook_manually(X, 41, "kmeans")
    cb = pq.codebook.detach().numpy().copy()
    save("pq_codebook.npy", cb)
    save("pq_last_preds.npy", np.asarray(pq.last_preds).astype(np.int32))
    clus, mapping = pq.get_document_cluster(X, 0, 1, 128, True)
    codes = np.array([mapping[i] for i in range(X.shape[0])], dtype=np.int32)
    save("pq_codes_l2.npy", codes)
    pq_ip = refpq.ProductQuantization("pq", M, BITS, "ip", d, "kmeans", "grad")
    with torch.no_grad():
        pq_ip.codebook.copy_(torch.tensor(cb))
    _, mapping = pq_ip.get_document_cluster(X, 0, 1, 128, True)
    save("pq_codes_ip.npy", np.array([mapping[i] for i in range(X.shape[0])], dtype=np.int32))

    # ---- opq encode ----------------------------------------------------------
    rs = np.random.RandomState(5)
    rot, _ = np.linalg.qr(rs.standard_normal((d, d)))
    rot = rot.astype(np.float32)
    save("opq_rotate.npy", rot)
    opq = refpq.ProductQuantization("opq", M, BITS, "l2", d, "kmeans", "grad")
    with torch.no_grad():
        opq.codebook.copy_(torch.tensor(cb))
        opq.rotate.copy_(torch.tensor(rot))
    _, mapping = opq.get_document_cluster(X, 0, 1, 128, True)
    save("opq_codes_l2.npy", np.array([mapping[i] for i in range(X.shape[0])], dtype=np.int32))

    # ---- pq beam search, forward_pq -------------------------------------------
    lab, sc = pq.beam_search(torch.tensor(Q), 8, return_proba=True)
    save("pq_beam8_labels.npy", lab.numpy())
    save("pq_beam8_scores.npy", sc.numpy())
    proba, index, loss = pq.forward(torch.tensor(X[:256].copy()))
    save("pq_forward_proba.npy", proba.detach().numpy())
    save("pq_forward_index.npy", index.numpy())

    # ---- EMA update -----------------------------------------------------------
    rq_cb = torch.load(os.path.join(HERE, "small64", "codebook.pt"), map_location="cpu", weights_only=False).detach()
    for kind, cbk, m, bits in (("rq", rq_cb, rq_cb.shape[0], 4), ("pq", torch.tensor(cb), M, BITS)):
        e = refpq.ProductQuantization(kind, m, bits, "l2", d, "kmeans", "ema")
        e.restart_unused_codes = False
        with torch.no_grad():
            e.codebook.copy_(cbk)
            e.embed_ema.copy_(cbk)
            e.cluster_size_ema.fill_(1.0)
        e.train()
        vec = torch.tensor(X[:512].copy())
        proba, index, loss = e.forward(vec)
        save(f"ema_{kind}_index.npy", index.numpy())
        save(f"ema_{kind}_codebook.npy", e.codebook.detach().numpy())
        save(f"ema_{kind}_embed.npy", e.embed_ema.numpy())
        save(f"ema_{kind}_size.npy", e.cluster_size_ema.numpy())
        save(f"ema_{kind}_vecs_after.npy", vec.numpy())

    meta = {"case": "small64", "M": M, "bits": BITS, "torch": torch.__version__, "numpy": np.__version__,
            "x_sha256": datasets.sha256(X), "q_sha256": datasets.sha256(Q)}
    import sklearn
    meta["sklearn"] = sklearn.__version__
    with open(os.path.join(out, "meta.json"), "w") as fw:
        json.dump(meta, fw, indent=1)
    print("wrote", sorted(os.listdir(out)))


if __name__ == "__main__":
    main()
