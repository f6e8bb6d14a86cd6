"""Mint golden vectors by running the UNMODIFIED reference (`MEVI/pq.py`).

Run once in the authoring container (needs /root/reference):
    python tests/golden/make_golden.py [case ...]

For every case in `datasets.CASES` it executes, through the reference's own
entry points (pq.py line numbers in brackets):
  * ProductQuantization(...).initialize(index_file, X, 0, 41, None, 1024)   [441]
      -> sklearn MiniBatchKMeans build [551-598]; the reference itself
         torch.save()s the codebook Parameter to `codebook.pt`           [469-470]
      -> last_preds (fit_predict codes)                                   [595-596]
  * get_document_cluster_simple(True)                                     [201]
  * get_document_cluster(X, 0, 1, 128, True) and the 2-rank sharded form  [217]
  * beam_search(Q, 10 / 100, return_proba=True)                           [614]
  * dist_mode='ip' encode with the same codebook                          [124-127]
and stores the outputs as small fixtures next to this script.  X itself is not
stored: it regenerates from the seed (sha256 pinned in meta.json).
"""
from __future__ import annotations

import json
import os
import pickle
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import datasets  # noqa: E402
from oracle.ref_import import load_reference_pq  # noqa: E402


def run_case(name: str) -> None:
    import sklearn

    refpq = load_reference_pq()
    kind, n, d, M, bits, seed = datasets.CASES[name]
    out = os.path.join(HERE, name)
    os.makedirs(out, exist_ok=True)
    X = datasets.case_docs(name)
    Q = datasets.make_queries(d)

    pq = refpq.ProductQuantization("rq", M, bits, "l2", d, "kmeans", "grad")
    index_file = os.path.join(out, "codebook.pt")
    if os.path.exists(index_file):
        os.remove(index_file)
    t0 = time.time()
    pq.initialize(index_file, X, 0, 41, None, 1024)
    t_build = time.time() - t0
    assert pq.get_preds
    last_preds = np.asarray(pq.last_preds).astype(np.int32)
    np.save(os.path.join(out, "last_preds.npy"), last_preds)
    clus_simple, map_simple = pq.get_document_cluster_simple(True)
    with open(os.path.join(out, "rqclus_simple.pkl"), "wb") as fw:
        pickle.dump(clus_simple, fw)
    with open(os.path.join(out, "rqmapping_simple.pkl"), "wb") as fw:
        pickle.dump(map_simple, fw)

    t0 = time.time()
    clus, mapping = pq.get_document_cluster(X, 0, 1, 128, True)
    t_enc = time.time() - t0
    codes = np.array([mapping[i] for i in range(n)], dtype=np.int32)
    np.save(os.path.join(out, "codes.npy"), codes)
    with open(os.path.join(out, "rqclus.pkl"), "wb") as fw:
        pickle.dump(clus, fw)
    with open(os.path.join(out, "rqmapping.pkl"), "wb") as fw:
        pickle.dump(mapping, fw)

    # residual after all levels (the reference keeps it in a local; recompute through its own function)
    cluster_t = torch.empty((n, M), dtype=torch.int32)
    pq.get_rq_document_cluster(X, cluster_t, 0, n, 0, 1024)
    assert (cluster_t.numpy() == codes).all(), "batch size changed the reference's codes"

    # two-rank sharded encode, merged the way LogPklFile('cluster'/'dict') merges (main_models.py:289-310)
    merged_c, merged_m = {}, {}
    for r in range(2):
        c, m = pq.get_document_cluster(X, r, 2, 128, True)
        for k, v in c.items():
            merged_c.setdefault(k, []).extend(v)
        merged_m.update(m)
    assert merged_c == clus and merged_m == mapping

    # ip-metric encode with the same codebook
    pq_ip = refpq.ProductQuantization("rq", M, bits, "ip", d, "kmeans", "grad")
    with torch.no_grad():
        pq_ip.codebook.copy_(pq.codebook)
    _, map_ip = pq_ip.get_document_cluster(X, 0, 1, 128, True)
    codes_ip = np.array([map_ip[i] for i in range(n)], dtype=np.int32)
    np.save(os.path.join(out, "codes_ip.npy"), codes_ip)

    beams = {}
    for nb in (10, 100):
        if nb > 2 ** bits * 2 ** bits:
            continue
        lab, sc = pq.beam_search(torch.tensor(Q), nb, return_proba=True)
        beams[nb] = (lab.numpy(), sc.numpy())
        np.save(os.path.join(out, f"beam{nb}_labels.npy"), lab.numpy())
        np.save(os.path.join(out, f"beam{nb}_scores.npy"), sc.numpy())

    meta = {
        "case": name, "kind": kind, "n": n, "d": d, "M": M, "bits": bits, "seed": seed,
        "kmeans_seed": 41, "x_sha256": datasets.sha256(X), "q_sha256": datasets.sha256(Q),
        "n_leaves": len(clus), "preds_vs_codes_mismatch_rows": int((last_preds != codes).any(1).sum()),
        "ref_build_s": round(t_build, 2), "ref_encode_s": round(t_enc, 3),
        "versions": {"torch": torch.__version__, "numpy": np.__version__, "sklearn": sklearn.__version__},
        "beam_label_dtype": str(beams[10][0].dtype) if 10 in beams else None,
    }
    with open(os.path.join(out, "meta.json"), "w") as fw:
        json.dump(meta, fw, indent=1)
    print(json.dumps(meta))


if __name__ == "__main__":
    names = sys.argv[1:] or list(datasets.CASES)
    for nm in names:
        run_case(nm)
