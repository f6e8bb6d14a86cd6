"""Golden vectors for BASELINE.json configs[0] (synthetic 100k x 768 fp32: pq.py RQ build + encode on the reference CPU
path), minted by running the UNMODIFIED reference (`MEVI/pq.py`).  Run once in the authoring container:

    python tests/golden/make_golden_100k.py

X = the seeded `gauss100k` documents of datasets.py (N(0,1), seed 1234).  Outputs (tests/golden/gauss100k/):
  codebook.pt       torch.save of the reference's Parameter after initialize(...)               [pq.py:440-470, 551-598]
  codes_u8.npy      reference get_document_cluster codes, uint8 [100000, 4] (K = 32)             [pq.py:216-305]
  last_preds_u8.npy sklearn fit_predict labels kept by the reference (get_document_cluster_simple input)
  meta.json         sha256 of X, timings of the reference on this container's cores, versions
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import sklearn
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import datasets  # noqa: E402
from oracle.ref_import import load_reference_pq  # noqa: E402


def main():
    refpq = load_reference_pq()
    name = "gauss100k"
    kind, n, d, M, bits, seed = datasets.CASES[name]
    out = os.path.join(HERE, name)
    os.makedirs(out, exist_ok=True)
    X = datasets.case_docs(name)
    pq = refpq.ProductQuantization("rq", M, bits, "l2", d, "kmeans", "grad")
    index_file = os.path.join(out, "codebook.pt")
    if os.path.exists(index_file):
        os.remove(index_file)
    t0 = time.time()
    pq.initialize(index_file, X, 0, 41, None, 1024)
    t_build = time.time() - t0
    last_preds = np.asarray(pq.last_preds)
    t0 = time.time()
    _, mapping = pq.get_document_cluster(X, 0, 1, 128, True)
    t_enc = time.time() - t0
    codes = np.array([mapping[i] for i in range(n)], dtype=np.int32)
    assert codes.max() < 256 and last_preds.max() < 256
    np.save(os.path.join(out, "codes_u8.npy"), codes.astype(np.uint8))
    np.save(os.path.join(out, "last_preds_u8.npy"), last_preds.astype(np.uint8))
    meta = {"case": name, "kind": kind, "n": n, "d": d, "M": M, "bits": bits, "seed": seed, "kmeans_seed": 41,
            "x_sha256": datasets.sha256(X), "n_leaves": len(set(map(tuple, codes.tolist()))),
            "preds_vs_codes_mismatch_rows": int((last_preds != codes).any(1).sum()),
            "ref_build_s": round(t_build, 2), "ref_encode_s": round(t_enc, 3), "host_cpus": os.cpu_count(),
            "versions": {"torch": torch.__version__, "numpy": np.__version__, "sklearn": sklearn.__version__}}
    with open(os.path.join(out, "meta.json"), "w") as fw:
        json.dump(meta, fw, indent=1)
    print(meta)


if __name__ == "__main__":
    main()
