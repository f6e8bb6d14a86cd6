"""Pin the re-rank oracle to the reference's OWN code: execute the source lines of MEVI/main_models.py's re-rank loop.

`main_models.py` cannot be imported here (pytorch_lightning, faiss and a vendored transformers fork are absent), but
the cluster-restricted re-rank is a self-contained block of `T5FineTuner.infer` (the `else:` branch that starts at
`ndoc = []` and ends at `assert q_ind == len(query_embedding)`, main_models.py:3912-4055).  This script reads those
lines from /root/reference, dedents them and exec()s them UNMODIFIED in a namespace that supplies the names the block
reads (`args`, `self`, `dec`, `doc_cluster`, `query_embedding`, ...).  The helpers it calls are exec()ed from the
reference source as well: `get_inference_scores` (main_models.py), `compute_similarity` / `generate`
(document_encoder.py) and `LogTxtFile.flush`'s line format.  `Tensor.cuda()` is the identity here (no GPU).

Outputs (tests/golden/rerank/<case>_<variant>.pkl): per query the reference's sorted document list, the scores it
writes, its candidate count, and the hard-negative lines.  tests/test_oracle_golden.py checks oracle.cluster_rerank,
oracle.hn_result_line and the doc_multiclus / knn_topk_by_step restatements against them.

Run in the build container only (needs /root/reference):  python tests/golden/make_rerank_golden.py
"""
import ast
import io
import os
import pickle
import sys
import textwrap
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import datasets  # noqa: E402

REF = "/root/reference/MEVI"
BEAMS = {"gauss768": 100, "small64": 10}
OUT = os.path.join(HERE, "rerank")


def _method_source(path, cls, name):
    src = open(path).read()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.ClassDef) and (cls is None or node.name == cls):
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name == name:
                    seg = ast.get_source_segment(src, f)
                    return textwrap.dedent(" " * f.col_offset + seg)
    raise KeyError((cls, name))


def _rerank_block():
    """The reference lines, from `ndoc = []` of the cluster branch to `assert q_ind == len(query_embedding)`."""
    lines = open(os.path.join(REF, "main_models.py")).read().split("\n")
    end = next(i for i, l in enumerate(lines) if l.strip() == "assert q_ind == len(query_embedding)")
    start = max(i for i, l in enumerate(lines[:end]) if l.strip() == "ndoc = []")
    assert lines[start - 1].strip() == "else:" and 3900 < start < 3930 and 4040 < end < 4070, (start, end)
    return textwrap.dedent("\n".join(lines[start:end + 1])), (start + 1, end + 1)


class _Embeddings:
    """IndexedData.__getitem__ (main_models.py:1011-1017) with one array and torch_dtype=float32."""

    def __init__(self, X):
        self.X = X

    def __getitem__(self, keys):
        return torch.tensor(self.X[keys], dtype=torch.float32)


class _Log:
    def __init__(self):
        self.lines = []

    def add(self, item):
        self.lines.append(item)

    def text(self):  # LogTxtFile.flush (main_models.py:254-257)
        buf = io.StringIO()
        for line in self.lines:
            print(*line, file=buf, sep="\t")
        return buf.getvalue()


def run_reference(X, Q, doc_cluster, dec, beam_scores, *, doc_multiclus=1, aggr="add", knn_topk_by_step=0, pool_size=100,
                  save_hard_neg=0, batch_size=1024, gt_docs=None, pq_mapping=None):
    block, _ = _rerank_block()
    ns = {}
    exec(_method_source(os.path.join(REF, "main_models.py"), "T5FineTunerWithValidation", "get_inference_scores"),
         {"torch": torch}, ns)
    enc_ns = {}
    glob = {"torch": torch, "Dict": dict, "Tensor": torch.Tensor, "DocEncOutput": SimpleNamespace}
    for name in ("compute_similarity", "generate"):
        exec(_method_source(os.path.join(REF, "document_encoder.py"), "DocumentEncoder", name), glob, enc_ns)
    Encoder = type("Encoder", (), enc_ns)
    args = SimpleNamespace(knn_topk_by_step=knn_topk_by_step, codebook=1, doc_multiclus=doc_multiclus,
                           multiclus_score_aggr=aggr, infer_reconstruct_vector=0, save_hard_neg=save_hard_neg,
                           dataset="marco" if gt_docs is not None else "nq_dpr", eval_all_documents=0, use_topic_model=0)
    me = SimpleNamespace(additional_reconstruct=False, all_embeddings=_Embeddings(X), pemb_projection=None,
                         unified_projection=None, document_encoder=Encoder(), args=args, hn_log_texts=_Log(),
                         pq_mapping=pq_mapping)
    me.get_inference_scores = lambda *a, **k: ns["get_inference_scores"](me, *a, **k)
    nq, L = dec.shape[:2]
    # the loop scores leaf i of query didx with query_embedding[q_ind], q_ind running over all (query, leaf) pairs: the
    # caller repeats every query embedding once per beam (main_models.py: query_embedding is [bs * num_beams, d])
    env = dict(torch=torch, np=np, args=args, self=me, dec=dec, doc_cluster=doc_cluster,
               query_embedding=torch.tensor(np.repeat(Q, L, axis=0)), nci_scores=torch.tensor(beam_scores),
               batch={"doc_ids": gt_docs}, texts=[f"query {i}" for i in range(nq)], batch_size=batch_size,
               pool_size=pool_size, eos_idx=None, torch_dec=None, pemb_dec=None)
    cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda t, *a, **k: t
    try:
        exec(block, env)
    finally:
        torch.Tensor.cuda = cuda
    return env["result_docs"], env["ndoc"], me.hn_log_texts


def multi_cluster_dict(codes, rs, extra=0.3):
    """doc_cluster of a --doc_multiclus > 1 index: a fraction of the documents is also listed under a second leaf."""
    from collections import defaultdict

    clus = defaultdict(list)
    for i, c in enumerate(codes):
        clus[tuple(int(v) for v in c)].append(i)
    keys = list(clus.keys())
    for i in np.nonzero(rs.rand(len(codes)) < extra)[0]:
        clus[keys[rs.randint(len(keys))]].append(int(i))
    return dict(clus)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    block, span = _rerank_block()
    print(f"reference block: main_models.py:{span[0]}-{span[1]} ({len(block.splitlines())} lines)")
    for case in ("gauss768", "small64"):
        kind, n, d, M, bits, seed = datasets.CASES[case]
        X = datasets.case_docs(case)
        Q = datasets.make_queries(d)
        cdir = os.path.join(HERE, case)
        codes = np.load(os.path.join(cdir, "codes.npy"))
        clus = pickle.load(open(os.path.join(cdir, "rqclus.pkl"), "rb"))
        mapping = pickle.load(open(os.path.join(cdir, "rqmapping.pkl"), "rb"))
        nb = BEAMS[case]  # 3,000 documents over ~2,700 leaves: 100 beams give a gauss768 query ~100 candidates
        dec = np.load(os.path.join(cdir, f"beam{nb}_labels.npy"))
        bsc = np.load(os.path.join(cdir, f"beam{nb}_scores.npy"))
        rs = np.random.RandomState(11)
        gt = [[int(rs.randint(n))] for _ in range(len(Q))]
        variants = {
            # the shipped recipe: every candidate, sorted, hard-negative lines with the ground-truth score (marco) ...
            "shipped": dict(save_hard_neg=n, gt_docs=gt, pq_mapping=mapping),
            # ... and without it (nq_dpr), truncated lists
            "hn100": dict(save_hard_neg=100),
            "topk_by_step": dict(knn_topk_by_step=1, pool_size=50, save_hard_neg=50),
        }
        for name, kw in variants.items():
            docs, ndoc, log = run_reference(X, Q, clus, dec, bsc, **kw)
            pickle.dump({"docs": docs, "ndoc": ndoc, "lines": log.text(), "span": span},
                        open(os.path.join(OUT, f"{case}_{name}.pkl"), "wb"))
            print(case, name, "queries", len(docs), "first list", docs[0][:5], "ndoc", ndoc[:4])
        mc = multi_cluster_dict(codes, np.random.RandomState(5))
        pickle.dump(mc, open(os.path.join(OUT, f"{case}_multiclus_dict.pkl"), "wb"))
        for aggr in ("add", "max"):
            docs, ndoc, log = run_reference(X, Q, mc, dec, bsc, doc_multiclus=2, aggr=aggr, save_hard_neg=200)
            pickle.dump({"docs": docs, "ndoc": ndoc, "lines": log.text(), "span": span},
                        open(os.path.join(OUT, f"{case}_multiclus_{aggr}.pkl"), "wb"))
            print(case, "multiclus", aggr, "first list", docs[0][:5], "ndoc", ndoc[:4])
