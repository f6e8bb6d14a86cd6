"""Result fusion of MEVI/ensemble_marco.py and MEVI/ensemble_nqdpr.py.

Consumes the files the index hot path writes (faiss_search.to_file / hn result lines / coarse lines /
rqmapping pickle) and reproduces the reference's fusion and report text exactly:
    score' = score + alpha / (beta * crank + 1),   times (1 - gamma * alpha) when the document's RQ leaf
    is not among the query's beam-search leaves (crank == number of leaves)
(ensemble_marco.py:221-240, ensemble_nqdpr.py:232-251), recall / MRR / hit-rate bookkeeping
(ensemble_marco.py:8-72, ensemble_nqdpr.py:9-60), text-parse caches next to the inputs (130-140) and the
`_cr4gt.pkl` / `_cr.pkl` rank caches (176-209).  Pure host-side dictionary work — there is no tensor
math here; it is provided so the reference's last pipeline stage runs against this package's outputs.
"""
from __future__ import annotations

import ast
import os.path as osp
import pickle
from itertools import chain
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np


# ---- files ---------------------------------------------------------------------------------------
def _as_list(item: str):
    """The reference evals 'a,b,c' or '[[..],[..]]' (ensemble_marco.py:85-89); literal_eval is the safe equivalent."""
    if not item:
        return []
    if item[0] != "[":
        item = f"[{item}]"
    return list(ast.literal_eval(item))


def parse_file(fpath: str, template: Dict[str, int], key_file: Optional[str] = None):
    """ensemble_marco.py:92-111 (keys = query text) / ensemble_nqdpr.py:81-113 (keys = line index of the
    query in `key_file`, or the line number when key_file is None)."""
    qind, pind, sind, cind = template["query"], template.get("pred"), template.get("score"), template.get("cluster")
    preds, scores, clusters = {}, {}, {}
    index_of = None
    if key_file is not None:
        with open(key_file, "r") as fr:
            index_of = {line.rstrip("\n").split("\t")[0]: i for i, line in enumerate(fr)}
    with open(fpath, "r") as fr:
        for i, line in enumerate(fr):
            items = line.rstrip("\n").split("\t")
            key = items[qind]
            if key_file is not None:
                key = index_of[key]
            elif template.get("_by_line"):
                key = i
            if pind is not None:
                preds[key] = _as_list(items[pind])
            if sind is not None:
                scores[key] = _as_list(items[sind])
            if cind is not None:
                clusters[key] = _as_list(items[cind])
    return preds, scores, clusters


def _strip_ext(fpath: str) -> str:
    return fpath[: -(len(fpath.split(".")[-1]) + 1)]


def check_cache(fpath: str, template: Dict[str, int], key_file: Optional[str] = None):
    """ensemble_marco.py:130-140: parsed triples are cached as <file>.pkl next to the text file."""
    cache = _strip_ext(fpath) + ".pkl"
    if fpath.endswith(".pkl"):
        with open(fpath, "rb") as fr:
            return pickle.load(fr)
    if osp.exists(cache):
        with open(cache, "rb") as fr:
            return pickle.load(fr)
    res = parse_file(fpath, template, key_file)
    with open(cache, "wb") as fw:
        pickle.dump(res, fw)
    return res


def _resolve(fpath, dpath, nonexist_ok=False):
    if fpath is not None and not osp.exists(fpath):
        fpath = osp.join(dpath, fpath)
    if nonexist_ok:
        return fpath, fpath is not None and osp.exists(fpath)
    assert osp.exists(fpath), fpath
    return fpath


# ---- arithmetic ------------------------------------------------------------------------------------
def cluster_rankings(preds_by_query, coarse_clusters, mapping):
    """Rank of each predicted document's RQ leaf in the query's ordered leaf list, len(leaves) when the
    leaf is not there or the id is the -1 padding (ensemble_marco.py:181-191).  Returns (ranks, num_leaves)."""
    out, num = {}, None
    for q, docs in preds_by_query.items():
        cr = {}
        for i, clus in enumerate(coarse_clusters[q]):
            cr[tuple(clus)] = i
        assert num in (None, len(cr)), "queries must all carry the same number of leaves"
        num = len(cr)
        out[q] = [cr.get(mapping[p] if p != -1 else -1, len(cr)) for p in docs]
    return out, num


def fuse(preds: Sequence[int], scores: Sequence[float], cranks: Iterable[int], alpha: float, beta: float, gamma: float,
         num_clusters: int) -> Dict[int, float]:
    """One query: ensemble_marco.py:233-237.  Later occurrences of a document overwrite earlier ones."""
    fused = {}
    for p, s, crank in zip(preds, scores, cranks):
        v = s + alpha / (beta * crank + 1)
        if crank == num_clusters:
            v *= (1 - gamma * alpha)
        fused[p] = v
    return fused


def ranking_of(fused: Dict[int, float]) -> List[int]:
    """Documents by descending fused score; ties keep insertion order (Python's stable sort, as the reference)."""
    return [p for p, _ in sorted(fused.items(), key=lambda x: -x[1])]


# ---- reports -----------------------------------------------------------------------------------------
def _report(scoring, lines, ofile):
    print(f"{scoring}")
    for ln in lines:
        print(*ln)
    print()
    if ofile is not None:
        with open(ofile, "a") as fw:
            print(f"Scoring {scoring}", file=fw)
            for ln in lines:
                print(*ln, file=fw)
            print(file=fw)


def evaluate_marco(scoring, recall_num, ofile, gts, scores=None, ranks=None):
    """ensemble_marco.py:8-72: recall@n averaged over ground truths, MRR@n of the best-ranked ground truth."""
    recalls = {r: 0 for r in recall_num}
    mrrs = {r: 0 for r in recall_num}
    for q in gts:
        preds = ranks[q] if ranks is not None else ranking_of(scores[q])
        vs = [preds.index(g) if g in preds else None for g in gts[q]]
        valid = [v for v in vs if v is not None]
        best = min(valid) if valid else None
        for n in recall_num:
            if valid:
                recalls[n] += sum(v < n for v in valid) / len(vs)
                mrrs[n] += 1 / (best + 1) if best < n else 0
    nq = len(gts)
    lines = [(f"Recall{k}", v / nq) for k, v in recalls.items()] + [(f"MRR{k}", v / nq) for k, v in mrrs.items()]
    _report(scoring, lines, ofile)
    return {k: v / nq for k, v in recalls.items()}, {k: v / nq for k, v in mrrs.items()}


def evaluate_nqdpr(scoring, recall_num, ofile, nq_eval, scores=None, ranks=None):
    """ensemble_nqdpr.py:9-60: a document is a hit for query qind if qind is in its inverse-answer list."""
    offsets, array = nq_eval
    src = scores if scores is not None else ranks
    mrrs = {r: 0 for r in recall_num}
    hits = {r: 0 for r in recall_num}
    for qind in src.keys():
        preds = ranks[qind] if ranks is not None else ranking_of(scores[qind])
        ind = None
        for j, res in enumerate(preds):
            if qind in array[offsets[res]:offsets[res + 1]]:
                ind = j
                break
        for n in recall_num:
            if ind is not None:
                mrrs[n] += 1 / (ind + 1) if ind < n else 0
                hits[n] += ind < n
    nq = len(src)
    lines = [(f"MRR{k}", v / nq) for k, v in mrrs.items()] + [(f"HitRate{k}", v / nq) for k, v in hits.items()]
    _report(scoring, lines, ofile)
    return {k: v / nq for k, v in mrrs.items()}, {k: v / nq for k, v in hits.items()}


# ---- drivers -----------------------------------------------------------------------------------------
def _floats(xs, dtype=float):
    return [dtype(x) for x in xs.split(",")]


def _rank_cache(path, preds, coarse, mapping, num_clusters):
    if osp.exists(path):
        with open(path, "rb") as fr:
            ranks, num = pickle.load(fr)
    else:
        ranks, num = cluster_rankings(preds, coarse, mapping)
        with open(path, "wb") as fw:
            pickle.dump((ranks, num), fw)
    assert num_clusters in (None, num)
    return ranks, num


def _sweep(args, keys, ance_preds, ance_scores, ranks_gt, fexists, fine_preds, fine_scores, ranks_fine, num_clusters, evaluate, truth):
    for alpha in args.alphas:
        for beta in args.betas:
            for gamma in args.gammas:
                scores = {q: {} for q in keys}
                for q, apreds in ance_preds.items():
                    ascores, cr = ance_scores[q], ranks_gt[q]
                    if fexists:
                        apreds = apreds + fine_preds[q]
                        ascores = ascores + fine_scores[q]
                        cr = chain(cr, ranks_fine[q])
                    scores[q] = fuse(apreds, ascores, cr, alpha, beta, gamma, num_clusters)
                evaluate(f"score + {alpha} / ({beta} * crank + 1); punishment (1 - {gamma} * {alpha})", args.recall_num,
                         args.ofile, truth, scores=scores)


def combine_main_marco(args):
    """ensemble_marco.py:152-240 — same flags, caches, printed report and --ofile contents."""
    assert osp.exists(args.mapping_file)
    args.alphas, args.betas, args.gammas = _floats(args.alphas), _floats(args.betas), _floats(args.gammas)
    args.recall_num = _floats(args.recall_num, int)
    args.gt_file = _resolve(args.gt_file, args.dir_path)
    args.ance_file = _resolve(args.ance_file, args.dir_path)
    args.fine_file, fexists = _resolve(args.fine_file, args.dir_path, True)
    args.coarse_file = _resolve(args.coarse_file, args.dir_path)
    fine_t = {"query": 0, "pred": 2, "score": 3}
    gts, _, _ = check_cache(args.gt_file, {"query": 0, "pred": -1})
    ance_preds, ance_scores, _ = check_cache(args.ance_file, fine_t)
    fine_preds = fine_scores = ranks_fine = None
    if fexists:
        fine_preds, fine_scores, _ = check_cache(args.fine_file, fine_t)
    _, _, coarse = check_cache(args.coarse_file, {"query": 0, "cluster": 1})
    with open(args.mapping_file, "rb") as fr:
        mapping = pickle.load(fr)
    ranks_gt, num = _rank_cache(_strip_ext(args.coarse_file) + "_cr4gt.pkl", ance_preds, coarse, mapping, None)
    if fexists:
        # the reference ranks the ANCE documents again here (ensemble_marco.py:199-209 iterate ance_preds)
        ranks_fine, num = _rank_cache(_strip_ext(args.fine_file) + "_cr.pkl", ance_preds, coarse, mapping, num)
    if args.ofile is not None:
        open(args.ofile, "w").close()
    evaluate_marco("ANCE Pred", args.recall_num, args.ofile, gts, ranks=ance_preds)
    if fexists:
        evaluate_marco("Fine Pred", args.recall_num, args.ofile, gts, ranks=fine_preds)
    _sweep(args, gts, ance_preds, ance_scores, ranks_gt, fexists, fine_preds, fine_scores, ranks_fine, num, evaluate_marco, gts)


def combine_main_nqdpr(args):
    """ensemble_nqdpr.py:152-251."""
    assert osp.exists(args.mapping_file)
    args.alphas, args.betas, args.gammas = _floats(args.alphas), _floats(args.betas), _floats(args.gammas)
    args.recall_num = _floats(args.recall_num, int)
    args.ance_file = _resolve(args.ance_file, args.dir_path)
    args.fine_file, fexists = _resolve(args.fine_file, args.dir_path, True)
    fine_t = {"query": 0, "pred": 2, "score": 3, "_by_line": True}
    offsets = np.memmap(osp.join(args.dir_path, "test_inverse_offsets.bin"), mode="r", dtype=np.int32)
    array = np.memmap(osp.join(args.dir_path, "test_inverse_array.bin"), mode="r", dtype=np.int32)
    nq_eval = (offsets, array)
    ance_preds, ance_scores, _ = check_cache(args.ance_file, fine_t)
    fine_preds = fine_scores = ranks_fine = None
    if fexists:
        fine_preds, fine_scores, _ = check_cache(args.fine_file, {"query": 0, "pred": 2, "score": 3}, args.ance_file)
    if args.ofile is not None:
        open(args.ofile, "w").close()
    evaluate_nqdpr("ANCE Pred", args.recall_num, args.ofile, nq_eval, ranks=ance_preds)
    if fexists:
        evaluate_nqdpr("Fine Pred", args.recall_num, args.ofile, nq_eval, ranks=fine_preds)
    if args.noensemble:
        return
    args.coarse_file = _resolve(args.coarse_file, args.dir_path)
    _, _, coarse = check_cache(args.coarse_file, {"query": 0, "cluster": 1}, args.ance_file)
    with open(args.mapping_file, "rb") as fr:
        mapping = pickle.load(fr)
    ranks_gt, num = _rank_cache(_strip_ext(args.coarse_file) + "_cr4gt.pkl", ance_preds, coarse, mapping, None)
    if fexists:
        ranks_fine, num = _rank_cache(_strip_ext(args.fine_file) + "_cr.pkl", ance_preds, coarse, mapping, num)
    _sweep(args, ance_preds, ance_preds, ance_scores, ranks_gt, fexists, fine_preds, fine_scores, ranks_fine, num,
           evaluate_nqdpr, nq_eval)
