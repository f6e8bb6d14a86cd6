"""Result fusion of MEVI/ensemble_marco.py and MEVI/ensemble_nqdpr.py.

Consumes the files the index hot path writes (faiss_search.to_file / hn result lines / coarse lines /
rqmapping pickle) and reproduces the reference's fusion and report text exactly:
    score' = score + alpha / (beta * crank + 1),   times (1 - gamma * alpha) when the document's RQ leaf
    is not among the query's beam-search leaves (crank == number of leaves)
(ensemble_marco.py:221-240, ensemble_nqdpr.py:232-251), recall / MRR / hit-rate bookkeeping
(ensemble_marco.py:8-72, ensemble_nqdpr.py:9-60), text-parse caches next to the inputs (130-140) and the
`_cr4gt.pkl` / `_cr.pkl` rank caches (176-209).

Two arithmetic paths behind the same drivers (`args.device`, CLI flag `--device`):
  * "cuda" (the CLI default): `DeviceFusion` — the candidate lists of all queries live on the GPU as dense
    [nq, P] arrays; leaf ranks (`mevi_ensemble_cluster_ranks`), fusion + de-duplication + ranking
    (`mevi_ensemble_fuse`, float64 with individually rounded operations) and the evaluators' list look-ups
    (`mevi_ensemble_positions`, `mevi_ensemble_first_hit`) are kernels of libmevi_b200.so; only the per-query
    hit positions come back, and the report is accumulated from them in the reference's order.  A whole
    (alpha, beta, gamma) sweep re-uses the uploaded lists.
  * "host": the reference's own per-query python dictionaries, kept as a mirror of the scripts (and used by the
    CPU-only tests); it is never chosen implicitly.
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


def evaluate_marco(scoring, recall_num, ofile, gts, scores=None, ranks=None, positions=None):
    """ensemble_marco.py:8-72: recall@n averaged over ground truths, MRR@n of the best-ranked ground truth.
    `positions[q]` (index of every ground truth in the ranking, None if absent) replaces the list look-ups when the
    ranking lives on the device."""
    recalls = {r: 0 for r in recall_num}
    mrrs = {r: 0 for r in recall_num}
    for q in gts:
        if positions is not None:
            vs = positions[q]
        else:
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


def evaluate_nqdpr(scoring, recall_num, ofile, nq_eval, scores=None, ranks=None, first_hits=None):
    """ensemble_nqdpr.py:9-60: a document is a hit for query qind if qind is in its inverse-answer list.
    `first_hits[qind]` (rank of the first hit or None) replaces the scan when the ranking lives on the device."""
    offsets, array = nq_eval
    src = first_hits if first_hits is not None else (scores if scores is not None else ranks)
    mrrs = {r: 0 for r in recall_num}
    hits = {r: 0 for r in recall_num}
    for qind in src.keys():
        if first_hits is not None:
            ind = first_hits[qind]
        else:
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


# ---- device path -------------------------------------------------------------------------------------
def _wants_device(args) -> bool:
    dev = getattr(args, "device", "cuda")
    if dev not in ("cuda", "host"):
        raise ValueError(f"--device must be 'cuda' or 'host', got {dev!r}")
    return dev == "cuda"


def mapping_to_codes(mapping):
    """rqmapping dictionary (doc -> tuple of M codes) -> dense int32 [max doc + 1, M]; documents that are not in the
    mapping get INT32_MIN rows, which the rank kernel reports as a missing key."""
    n = max(mapping) + 1 if len(mapping) else 0
    M = len(next(iter(mapping.values()))) if n else 1
    codes = np.full((n, M), np.iinfo(np.int32).min, dtype=np.int32)
    keys = np.fromiter(mapping.keys(), dtype=np.int64, count=len(mapping))
    vals = np.array(list(mapping.values()), dtype=np.int32).reshape(len(mapping), M)
    codes[keys] = vals
    return codes


class DeviceFusion:
    """The candidate lists of `keys` (query order of the evaluator) as dense device arrays, and the kernels over them.

    preds / scores: dictionaries query -> list (queries without an entry get an empty list, as the reference's
    `scores = {q: {} for q in gts}` does); ranks: query -> list of leaf ranks, or computed on the device by
    `cluster_ranks`.  The per-query length is the shortest of the three lists (python's zip)."""

    def __init__(self, keys, preds, scores=None, ranks=None, device=None):
        import torch

        from ._lib import get_context

        self.ctx = get_context(device)
        self.dev = torch.device("cuda", self.ctx.device)
        self.keys = list(keys)
        nq = len(self.keys)
        lens = np.zeros(nq, dtype=np.int32)
        for i, q in enumerate(self.keys):
            if q in preds:
                n = len(preds[q])
                if scores is not None:
                    n = min(n, len(scores[q]))
                if ranks is not None:
                    n = min(n, len(ranks[q]))
                lens[i] = n
        P = max(int(lens.max()) if nq else 0, 1)
        ids = np.full((nq, P), -1, dtype=np.int64)
        sc = np.zeros((nq, P), dtype=np.float64)
        cr = np.zeros((nq, P), dtype=np.int32)
        for i, q in enumerate(self.keys):
            n = lens[i]
            if n:
                ids[i, :n] = preds[q][:n]
                if scores is not None:
                    sc[i, :n] = scores[q][:n]
                if ranks is not None:
                    cr[i, :n] = ranks[q][:n]
        self.P = P
        self.ids = torch.from_numpy(ids).to(self.dev)
        self.scores = torch.from_numpy(sc).to(self.dev)
        self.count = torch.from_numpy(lens).to(self.dev)
        self.cranks = torch.from_numpy(cr).to(self.dev) if ranks is not None else None

    def cluster_ranks(self, coarse_clusters, codes):
        """ensemble_marco.py:181-191 on the device.  coarse_clusters: query -> ordered list of leaves (lists of M
        codes); codes: int32 [N, M] (numpy or CUDA tensor; `mapping_to_codes(rqmapping)` or the encoder's output).
        Returns (ranks dictionary query -> list, number of distinct leaves) in the format of the reference's caches."""
        import torch

        if not isinstance(codes, torch.Tensor):
            codes = torch.from_numpy(np.ascontiguousarray(codes, dtype=np.int32))
        codes = codes.to(self.dev)
        leaves = np.array([coarse_clusters[q] for q in self.keys], dtype=np.int32)
        if leaves.ndim != 3 or leaves.shape[2] != codes.shape[1]:
            raise ValueError(f"leaf lists must be [nq, L, M={codes.shape[1]}], got {leaves.shape}")
        cranks, num = self.ctx.ensemble_cluster_ranks(self.ids, self.count, codes, torch.from_numpy(leaves).to(self.dev))
        num = num.cpu().numpy()
        if len(num) and not (num == num[0]).all():
            raise AssertionError("queries must all carry the same number of leaves")
        host = cranks.cpu().numpy()
        lens = self.count.cpu().numpy()
        if any((host[i, :lens[i]] == -2).any() for i in range(len(lens))):
            raise KeyError("a predicted document is not in the rqmapping")
        self.cranks = cranks
        return {q: host[i, :lens[i]].tolist() for i, q in enumerate(self.keys)}, (int(num[0]) if len(num) else None)

    def fuse(self, alpha, beta, gamma, num_leaves):
        """ensemble_marco.py:233-237 + the ranking of its evaluator: (ranked ids, fused scores, counts) on the device."""
        return self.ctx.ensemble_fuse(self.ids, self.scores, self.cranks, self.count, alpha, beta, gamma, num_leaves)

    def positions(self, ranked, ranked_count, truth):
        """truth: query -> list of ground-truth documents; -> query -> [index or None, ...]."""
        import torch

        G = max(max((len(truth[q]) for q in self.keys), default=0), 1)
        tg = np.full((len(self.keys), G), -1, dtype=np.int64)
        tc = np.zeros(len(self.keys), dtype=np.int32)
        for i, q in enumerate(self.keys):
            tc[i] = len(truth[q])
            tg[i, :tc[i]] = truth[q]
        pos = self.ctx.ensemble_positions(ranked, ranked_count, torch.from_numpy(tg).to(self.dev),
                                          torch.from_numpy(tc).to(self.dev)).cpu().numpy()
        return {q: [int(v) if v >= 0 else None for v in pos[i, :tc[i]]] for i, q in enumerate(self.keys)}

    def first_hits(self, ranked, ranked_count, offsets, array):
        """NQ-DPR: query (an integer line index) -> rank of the first document that answers it, or None."""
        import torch

        qidx = torch.tensor([int(q) for q in self.keys], dtype=torch.int64, device=self.dev)
        if not isinstance(offsets, torch.Tensor):
            offsets = torch.from_numpy(np.ascontiguousarray(offsets, dtype=np.int32)).to(self.dev)
            array = torch.from_numpy(np.ascontiguousarray(array, dtype=np.int32)).to(self.dev)
        fh = self.ctx.ensemble_first_hit(ranked, ranked_count, qidx, offsets, array).cpu().numpy()
        return {q: (int(v) if v >= 0 else None) for q, v in zip(self.keys, fh)}


def _device_rank_cache(path, fusion, coarse, codes, num_clusters):
    """`_rank_cache` with the ranks computed by the device when the cache file is not there yet."""
    if osp.exists(path):
        with open(path, "rb") as fr:
            ranks, num = pickle.load(fr)
    else:
        ranks, num = fusion.cluster_ranks(coarse, codes)
        with open(path, "wb") as fw:
            pickle.dump((ranks, num), fw)
    assert num_clusters in (None, num)
    return ranks, num


def _sweep_device(args, keys, ance_preds, ance_scores, ranks_gt, fexists, fine_preds, fine_scores, ranks_fine,
                  num_clusters, report):
    """`_sweep` with the lists resident on the device: one upload, one fuse launch per (alpha, beta, gamma)."""
    preds, scores, ranks = ance_preds, ance_scores, ranks_gt
    if fexists:
        preds = {q: ance_preds[q] + fine_preds[q] for q in ance_preds}
        scores = {q: ance_scores[q] + fine_scores[q] for q in ance_preds}
        ranks = {q: list(ranks_gt[q]) + list(ranks_fine[q]) for q in ance_preds}
    fusion = DeviceFusion(keys, preds, scores, ranks)
    for alpha in args.alphas:
        for beta in args.betas:
            for gamma in args.gammas:
                ranked, _, counts = fusion.fuse(alpha, beta, gamma, num_clusters)
                report(f"score + {alpha} / ({beta} * crank + 1); punishment (1 - {gamma} * {alpha})", fusion, ranked, counts)


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
    if _wants_device(args):
        return _combine_marco_device(args, gts, ance_preds, ance_scores, fexists, fine_preds, fine_scores, coarse, mapping)
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


def _combine_marco_device(args, gts, ance_preds, ance_scores, fexists, fine_preds, fine_scores, coarse, mapping):
    """The arithmetic of combine_main_marco on the GPU; files, caches and report text as the host path."""
    codes = mapping_to_codes(mapping)
    ance = DeviceFusion(list(ance_preds), ance_preds)  # the reference ranks the ANCE lists, query order of the file
    ranks_gt, num = _device_rank_cache(_strip_ext(args.coarse_file) + "_cr4gt.pkl", ance, coarse, codes, None)
    ranks_fine = None
    if fexists:
        ranks_fine, num = _device_rank_cache(_strip_ext(args.fine_file) + "_cr.pkl", ance, coarse, codes, num)
    if args.ofile is not None:
        open(args.ofile, "w").close()
    keys = list(gts)
    raw = DeviceFusion(keys, ance_preds)
    evaluate_marco("ANCE Pred", args.recall_num, args.ofile, gts, positions=raw.positions(raw.ids, raw.count, gts))
    if fexists:
        raw = DeviceFusion(keys, fine_preds)
        evaluate_marco("Fine Pred", args.recall_num, args.ofile, gts, positions=raw.positions(raw.ids, raw.count, gts))

    def report(scoring, fusion, ranked, counts):
        evaluate_marco(scoring, args.recall_num, args.ofile, gts, positions=fusion.positions(ranked, counts, gts))

    _sweep_device(args, keys, ance_preds, ance_scores, ranks_gt, fexists, fine_preds, fine_scores, ranks_fine, num, report)


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
    if _wants_device(args):
        return _combine_nqdpr_device(args, nq_eval, ance_preds, ance_scores, fexists, fine_preds, fine_scores)
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


def _combine_nqdpr_device(args, nq_eval, ance_preds, ance_scores, fexists, fine_preds, fine_scores):
    """The arithmetic of combine_main_nqdpr on the GPU; files, caches and report text as the host path."""
    import torch

    offsets, array = nq_eval
    keys = list(ance_preds)
    ance = DeviceFusion(keys, ance_preds)
    d_off = torch.from_numpy(np.ascontiguousarray(offsets, dtype=np.int32)).to(ance.dev)
    d_arr = torch.from_numpy(np.ascontiguousarray(array, dtype=np.int32)).to(ance.dev)
    evaluate_nqdpr("ANCE Pred", args.recall_num, args.ofile, nq_eval,
                   first_hits=ance.first_hits(ance.ids, ance.count, d_off, d_arr))
    if fexists:
        raw = DeviceFusion(list(fine_preds), fine_preds)
        evaluate_nqdpr("Fine Pred", args.recall_num, args.ofile, nq_eval,
                       first_hits=raw.first_hits(raw.ids, raw.count, d_off, d_arr))
    if args.noensemble:
        return
    args.coarse_file = _resolve(args.coarse_file, args.dir_path)
    _, _, coarse = check_cache(args.coarse_file, {"query": 0, "cluster": 1}, args.ance_file)
    with open(args.mapping_file, "rb") as fr:
        mapping = pickle.load(fr)
    codes = mapping_to_codes(mapping)
    ranks_gt, num = _device_rank_cache(_strip_ext(args.coarse_file) + "_cr4gt.pkl", ance, coarse, codes, None)
    ranks_fine = None
    if fexists:
        ranks_fine, num = _device_rank_cache(_strip_ext(args.fine_file) + "_cr.pkl", ance, coarse, codes, num)

    def report(scoring, fusion, ranked, counts):
        evaluate_nqdpr(scoring, args.recall_num, args.ofile, nq_eval,
                       first_hits=fusion.first_hits(ranked, counts, d_off, d_arr))

    _sweep_device(args, keys, ance_preds, ance_scores, ranks_gt, fexists, fine_preds, fine_scores, ranks_fine, num, report)
