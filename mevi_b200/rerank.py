"""Cluster-restricted re-rank: the loop of MEVI/main_models.py:3911-4053 as one kernel launch.

The reference walks, per query, the RQ leaves emitted by NCI beam search,
looks each leaf up in the `rqclus` dictionary (3928), fancy-indexes the
candidate rows out of the doc-embedding memmap (3944 / IndexedData 1011-1017),
copies them to the GPU in chunks of `encode_batch_size` (3948-3952), scores
q.P^T (3967-3968), concatenates and sorts (4012-4014), then prints the hn /
fine lines (4046-4053, 4082).  Here the inverted lists live on the device as
CSR, all queries go through `mevi_cluster_rerank` at once, and only the k best
come back.  Text output keeps the reference's byte format.
"""
from __future__ import annotations

import pickle
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .dist_utils import all_gather_stack, dist_on, rank_world, shard_bounds


class ClusterIndex:
    """Inverted lists leaf -> doc ids on the device (CSR over the non-empty leaves).

    leaf_keys    int64 [n_leaves]   sorted base-K value of the code tuple
    leaf_offsets int64 [n_leaves+1]
    leaf_docids  int32 [n]          local row indices, ascending inside a leaf
    """

    def __init__(self, leaf_keys, leaf_offsets, leaf_docids, M: int, K: int, id_base: int = 0):
        self.leaf_keys, self.leaf_offsets, self.leaf_docids = leaf_keys, leaf_offsets, leaf_docids
        self.M, self.K, self.id_base = int(M), int(K), int(id_base)

    @property
    def n_leaves(self) -> int:
        return int(self.leaf_keys.numel())

    @property
    def device(self):
        return self.leaf_keys.device

    # -- construction -------------------------------------------------------
    @classmethod
    def from_codes(cls, codes, K: int, id_base: int = 0, device_index: Optional[int] = None) -> "ClusterIndex":
        """Replaces the dict loops of pq.py:236-242 / 200-214 with a device sort by leaf key."""
        ctx = _lib.get_context(device_index)
        dev = torch.device("cuda", ctx.device)
        if not isinstance(codes, torch.Tensor):
            codes = torch.from_numpy(np.ascontiguousarray(codes, dtype=np.int32))
        codes = codes.to(device=dev, dtype=torch.int32).contiguous()
        M = codes.shape[1]
        docids, keys = ctx.build_inverted_lists(codes, K)
        leaf_keys, counts = torch.unique_consecutive(keys, return_counts=True)
        offsets = torch.zeros(leaf_keys.numel() + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, 0, out=offsets[1:])
        return cls(leaf_keys, offsets, docids, M, K, id_base)

    @classmethod
    def from_cluster_dict(cls, doc_cluster: Dict[Tuple[int, ...], Sequence[int]], K: int, id_base: int = 0,
                          device_index: Optional[int] = None, row_range: Optional[Tuple[int, int]] = None) -> "ClusterIndex":
        """Load a reference-format `rqclus*.pkl` dictionary (main_models.py:3214-3219).  With
        `row_range=(start, end)` only doc ids of that row block are kept (doc-sharded re-rank)."""
        ctx = _lib.get_context(device_index)
        dev = torch.device("cuda", ctx.device)
        M = len(next(iter(doc_cluster))) if doc_cluster else 1
        keys, lens, ids = [], [], []
        for t, docs in doc_cluster.items():
            key = 0
            for c in t:
                key = key * K + int(c)
            d = np.asarray(docs, dtype=np.int64)
            if row_range is not None:
                d = d[(d >= row_range[0]) & (d < row_range[1])]
            if d.size == 0:
                continue
            keys.append(key)
            lens.append(d.size)
            ids.append(np.sort(d) - id_base)
        order = np.argsort(np.asarray(keys, dtype=np.int64), kind="stable")
        keys_s = np.asarray(keys, dtype=np.int64)[order]
        lens_s = np.asarray(lens, dtype=np.int64)[order]
        docids = np.concatenate([ids[i] for i in order]).astype(np.int32) if len(ids) else np.zeros(0, np.int32)
        offsets = np.zeros(len(keys_s) + 1, dtype=np.int64)
        np.cumsum(lens_s, out=offsets[1:])
        return cls(torch.from_numpy(keys_s).to(dev), torch.from_numpy(offsets).to(dev), torch.from_numpy(docids).to(dev),
                   M, K, id_base)

    # -- queries --------------------------------------------------------------
    def lookup(self, dec) -> torch.Tensor:
        """Beam-search leaves [nq, L, M] (any int dtype, host or device) -> CSR leaf index int32
        [nq, L], -1 where the leaf holds no document (`doc_cluster.get(d, None)`, 3928)."""
        if not isinstance(dec, torch.Tensor):
            dec = torch.from_numpy(np.asarray(dec))
        dec = dec.to(self.device, torch.int64)
        key = torch.zeros(dec.shape[:-1], dtype=torch.int64, device=self.device)
        valid = torch.ones(dec.shape[:-1], dtype=torch.bool, device=self.device)
        for j in range(dec.shape[-1]):
            c = dec[..., j]
            valid &= (c >= 0) & (c < self.K)
            key = key * self.K + c.clamp(0, self.K - 1)
        if self.n_leaves == 0:
            return torch.full(key.shape, -1, dtype=torch.int32, device=self.device)
        pos = torch.searchsorted(self.leaf_keys, key).clamp(max=self.n_leaves - 1)
        hit = (self.leaf_keys[pos] == key) & valid
        return torch.where(hit, pos, torch.full_like(pos, -1)).to(torch.int32).contiguous()

    # -- reference-format dictionaries ---------------------------------------
    def to_dicts(self):
        """(rqclus, rqmapping) python dictionaries in the reference's format (global doc ids)."""
        keys = self.leaf_keys.cpu().numpy()
        offs = self.leaf_offsets.cpu().numpy()
        docs = self.leaf_docids.cpu().numpy().astype(np.int64) + self.id_base
        cluster, mapping = {}, {}
        for i, key in enumerate(keys.tolist()):
            t = []
            for _ in range(self.M):
                t.append(key % self.K)
                key //= self.K
            t = tuple(reversed(t))
            lst = docs[offs[i] : offs[i + 1]].tolist()
            cluster[t] = lst
            for dd in lst:
                mapping[dd] = t
        return cluster, mapping


class ClusterReranker:
    """Holds a (shard of the) doc-embedding matrix on the device plus its inverted lists."""

    def __init__(self, all_embeddings, index: ClusterIndex, device_index: Optional[int] = None,
                 leaf_ordered: bool = True):
        """`leaf_ordered=True` (default) keeps a copy of the matrix permuted into CSR order so that each
        leaf is one contiguous byte range (streamed with bulk async copies); False gathers candidate
        rows one by one from the document-ordered matrix."""
        self.ctx = _lib.get_context(device_index if device_index is not None else index.device.index)
        dev = torch.device("cuda", self.ctx.device)
        if isinstance(all_embeddings, torch.Tensor):
            D = all_embeddings.to(device=dev, dtype=torch.float32).contiguous()
        else:
            from .trainer import _upload_rows

            D = _upload_rows(all_embeddings, 0, all_embeddings.shape[0], dev)
        self.index = index
        self.leaf_ordered = bool(leaf_ordered)
        self.D = self.ctx.gather_rows(D, index.leaf_docids) if self.leaf_ordered else D

    @torch.no_grad()
    def rerank(self, query_embedding, dec, topk: int = 100):
        """-> (scores [nq,k] fp32 desc, ids [nq,k] int64 (-1 padded), n_candidates [nq] int32), on the device.
        With torch.distributed initialised the documents are sharded: every rank scores the
        candidates it owns, the per-shard lists are all-gathered and merged on every rank."""
        dev = self.D.device
        if not isinstance(query_embedding, torch.Tensor):
            query_embedding = torch.from_numpy(np.ascontiguousarray(query_embedding, dtype=np.float32))
        Q = query_embedding.to(device=dev, dtype=torch.float32).contiguous()
        ql = self.index.lookup(dec)
        scores, ids, ncand = self.ctx.cluster_rerank(Q, self.D, self.index.leaf_offsets, self.index.leaf_docids, ql, topk,
                                                     id_base=self.index.id_base, leaf_ordered=self.leaf_ordered)
        if dist_on():
            s_all = all_gather_stack(scores)
            i_all = all_gather_stack(ids)
            scores, ids = self.ctx.topk_merge(s_all.contiguous(), i_all.contiguous())
            torch.distributed.all_reduce(ncand)
        return scores, ids, ncand


def hn_lines(texts: Sequence[str], scores, ids, gt_outputs: Optional[Sequence[str]] = None) -> List[str]:
    """The hard-negative / fine result lines of main_models.py:4046-4053 (LogTxtFile joins the
    four fields with tabs, 254-257): query, gt scores, 'd1,d2,...', 's1,s2,...' where each score
    is str(tensor.item()) — the fp32 value widened to a python float."""
    s = scores.detach().cpu().numpy() if isinstance(scores, torch.Tensor) else np.asarray(scores)
    i = ids.detach().cpu().numpy() if isinstance(ids, torch.Tensor) else np.asarray(ids)
    lines = []
    for r, text in enumerate(texts):
        keep = i[r] >= 0
        docs = ",".join(str(int(x)) for x in i[r][keep])
        sc = ",".join(str(float(x)) for x in s[r][keep])
        gt = gt_outputs[r] if gt_outputs is not None else ""
        lines.append("\t".join([text, gt, docs, sc]))
    return lines


def write_hn_file(path: str, texts, scores, ids, gt_outputs=None) -> None:
    with open(path, "w") as fw:
        for line in hn_lines(texts, scores, ids, gt_outputs):
            print(line, file=fw)


def save_index_files(cluster_path: str, doc_cluster: dict, mapping: dict) -> None:
    """rqclus*.pkl / rqmapping*.pkl exactly as main_models.py:3198-3203 writes them
    (mapping path = cluster path with 'clus' -> 'mapping', 3192-3193)."""
    with open(cluster_path, "wb") as fw:
        pickle.dump(doc_cluster, fw)
    with open(cluster_path.replace("clus", "mapping"), "wb") as fw:
        pickle.dump(mapping, fw)


@torch.no_grad()
def eval_all_documents(query_embedding, all_embeddings, pool_size: int, batch_size: int = 1 << 20,
                       device_index: Optional[int] = None):
    """The `--eval_all_documents` branch of MEVI/main_models.py:3818-3876 (twin tower): the reference walks the
    corpus in blocks of `encode_batch_size`, scores q @ p^T and keeps a running torch.topk pool over the
    concatenation.  Here every block goes through `mevi_flat_ip_topk` (tcgen05 prefilter + exact fp32
    re-score) with `id_base` = the block's first row, and the per-block lists are merged by
    `mevi_topk_merge`.  `all_embeddings` may be a CUDA tensor (scored where it lies), a host tensor or an
    np.ndarray / np.memmap (streamed block by block).  With torch.distributed initialised the rows are
    sharded as pq.py:218-225 and the per-shard pools are all-gathered and merged on every rank.
    Returns (stack_scores [nq, min(pool_size, N)] fp32 descending, sorted_docs int32) on the device,
    like the reference's tensors."""
    ctx = _lib.get_context(device_index)
    dev = torch.device("cuda", ctx.device)
    if not isinstance(query_embedding, torch.Tensor):
        query_embedding = torch.from_numpy(np.ascontiguousarray(query_embedding, dtype=np.float32))
    Q = query_embedding.to(device=dev, dtype=torch.float32).contiguous()
    N = all_embeddings.shape[0]
    rank, world = rank_world()
    start, end = shard_bounds(N, rank, world) if dist_on() else (0, N)
    k = int(min(pool_size, N))
    fan_in = max(2, 16384 // max(k, 1))  # mevi_topk_merge sorts at most 16,384 entries per query
    parts_s, parts_i = [], []
    for a in range(start, end, batch_size):
        b = min(a + batch_size, end)
        if isinstance(all_embeddings, torch.Tensor):
            blk = all_embeddings[a:b].to(device=dev, dtype=torch.float32).contiguous()
        else:
            blk = torch.from_numpy(np.ascontiguousarray(all_embeddings[a:b], dtype=np.float32)).to(dev)
        s, i = ctx.flat_ip_topk(Q, blk, k, id_base=a)
        parts_s.append(s)
        parts_i.append(i)
        if len(parts_s) == fan_in:
            s, i = ctx.topk_merge(torch.stack(parts_s).contiguous(), torch.stack(parts_i).contiguous())
            parts_s, parts_i = [s], [i]
    if not parts_s:
        parts_s = [torch.full((Q.shape[0], k), float("-inf"), device=dev)]
        parts_i = [torch.full((Q.shape[0], k), -1, dtype=torch.int64, device=dev)]
    scores, ids = (parts_s[0], parts_i[0]) if len(parts_s) == 1 else ctx.topk_merge(
        torch.stack(parts_s).contiguous(), torch.stack(parts_i).contiguous())
    if dist_on():
        scores, ids = ctx.topk_merge(all_gather_stack(scores).contiguous(), all_gather_stack(ids).contiguous())
    return scores, ids.to(torch.int32)
