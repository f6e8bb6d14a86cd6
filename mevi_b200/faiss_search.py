"""Drop-in for MEVI/faiss_search.py with param='Flat' (exact inner-product search).

`read`, `search`, `to_file` and the CLI flags keep the reference's signatures and
file formats (faiss_search.py:9-21, 71-98).  `search` runs the exact flat search
on the GPU(s) through `mevi_flat_ip_topk`; approximate index strings (HNSW*,
IVF*) are out of scope and raise.
"""
from __future__ import annotations

import argparse
from typing import Optional

import numpy as np
import torch

from . import _lib
from .dist_utils import all_gather_stack, dist_on, rank_world, shard_bounds


def read(path, dim):  # faiss_search.py:9-10
    return np.fromfile(path, dtype=np.float32).reshape(-1, dim)


def _is_flat(param: str) -> bool:
    return param.strip().lower() in ("flat", "flatip", "idmap,flat")


def _is_approximate(param: str) -> bool:
    """IVF*/HNSW*/PQ* factory strings: approximate faiss indexes whose result is a SUBSET of the exact one."""
    p = param.strip().lower()
    return any(tok.startswith(("ivf", "hnsw", "pq", "opq", "lsh", "imi")) for tok in p.split(","))


FAISS_PAD_SCORE = float(np.finfo(np.float32).min)  # faiss pads inner-product results with -3.4028235e38, id -1


class FlatIndex:
    """faiss `IndexFlatIP` as the reference uses it (faiss_search.py:14-20): `add(doc)` once, `search(query, k)` many.
    Documents live on the device in pieces of `piece_rows` rows, each with its persistent fp16 tile image
    (`mevi_flat_index_create`); a search runs the tensor-core prefilter + exact fp32 re-score per piece and merges.
    With torch.distributed initialised the rows are sharded (pq.py:218-225 rule): every rank adds its block, and a
    search ends with the all-gather of the per-shard lists + merge."""

    def __init__(self, dim: int, device_index: Optional[int] = None, piece_rows: int = 1 << 22, mode: str = "auto"):
        self.ctx = _lib.get_context(device_index)
        self.dev = torch.device("cuda", self.ctx.device)
        self.dim, self.piece_rows, self.mode = int(dim), int(piece_rows), mode
        self.pieces = []  # (id_base, tensor, handle)
        self.ntotal = 0
        self.is_trained = True

    def add(self, doc):
        assert doc.shape[1] == self.dim
        N = doc.shape[0]
        rank, world = rank_world()
        start, end = shard_bounds(N, rank, world)
        base = self.ntotal
        for a in range(start, end, self.piece_rows):
            b = min(a + self.piece_rows, end)
            if isinstance(doc, torch.Tensor):
                piece = doc[a:b].to(device=self.dev, dtype=torch.float32).contiguous()
            else:
                piece = torch.from_numpy(np.ascontiguousarray(doc[a:b], dtype=np.float32)).to(self.dev)
            self.pieces.append((base + a, piece, self.ctx.flat_index_create(piece)))
        self.ntotal += N

    @torch.no_grad()
    def search_device(self, Q, topk: int):
        """-> (scores [nq,k] fp32 descending, ids [nq,k] int64) on the device, -inf / -1 padded."""
        fan_in = max(2, 16384 // max(topk, 1))  # mevi_topk_merge sorts at most 16,384 entries per query
        parts_s, parts_i = [], []
        for id_base, piece, handle in self.pieces:
            s, i = self.ctx.flat_index_search(handle, Q, topk, id_base=id_base, mode=self.mode)
            parts_s.append(s)
            parts_i.append(i)
            if len(parts_s) == fan_in:
                s, i = self.ctx.topk_merge(torch.stack(parts_s).contiguous(), torch.stack(parts_i).contiguous())
                parts_s, parts_i = [s], [i]
        if not parts_s:
            parts_s = [torch.full((Q.shape[0], topk), float("-inf"), dtype=torch.float32, device=self.dev)]
            parts_i = [torch.full((Q.shape[0], topk), -1, dtype=torch.int64, device=self.dev)]
        run_s, run_i = (parts_s[0], parts_i[0]) if len(parts_s) == 1 else self.ctx.topk_merge(
            torch.stack(parts_s).contiguous(), torch.stack(parts_i).contiguous())
        if dist_on():
            run_s, run_i = self.ctx.topk_merge(all_gather_stack(run_s).contiguous(), all_gather_stack(run_i).contiguous())
        return run_s, run_i

    def search(self, query, topk: int):
        """faiss signature: (dists float32 [nq,topk] descending, indices int64 [nq,topk]) as numpy arrays; missing
        results are padded like faiss does (lowest float, id -1)."""
        Q = query if isinstance(query, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(query, dtype=np.float32))
        Q = Q.to(device=self.dev, dtype=torch.float32).contiguous()
        assert Q.shape[1] == self.dim
        run_s, run_i = self.search_device(Q, int(topk))
        dists, indices = run_s.cpu().numpy(), run_i.cpu().numpy()
        dists[indices < 0] = FAISS_PAD_SCORE
        return dists, indices

    def close(self):
        for _, _, handle in self.pieces:
            self.ctx.flat_index_destroy(handle)
        self.pieces = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@torch.no_grad()
def search(query, doc, dim, topk, param, piece_rows: int = 1 << 22, mode: str = "auto",
           device_index: Optional[int] = None):
    """faiss_search.py:13-21: returns (dists float32 [nq,topk] descending, indices int64 [nq,topk],
    -1 / lowest-float padded when the index holds fewer than topk vectors).
    `doc` may be a host array (streamed to the device in pieces of `piece_rows`) or a CUDA tensor.
    With torch.distributed initialised the documents are row-sharded (pq.py:218-225 rule), each
    rank searches its block and the per-shard lists are all-gathered and merged."""
    if not _is_flat(param):
        if not _is_approximate(param):
            raise NotImplementedError(f"param={param!r}: not a faiss index string this module understands")
        # The reference CLI defaults to 'IVF100,Flat' (faiss_search.py:88).  An approximate index returns a subset of
        # the exact neighbours; the exact search below is always at least as good and on a B200 faster than training one.
        print(f"Param {param}: approximate faiss indexes are not built here; running the exact 'Flat' search instead.")
    assert doc.shape[1] == dim
    index = FlatIndex(dim, device_index=device_index, piece_rows=piece_rows, mode=mode)  # faiss.index_factory(...)
    print(f"Param {param} trained: {index.is_trained}.")
    index.add(doc)                                                                         # index.add(doc)
    try:
        return index.search(query, topk)                                                   # index.search(query, topk)
    finally:
        index.close()


def to_file(query_path, output_path, dists, indices):  # faiss_search.py:71-77
    with open(query_path, "r") as fr, open(output_path, "w") as fw:
        for i, line in enumerate(fr):
            query = line.split("\t")[0]
            preds = ",".join([str(ind) for ind in indices[i].tolist()])
            scores = ",".join([str(sco) for sco in dists[i].tolist()])
            print(f"{query}\t\t{preds}\t{scores}", file=fw)


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--query_path", type=str, required=True)
    parser.add_argument("--doc_path", type=str, required=True)
    parser.add_argument("--output_path", type=str, required=True)
    parser.add_argument("--raw_query_path", type=str, required=True)
    parser.add_argument("--dim", type=int, default=768)
    parser.add_argument("--topk", type=int, default=1000)
    parser.add_argument("--param", type=str, default="IVF100,Flat")
    args = parser.parse_args(argv)
    query = read(args.query_path, args.dim)
    doc = np.memmap(args.doc_path, dtype=np.float32, mode="r").reshape(-1, args.dim)
    dists, indices = search(query, doc, args.dim, args.topk, args.param)
    print(indices.dtype, indices.shape, dists.dtype, dists.shape)
    to_file(args.raw_query_path, args.output_path, dists, indices)


if __name__ == "__main__":
    main()
