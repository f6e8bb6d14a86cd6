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


@torch.no_grad()
def search(query, doc, dim, topk, param, piece_rows: int = 1 << 21, mode: str = "auto",
           device_index: Optional[int] = None):
    """faiss_search.py:13-21: returns (dists float32 [nq,topk] descending, indices int64 [nq,topk],
    -1 / -inf padded when the index holds fewer than topk vectors).
    `doc` may be a host array (streamed to the device in pieces of `piece_rows`) or a CUDA tensor.
    With torch.distributed initialised the documents are row-sharded (pq.py:218-225 rule), each
    rank searches its block and the per-shard lists are all-gathered and merged."""
    if not _is_flat(param):
        if not _is_approximate(param):
            raise NotImplementedError(f"param={param!r}: not a faiss index string this module understands")
        # The reference CLI defaults to 'IVF100,Flat' (faiss_search.py:88).  An approximate index returns a subset of
        # the exact neighbours; the exact search below is always at least as good and on a B200 faster than training one.
        print(f"Param {param}: approximate faiss indexes are not built here; running the exact 'Flat' search instead.")
    ctx = _lib.get_context(device_index)
    dev = torch.device("cuda", ctx.device)
    print(f"Param {param} trained: True.")
    Q = query if isinstance(query, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(query, dtype=np.float32))
    Q = Q.to(device=dev, dtype=torch.float32).contiguous()
    assert Q.shape[1] == dim and doc.shape[1] == dim
    N = doc.shape[0]
    rank, world = rank_world()
    start, end = shard_bounds(N, rank, world)
    run_s = run_i = None
    for a in range(start, end, piece_rows) if end > start else []:
        b = min(a + piece_rows, end)
        if isinstance(doc, torch.Tensor):
            piece = doc[a:b].to(device=dev, dtype=torch.float32).contiguous()
        else:
            piece = torch.from_numpy(np.ascontiguousarray(doc[a:b], dtype=np.float32)).to(dev)
        s, i = ctx.flat_ip_topk(Q, piece, topk, id_base=a, mode=mode)
        if run_s is None:
            run_s, run_i = s, i
        else:
            run_s, run_i = ctx.topk_merge(torch.stack([run_s, s]), torch.stack([run_i, i]))
        del piece
    if run_s is None:
        run_s = torch.full((Q.shape[0], topk), float("-inf"), dtype=torch.float32, device=dev)
        run_i = torch.full((Q.shape[0], topk), -1, dtype=torch.int64, device=dev)
    if dist_on():
        run_s, run_i = ctx.topk_merge(all_gather_stack(run_s).contiguous(), all_gather_stack(run_i).contiguous())
    dists, indices = run_s.cpu().numpy(), run_i.cpu().numpy()
    dists[indices < 0] = FAISS_PAD_SCORE  # the library pads with -inf; faiss with the lowest float (to_file prints it)
    return dists, indices


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
