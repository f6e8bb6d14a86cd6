"""Host-side sharding helpers (one process per GPU, torch.distributed for plumbing)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def dist_on() -> bool:
    return dist.is_available() and dist.is_initialized()


def rank_world() -> Tuple[int, int]:
    if dist_on():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(num_docs: int, rank: int, nrank: int) -> Tuple[int, int]:
    """Contiguous row block of `rank`: N//nrank rows each, the last rank takes the
    remainder — exactly MEVI/pq.py:218-225 (also 724-731, main_models.py:3092-3098)."""
    per = num_docs // nrank
    start = per * rank
    ending = num_docs if rank + 1 == nrank else start + per
    return start, ending


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if dist_on():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def all_gather_stack(t: torch.Tensor) -> torch.Tensor:
    """[..] -> [world, ..] (same shape on every rank)."""
    if not dist_on():
        return t.unsqueeze(0)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t.contiguous())
    return torch.stack(out, dim=0)


def gather_rows_to_rank0(t: torch.Tensor, counts) -> "torch.Tensor | None":
    """Concatenate per-rank row blocks (possibly different lengths) on rank 0, in rank order —
    the collective form of the reference's /tmp part-file merge (main_models.py:289-310)."""
    if not dist_on():
        return t
    rank, world = dist.get_rank(), dist.get_world_size()
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    if rank != 0:
        return None
    return torch.cat([o[: counts[r]] for r, o in enumerate(out)], dim=0)


def all_gather_varlen(t: torch.Tensor) -> torch.Tensor:
    """Concatenation, in rank order, of every rank's 1-D (or row-block) tensor of differing length, on every rank."""
    if not dist_on():
        return t
    world = dist.get_world_size()
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(v.item()) for v in sizes]
    mx = max(max(sizes), 1)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:c] for o, c in zip(out, sizes)], dim=0)


def partition_rows_by_leaf(X: torch.Tensor, codes: torch.Tensor, K: int, id_base: int, gather_rows=None):
    """Re-shard a row-block-sharded corpus so that every RQ leaf lives on ONE rank (the re-rank's natural partition:
    the row blocks of pq.py:218-225 give every rank a slice of every leaf, so the per-rank work of a leaf-grouped call
    stops shrinking with the number of ranks).  Leaves, in ascending key order, are cut into `world` contiguous ranges
    of (nearly) equal row counts; rows, their codes and their GLOBAL document ids travel in one all-to-all each.
      X [n,d] fp32, codes [n,M] int32 (this rank's block, document ids id_base .. id_base+n)
      -> (X_own [m,d], codes_own [m,M], doc_ids_own int64 [m])
    `gather_rows(X, int32 index) -> rows` may be the library's row gather (a device kernel); torch indexing otherwise."""
    rank, world = rank_world()
    dev = X.device
    n, M = codes.shape
    gids = torch.arange(id_base, id_base + n, dtype=torch.int64, device=dev)
    if world == 1:
        return X, codes, gids
    key = torch.zeros(n, dtype=torch.int64, device=dev)
    for j in range(M):
        key = key * K + codes[:, j].long()
    uk, cnt = torch.unique(key, return_counts=True)
    gk, inv = torch.unique(all_gather_varlen(uk), return_inverse=True)          # every leaf of the corpus, ascending
    gcnt = torch.zeros(gk.numel(), dtype=torch.int64, device=dev).index_add_(0, inv, all_gather_varlen(cnt))
    first = torch.cumsum(gcnt, 0) - gcnt                                         # corpus position of a leaf's first row
    owner = torch.clamp(first * world // max(int(gcnt.sum().item()), 1), max=world - 1)
    dest = owner[torch.searchsorted(gk, key)]
    order = torch.argsort(dest, stable=True)
    send = torch.bincount(dest, minlength=world)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)
    send_l, recv_l = [int(v) for v in send.tolist()], [int(v) for v in recv.tolist()]
    m = sum(recv_l)

    def exchange(t_sorted):
        out = torch.empty((m,) + tuple(t_sorted.shape[1:]), dtype=t_sorted.dtype, device=dev)
        dist.all_to_all_single(out, t_sorted.contiguous(), output_split_sizes=recv_l, input_split_sizes=send_l)
        return out

    X_sorted = gather_rows(X, order.to(torch.int32)) if gather_rows is not None else X[order]
    X_own = exchange(X_sorted)
    del X_sorted
    return X_own, exchange(codes[order]), exchange(gids[order])
