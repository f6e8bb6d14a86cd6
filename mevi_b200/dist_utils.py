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
