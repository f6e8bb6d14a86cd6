"""Data-parallel full-batch Lloyd trainer for the RQ codebook.

Replaces the rq branch of MEVI/pq.py:550-598 (sklearn MiniBatchKMeans on rank 0,
residual subtraction in numpy) with the sharded form BASELINE.json's north star
describes: every rank owns the row block of pq.py:218-225, runs the assignment
+ per-centroid sum/count kernels on it, and ONE all-reduce(SUM) of the fused
[K*d + K] fp32 buffer per iteration (precedent: pq.py:396-397) gives all ranks
identical new centroids — no broadcast needed afterwards.

The compute backend is `mevi_b200._lib.Context` (CUDA kernels via the C ABI);
the trainer itself only sequences kernels and collectives.
"""
from __future__ import annotations

import math
import os
import time
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .dist_utils import dist_on, gather_rows_to_rank0, rank_world, shard_bounds


FUSED_LLOYD = os.environ.get("MEVI_KMEANS_FUSED", "1") != "0"  # one-pass Lloyd iterations (mevi_kmeans_step_fused)
# how a Lloyd iteration after the first gets its sums: "delta" (default) = assignment pass + correction of float64 running
# sums by the rows whose assignment changed (mevi_kmeans_step_delta; no host synchronisation); "fused" = the assignment
# pass also sums every row under the previous assignment, changed rows moved afterwards (mevi_kmeans_step_fused);
# "twopass" = assignment pass + accumulation pass (mevi_kmeans_step)
LLOYD_ITERATION = os.environ.get("MEVI_KMEANS_ITER", "delta" if FUSED_LLOYD else "twopass")


TRACE = os.environ.get("MEVI_TRAIN_TRACE", "0") != "0"  # phase times (with device synchronisation) into last_info["trace"]
_trace = []


def _mark(name: str):
    if TRACE:
        torch.cuda.synchronize()
        _trace.append((name, time.perf_counter()))


def kmeanspp_init(sample, K: int, rs: np.random.RandomState) -> torch.Tensor:
    """k-means++ seeding (D^2 sampling) on a small sample, float64 arithmetic, on the device the sample lives on
    (32 passes over a 16,384 x 768 sample are seconds of numpy but milliseconds of device time).  The reference
    seeds with sklearn's init='k-means++' (pq.py:559); this is the textbook algorithm, with every random number
    drawn from `rs` (`random_state=seed` like the reference) so the seeds depend only on (seed, sample)."""
    x = torch.as_tensor(sample).to(torch.float64)
    n = x.shape[0]
    centers = torch.empty((K, x.shape[1]), dtype=torch.float64, device=x.device)
    first = int(rs.randint(n))
    u = rs.random_sample(K)  # one uniform per further centre, drawn up front
    centers[0] = x[first]
    d2 = ((x - centers[0]) ** 2).sum(1)
    for k in range(1, K):
        cum = torch.cumsum(d2, 0)
        tot = float(cum[-1].item())
        if not np.isfinite(tot) or tot <= 0:
            idx = int(u[k] * n) % n
        else:
            idx = int(torch.searchsorted(cum, torch.tensor(u[k] * tot, dtype=torch.float64, device=x.device)).item())
            idx = min(idx, n - 1)
        centers[k] = x[idx]
        d2 = torch.minimum(d2, ((x - centers[k]) ** 2).sum(1))
    return centers.to(torch.float32)


def _upload_rows(doc_emb, start: int, end: int, device: torch.device, chunk: int = 1 << 19) -> torch.Tensor:
    if isinstance(doc_emb, torch.Tensor):
        return doc_emb[start:end].to(device=device, dtype=torch.float32).clone()
    n, d = end - start, doc_emb.shape[1]
    out = torch.empty((n, d), dtype=torch.float32, device=device)
    for a in range(0, n, chunk):
        b = min(a + chunk, n)
        host = torch.from_numpy(np.ascontiguousarray(doc_emb[start + a : start + b], dtype=np.float32))
        out[a:b].copy_(host, non_blocking=False)
    return out


def _init_sample(R, init_sample: int, rs: np.random.RandomState, dev):
    """Rows for the k-means++ seeding, drawn from EVERY shard (init_sample // world rows per rank, seeded per rank) and
    gathered to rank 0 — a corpus stored in some order must not be seeded from its first N/world rows only.
    Returns a float32 tensor (on the compute device) on rank 0, None elsewhere.  Every rank draws from its own
    RandomState stream so the sample is reproducible for a given (seed, world)."""
    rank, world = rank_world()
    n = R.shape[0]
    per = max(1, init_sample // world)
    s = min(per, n)
    rs_r = np.random.RandomState(rs.randint(1 << 30) + 7919 * rank) if world > 1 else rs
    if s >= n:
        idx = np.arange(n)
    elif 4 * s >= n:
        idx = np.sort(rs_r.choice(n, size=s, replace=False))
    else:
        # numpy's choice(replace=False) permutes all n row numbers (155 ms per level at n = 8.8 M, more than the 25 Lloyd
        # iterations): for s << n draw with replacement, drop the repeats, and keep s of the distinct rows
        picked = np.unique(rs_r.randint(0, n, size=s + s // 4 + 16))
        while picked.size < s:
            picked = np.unique(np.concatenate([picked, rs_r.randint(0, n, size=s)]))
        idx = picked[np.sort(rs_r.permutation(picked.size)[:s])]
    mine = R[torch.from_numpy(idx).to(dev)].contiguous()
    if world == 1:
        return mine
    counts = torch.tensor([s], dtype=torch.int64, device=dev)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    g = gather_rows_to_rank0(mine, [int(c.item()) for c in all_counts])
    return g


def _reseed_empty(R, C, buf, K, rs: np.random.RandomState, dev):
    """Empty clusters (count 0 after the all-reduce, so every rank sees the same set) are re-seeded from rows of rank
    0's shard chosen with the trainer's RandomState, then broadcast: sklearn relocates empty / low-count centres too
    (MiniBatchKMeans reassignment_ratio, pq.py:559-560); leaving them in place wastes codes for good."""
    w = C.shape[1]
    empty = torch.nonzero(buf[K * w :] <= 0).squeeze(1)
    if empty.numel() == 0:
        return 0
    rank, _ = rank_world()
    if rank == 0:
        n = R.shape[0]
        pick = rs.choice(n, size=int(empty.numel()), replace=n < int(empty.numel()))
        C[empty] = R[torch.from_numpy(np.asarray(pick, dtype=np.int64)).to(dev)]
    if dist_on():
        dist.broadcast(C, 0)
    return int(empty.numel())


def _lloyd_level(be, R, K, col, stride, rs, iters, tol, init_sample, mode, inertia, buf, n_empty, dev, check_every=5, scratch=None):
    """One k-means problem on the local rows R [n, w] (all ranks in lockstep): k-means++ seeds from a sample of all
    shards, `iters` Lloyd iterations with ONE all-reduce of the fused sums|counts buffer each, then the labels under
    the FINAL centroids written to `col` (stride `stride`), like sklearn's fit_predict.  Convergence (relative inertia
    decrease <= tol) and empty clusters are looked at every `check_every` iterations only: that is the one point where
    the host waits for the device (an all-reduce of the inertia scalar + `.item()`).
    Returns (centroids [K, w] on the device, iterations run); `inertia` holds the global final inertia."""
    rank, _ = rank_world()
    n, w = R.shape
    C = torch.empty((K, w), dtype=torch.float32, device=dev)
    _mark("level_start")
    sample = _init_sample(R, init_sample, rs, dev)
    _mark("init_sample")
    if rank == 0:
        C.copy_(kmeanspp_init(sample, K, rs))
    del sample
    if dist_on():
        dist.broadcast(C, 0)
    _mark("kmeanspp")
    prev = math.inf
    n_it = 0
    # One pass per iteration where the library offers it (mevi_kmeans_step_fused): the pass that assigns the rows to the
    # current centroids also sums them under the PREVIOUS iteration's assignment; the rows whose assignment changed
    # (few after the first iterations) are then moved between their old and new centroid's sums.  The first iteration
    # has no previous assignment and takes the two-pass step.
    delta = LLOYD_ITERATION == "delta" and hasattr(be, "kmeans_step_delta")
    fused = not delta and LLOYD_ITERATION == "fused" and hasattr(be, "kmeans_step_fused") and n >= 4096
    a_prev = a_cur = moved_buf = master = n_changed = None
    if fused or delta:
        # buffers of the one-pass iterations, allocated once per training (not per iteration: the changed-row count differs
        # every time and a fresh multi-GB allocation per iteration costs more than the pass itself)
        scratch = scratch if scratch is not None else {}
        if scratch.get("n") != (n, w, delta):
            scratch.clear()
            scratch.update(n=(n, w, delta), a=torch.empty(n, dtype=torch.int32, device=dev), b=torch.empty(n, dtype=torch.int32, device=dev))
            if fused:
                scratch.update(moved=torch.empty((n // 4 + 1, w), dtype=torch.float32, device=dev))
        a_prev, a_cur, moved_buf = scratch["a"], scratch["b"], scratch.get("moved")
    if delta:
        master = torch.empty(K * w + K, dtype=torch.float64, device=dev)
        n_changed = torch.zeros(1, dtype=torch.int32, device=dev)
        changed_total = torch.zeros(1, dtype=torch.int64, device=dev)
    have_prev = False
    stats = {"delta_iters": 0, "fused_iters": 0, "two_pass_iters": 0, "changed_rows": 0}
    t_loop = time.perf_counter()
    for it in range(iters):
        if delta and have_prev:
            try:
                be.kmeans_step_delta(R, C, a_prev, a_cur, master, buf, n_changed=n_changed, inertia=inertia, mode=mode)
                changed_total += n_changed
                stats["delta_iters"] += 1
            except _lib.MeviError as e:
                if "unsupported" not in str(e):
                    raise
                delta = False
        if fused and have_prev:
            try:
                be.kmeans_step_fused(R, C, a_prev, a_cur, buf, inertia=inertia)
            except _lib.MeviError as e:
                if "unsupported" not in str(e):
                    raise
                fused = False
        if delta and have_prev:
            pass
        elif fused and have_prev:
            changed = torch.nonzero(a_cur != a_prev).squeeze(1)
            nc = int(changed.numel())
            stats["changed_rows"] += nc
            if nc > n // 4:  # cheaper to sum everything again than to move a quarter of the rows
                be.accumulate_by_code(R, a_cur, K, buf)
            elif nc > 0:
                moved = be.gather_rows(R, changed.to(torch.int32), out=moved_buf)
                plus = be.accumulate_by_code(moved, a_cur[changed].contiguous(), K)
                minus = be.accumulate_by_code(moved, a_prev[changed].contiguous(), K)
                buf.add_(plus).sub_(minus)
            stats["fused_iters"] += 1
        else:
            one_pass = fused or delta
            be.kmeans_step(R, C, buf, assign=(a_cur if one_pass else col), assign_stride=(1 if one_pass else stride), inertia=inertia, mode=mode)
            if delta:
                master.copy_(buf)  # fp32 -> float64: the running sums the later iterations correct
            stats["two_pass_iters"] += 1
        if fused or delta:
            a_prev, a_cur = a_cur, a_prev
            have_prev = True
        if dist_on():
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        be.kmeans_update(buf, C, n_empty)
        n_it = it + 1
        if n_it % check_every == 0 or n_it == iters:
            if dist_on():
                dist.all_reduce(inertia, op=dist.ReduceOp.SUM)
            cur = float(inertia.item())  # the host waits for the device here (and only here)
            reseeded = _reseed_empty(R, C, buf, K, rs, dev) if int(n_empty.item()) > 0 and n_it < iters else 0
            if not reseeded and tol is not None and prev - cur <= tol * max(cur, 1e-30):
                break
            prev = cur
    if delta and stats["delta_iters"]:
        stats["changed_rows"] = int(changed_total.item())
    _lloyd_level.last_stats = stats
    _lloyd_level.last_loop_seconds = time.perf_counter() - t_loop  # ends on the .item() of the last check: device time
    _mark("iterations")
    be.kmeans_step(R, C, buf, assign=col, assign_stride=stride, inertia=inertia, mode=mode)
    if dist_on():
        dist.all_reduce(inertia, op=dist.ReduceOp.SUM)
    if hasattr(be, "check"):
        be.check()
    _mark("final_labels")
    return C, n_it


@torch.no_grad()
def train_rq_lloyd(doc_emb, M: int, K: int, seed: int, iters: int = 25, tol: Optional[float] = 1e-7, init_sample: int = 16384,
                   mode: str = "auto", device_index: Optional[int] = None, backend=None, metric: str = "l2",
                   gather_codes: bool = True, presharded: bool = False):
    """Returns (codebook [M,K,d] fp32 tensor on the compute device, codes np.int32 [N,M] on rank 0 or None).
    k-means is always L2 (as sklearn's in the reference), whatever `metric` the encoder uses later.
    `presharded=True`: `doc_emb` already is THIS rank's row block (a device tensor the caller sharded with the
    pq.py:218-225 rule) instead of the whole corpus; codes are then returned as the local device tensor.
    `tol=None` runs exactly `iters` iterations per level."""
    be = backend if backend is not None else _lib.get_context(device_index)
    dev = be.torch_device if backend is not None else torch.device("cuda", be.device)
    rank, world = rank_world()
    if presharded:
        n, d = doc_emb.shape
        cnt = torch.tensor([n], dtype=torch.int64, device=dev)
        if dist_on():
            dist.all_reduce(cnt)
        N, start, end = int(cnt.item()), 0, n
    else:
        N, d = doc_emb.shape
        start, end = shard_bounds(N, rank, world)
        n = end - start
    del _trace[:]
    _mark("start")
    R = _upload_rows(doc_emb, start, end, dev)
    _mark("working_copy")
    codes = torch.zeros((n, M), dtype=torch.int32, device=dev)
    codebook = torch.empty((M, K, d), dtype=torch.float32, device=dev)
    buf = torch.empty(K * d + K, dtype=torch.float32, device=dev)
    inertia = torch.zeros(1, dtype=torch.float64, device=dev)
    n_empty = torch.zeros(1, dtype=torch.int32, device=dev)
    rs = np.random.RandomState(seed)
    info = {"levels": [], "world": world, "rows_local": n}
    scratch = {}
    for j in range(M):
        t0 = time.time()
        col = codes[:, j]
        C, n_it = _lloyd_level(be, R, K, col, M, rs, iters, tol, init_sample, mode, inertia, buf, n_empty, dev, scratch=scratch)
        codebook[j].copy_(C)
        if j != M - 1:  # pq.py:591-593
            be.residual_update(R, C, col, assign_stride=M)
        _mark("residual")
        info["levels"].append({"level": j, "iters": n_it, "inertia": float(inertia.item()),
                               "mse": float(inertia.item()) / max(N, 1) / d, "seconds": time.time() - t0,
                               "loop_ms_per_iter": getattr(_lloyd_level, "last_loop_seconds", 0.0) / max(n_it, 1) * 1e3,
                               **getattr(_lloyd_level, "last_stats", {})})
    if TRACE:
        info["trace"] = [(b[0], round((b[1] - a[1]) * 1e3, 2)) for a, b in zip(_trace, _trace[1:])]
    train_rq_lloyd.last_info = info
    if presharded:
        return codebook, codes
    codes_all = None
    if gather_codes:
        counts = [shard_bounds(N, r, world)[1] - shard_bounds(N, r, world)[0] for r in range(world)]
        g = gather_rows_to_rank0(codes, counts)
        if g is not None:
            codes_all = g.cpu().numpy()
    return codebook, codes_all


@torch.no_grad()
def train_pq_lloyd(doc_emb, M: int, K: int, seed: int, iters: int = 25, tol: float = 1e-7, init_sample: int = 16384,
                   mode: str = "auto", device_index: Optional[int] = None, backend=None, gather_codes: bool = True):
    """pq branch of MEVI/pq.py:568-581: one independent k-means per sub-vector slice
    doc_emb[:, j*dsub:(j+1)*dsub].  Same sharding and collectives as train_rq_lloyd.
    Returns (codebook [M,K,d/M] on the compute device, codes np.int32 [N,M] on rank 0 or None)."""
    be = backend if backend is not None else _lib.get_context(device_index)
    dev = be.torch_device if backend is not None else torch.device("cuda", be.device)
    rank, world = rank_world()
    N, d = doc_emb.shape
    assert d % M == 0
    dsub = d // M
    start, end = shard_bounds(N, rank, world)
    n = end - start
    X = _upload_rows(doc_emb, start, end, dev)
    codes = torch.zeros((n, M), dtype=torch.int32, device=dev)
    codebook = torch.empty((M, K, dsub), dtype=torch.float32, device=dev)
    buf = torch.empty(K * dsub + K, dtype=torch.float32, device=dev)
    inertia = torch.zeros(1, dtype=torch.float64, device=dev)
    n_empty = torch.zeros(1, dtype=torch.int32, device=dev)
    rs = np.random.RandomState(seed)
    info = {"levels": [], "world": world, "rows_local": n}
    for j in range(M):
        t0 = time.time()
        R = X[:, j * dsub : (j + 1) * dsub].contiguous()
        C, n_it = _lloyd_level(be, R, K, codes[:, j], M, rs, iters, tol, init_sample, mode, inertia, buf, n_empty, dev)
        codebook[j].copy_(C)
        info["levels"].append({"level": j, "iters": n_it, "inertia": float(inertia.item()),
                               "mse": float(inertia.item()) / max(N, 1) / dsub, "seconds": time.time() - t0})
        del R
    train_pq_lloyd.last_info = info
    codes_all = None
    if gather_codes:
        counts = [shard_bounds(N, r, world)[1] - shard_bounds(N, r, world)[0] for r in range(world)]
        g = gather_rows_to_rank0(codes, counts)
        if g is not None:
            codes_all = g.cpu().numpy()
    return codebook, codes_all
