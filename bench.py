#!/usr/bin/env python
"""bench.py — headline benchmark of the MEVI index hot path on B200.

Metric (BASELINE.json): docs/sec of the RQ encode (M=4 levels x K=32 centroids, d=768, L2) on the
MSMARCO-shape corpus, 8,841,823 x 768 fp32 per GPU, resident in HBM when the timed region starts.
A "step" is one full encode pass over that corpus.  Documents shard across GPUs with no data-path
collective (weak scaling: every rank owns one MSMARCO-shape block).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # reference CPU path

The JSON line also carries `roofline` (dominant kernel vs measured HBM peak), `e2e` (same metric
through the reference-shaped Python entry point with HOST buffers, copies inside the timed region),
`cpu_baseline` (the reference's CPU arithmetic, oracle port, bounded sample), `clocks`, `gpu_launches`
and `extra` — the other BASELINE configs, each with its own roofline: `rerank` (streaming kernel) and
`rerank_grouped` (leaf-grouped tensor GEMMs, with parity against the streaming kernel), `kmeans_iteration`
(whole iteration + per-kernel), `flat_ip` (with parity against the fp32 kernel), `widened_rows` (PQ encode,
device beam search, inverted lists, leaf-order permutation).  At N > 1 the re-rank and flat extras are
doc-sharded: all-gather of per-shard top-k + merge inside their timed regions.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_MARCO, D, M_LEVELS, K_CENTS = 8841823, 768, 4, 32
NQ_MARCO, TOPK, LEAVES = 6980, 100, 100
GOLDEN_CB = os.path.join(ROOT, "tests", "golden", "gauss768", "codebook.pt")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), float(j.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.samples = []
        self.proc = None
        self.gpu = gpu_index
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "25", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.samples:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                clk, cmx = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            mx = cmx
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                try:
                    power.append(float(parts[2]))
                except ValueError:
                    pass
                for nm, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def make_corpus(n, d, device, seed):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    X = torch.empty((n, d), dtype=torch.float32, device=device)
    step = 1 << 20
    for a in range(0, n, step):
        b = min(a + step, n)
        X[a:b].normal_(generator=g)
    return X


def load_codebook():
    import torch

    return torch.load(GOLDEN_CB, map_location="cpu", weights_only=False).detach().contiguous()


# --------------------------------------------------------------------------- #
def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path (torch CPU ops of pq.py:281-305, batch 128
    as main_models.py:3208) on the host cores — the oracle port, since the reference is Python and
    /root/reference does not exist on the GPU box."""
    if rank != 0:
        return
    import numpy as np
    import torch

    from oracle import oracle

    torch.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1; the reference uses every core
    torch.manual_seed(1234)
    S = args.ref_sample
    X = torch.randn(S, D).numpy()
    cb = load_codebook()
    for _ in range(max(args.warmup, 1)):
        oracle.rq_encode(X[: min(S, 8192)], cb, batch_size=128)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        codes = oracle.rq_encode(X, cb, batch_size=128)
    dt = time.perf_counter() - t0
    value = S * args.steps / dt
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "rq_encode_docs_per_sec", "value": value, "unit": "docs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: MSMARCO-shape RQ encode, M=4 K=32 d=768 L2 (reference CPU path, pq.py:281-305, batch 128)",
                   "sample_rows_per_step": S, "host_threads": cores},
        "cpu_baseline": {"value": value, "unit": "docs/s", "cores": cores, "kind": "port",
                         "sample": f"{S} x {D} fp32 rows per step, N(0,1), reference-trained codebook"},
        "e2e": {"value": value, "unit": "docs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "codes_checksum": int(np.asarray(codes, dtype=np.int64).sum()),
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- #
def our_arm(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist

    import mevi_b200
    from mevi_b200.pq import ProductQuantization

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = mevi_b200.get_context(local_rank)
    hbm_peak, bf16_peak, peak_src = measured_peaks()
    n = args.docs
    log(f"[rank {rank}] generating {n} x {D} fp32 corpus on {torch.cuda.get_device_name(dev)}")
    X = make_corpus(n, D, dev, 1234 + rank)
    cb_cpu = load_codebook()
    cb = cb_cpu.to(dev)
    codes = torch.empty((n, M_LEVELS), dtype=torch.int32, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        ctx.rq_encode(X, cb, metric="l2", mode=args.mode, codes=codes)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launches
    barrier()
    t_wall0 = time.time()
    start.record()
    for _ in range(args.steps):
        step()
    stop.record()
    barrier()
    t_wall1 = time.time()
    launches = ctx.launches - l0
    ms = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * n / (ms_step / 1e3)
    _, stats = ctx.rq_encode(X[: min(n, 1 << 20)], cb, mode=args.mode, return_stats=True)
    flagged_frac = float(stats[0].item()) / max(1, int(stats[1].item()))
    checksum = int(codes.to(torch.int64).sum().item())

    # kernel-vs-kernel agreement of the fast path with the exact path on a sample (both product kernels)
    ns = min(n, 1 << 18)
    exact = ctx.rq_encode(X[:ns], cb, mode="exact")
    mismatch = int((exact != codes[:ns]).any(dim=1).sum().item())

    algo_bytes = n * D * 4 + n * M_LEVELS * 4
    achieved = algo_bytes / (ms_step / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        try:
            traffic = json.load(open(tpath)).get("rq_encode_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": "rq_encode (all 4 levels, one pass over the corpus)", "algorithmic_bytes_per_launch": algo_bytes}

    # ---- e2e through the reference-shaped entry point with host buffers ----------------------
    e2e = None
    try:
        import psutil

        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    n_e2e = n
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))  # every rank of the node pins its own host copy
    while n_e2e * D * 4 * 1.3 > avail * 0.6 / max(1, local_world) and n_e2e > 1 << 18:
        n_e2e //= 2
    pq = ProductQuantization("rq", M_LEVELS, 5, "l2", D, "kmeans", "grad")
    pq.kernel_mode = args.mode
    pq.device_index = local_rank
    with torch.no_grad():
        pq.codebook.copy_(cb_cpu)
    Xh = torch.empty((n_e2e, D), dtype=torch.float32, pin_memory=True)
    Xh.copy_(X[:n_e2e])
    cluster = torch.empty((n_e2e, M_LEVELS), dtype=torch.int32).pin_memory()
    Xh_np = Xh.numpy()
    pq.get_rq_document_cluster(Xh_np[: 1 << 16], cluster[: 1 << 16], 0, min(n_e2e, 1 << 16), rank, 128)  # warm-up
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        pq.get_rq_document_cluster(Xh_np, cluster, 0, n_e2e, rank, 128)
    e1.record()
    barrier()
    e_ms = torch.tensor([e0.elapsed_time(e1) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
    e2e_ok = bool((cluster.to(dev) == codes[:n_e2e]).all().item())
    e2e = {"value": world * n_e2e / (float(e_ms.item()) / 1e3), "unit": "docs/s",
           "h2d_bytes_per_step": n_e2e * D * 4 + int(cb_cpu.numel()) * 4, "d2h_bytes_per_step": n_e2e * M_LEVELS * 4,
           "rows_per_step": n_e2e, "ms_per_step": float(e_ms.item()), "steps": e2e_steps,
           "api": "ProductQuantization.get_rq_document_cluster(np.ndarray) -> mevi_rq_encode_host (pinned host rows, 256k-row chunks)",
           "codes_equal_device_path": e2e_ok}
    del Xh, Xh_np, cluster

    extra = {}
    if not args.no_extras:
        extra = run_extras(args, ctx, X, cb, codes, dev, rank, world, hbm_peak, bf16_peak)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_leg(X, cb_cpu)

    if rank == 0:
        line = {
            "metric": "rq_encode_docs_per_sec", "value": value, "unit": "docs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: MSMARCO-shape RQ encode, 8,841,823 x 768 fp32 per GPU, M=4 K=32 L2, "
                                   "codebook trained by the reference (sklearn, seed 41) on N(0,1) data",
                       "docs_per_gpu": n, "parallelism": f"doc-sharded x{world}, no data-path collective",
                       "kernel_mode": args.mode, "l2_policy": "inputs (27.2 GB per GPU) are larger than L2 (126 MB)",
                       "timing": "CUDA events on the launch stream, max over ranks"},
            "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu_baseline, "clocks": clocks, "gpu_launches": launches,
            "codes_checksum": checksum, "prefilter_flagged_fraction": flagged_frac,
            "fast_vs_exact_kernel_mismatch_rows": {"rows": ns, "mismatch": mismatch}, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline_leg(X, cb_cpu):
    """Reference CPU arithmetic (oracle port of pq.py:281-305, batch 128) on a bounded sample of the
    same corpus (~10-30 s of CPU work).  Runs in a CHILD process that never initialises CUDA: inside a
    CUDA process every munmap of the 12.6 MB torch temporaries goes through the UVM notifier and the
    CPU path runs ~10x slower than the reference would on its own."""
    import tempfile

    import numpy as np

    S = int(min(1 << 18, X.shape[0]))
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(shm, f"mevi_bench_sample_{os.getpid()}.npy")
    np.save(path, X[:S].cpu().numpy())
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "cpu-baseline-child", "--sample-file", path],
                             capture_output=True, text=True, timeout=600, env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
        line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
        return json.loads(line)
    except Exception as e:
        return {"error": repr(e)[:300]}
    finally:
        try:
            os.remove(path)
        except OSError:
            pass


def cpu_baseline_child(args):
    import numpy as np
    import torch

    from oracle import oracle

    torch.set_num_threads(os.cpu_count() or 1)
    sample = np.load(args.sample_file)
    cb_cpu = load_codebook()
    oracle.rq_encode(sample[:8192], cb_cpu, batch_size=128)  # warm-up (allocator, threads)
    t0 = time.perf_counter()
    oracle.rq_encode(sample[:32768], cb_cpu, batch_size=128)
    rate = 32768 / (time.perf_counter() - t0)
    reps = int(max(1, min(64, round(rate * 15.0 / len(sample)))))
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle.rq_encode(sample, cb_cpu, batch_size=128)
    dt = time.perf_counter() - t0
    print(json.dumps({"value": len(sample) * reps / dt, "unit": "docs/s", "cores": torch.get_num_threads(), "kind": "port",
                      "sample": f"first {len(sample)} rows of the bench corpus x {reps} passes, torch CPU restatement of "
                                f"pq.py:281-305 (batch 128) in a CUDA-free child process, {dt:.1f} s",
                      "host_cpus": os.cpu_count()}), flush=True)


def run_extras(args, ctx, X, cb, codes, dev, rank, world, hbm_peak, bf16_peak):
    """The other BASELINE.json configs, each one short: (iv) cluster-restricted re-rank of 6,980
    queries x 100 leaves -> top-100; (iii) one k-means iteration with the sums|counts all-reduce;
    (v) exact flat inner-product top-100 on a bounded shard."""
    import torch
    import torch.distributed as dist

    from mevi_b200.pq import ProductQuantization
    from mevi_b200.rerank import ClusterIndex

    n = X.shape[0]
    out = {}

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- (iv) re-rank ---------------------------------------------------------------------
    try:
        g = torch.Generator(device=dev)
        g.manual_seed(4321)
        Q = torch.empty((NQ_MARCO, D), device=dev).normal_(generator=g)
        pq = ProductQuantization("rq", M_LEVELS, 5, "l2", D, "kmeans", "grad")
        with torch.no_grad():
            pq.codebook.copy_(cb.cpu())
        dec = torch.cat([pq.beam_search(Q[a : a + 128], LEAVES) for a in range(0, NQ_MARCO, 128)])
        t0 = time.perf_counter()
        index = ClusterIndex.from_codes(codes, K_CENTS, id_base=0, device_index=dev.index)
        torch.cuda.synchronize()
        t_index = time.perf_counter() - t0
        ql = index.lookup(dec)
        t0 = time.perf_counter()
        D_leaf = ctx.gather_rows(X, index.leaf_docids)  # one-time permutation into CSR (leaf) order
        torch.cuda.synchronize()
        t_perm = time.perf_counter() - t0
        res = {}

        from mevi_b200.dist_utils import all_gather_stack

        def rr():
            # documents are sharded: every rank scores the candidates it owns for ALL queries, then the per-shard
            # top-k lists are all-gathered and merged on every rank (the same sequence as ClusterReranker.rerank)
            sc, ids, nc = ctx.cluster_rerank(Q, D_leaf, index.leaf_offsets, index.leaf_docids, ql, TOPK, id_base=rank * n,
                                             leaf_ordered=True)
            if world > 1:
                sc, ids = ctx.topk_merge(all_gather_stack(sc).contiguous(), all_gather_stack(ids).contiguous())
            res["out"] = (sc, ids, nc)

        ms = timed(rr, 2)
        ncand = res["out"][2].to(torch.float64)
        gathered = float(ncand.sum().item()) * D * 4
        out["rerank"] = {
            "metric": "rerank_queries_per_sec", "value": NQ_MARCO / (ms / 1e3), "unit": "queries/s",
            "corpus_docs": world * n, "sharding": f"documents row-sharded over {world} GPU(s); all-gather of [nq,k] (score,id) + merge "
                                                  "inside the timed region" if world > 1 else "single GPU, no collective",
            "candidates_scored_per_sec": world * float(ncand.sum().item()) / (ms / 1e3),
            "ms_per_step": ms, "queries": NQ_MARCO, "leaves_per_query": LEAVES, "topk": TOPK,
            "candidates_mean": float(ncand.mean().item()), "candidates_max": float(ncand.max().item()),
            "empty_leaf_fraction": float((ql < 0).float().mean().item()), "n_leaves": index.n_leaves,
            "index_build_s": t_index, "leaf_order_permutation_s": t_perm, "layout": "documents stored in CSR (leaf) order; leaves streamed with bulk async copies",
            "roofline": {"bound": "hbm", "achieved": gathered / (ms / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": gathered / (ms / 1e3) / 1e9 / hbm_peak, "traffic": None,
                         "note": "bytes = sum_q candidates_q * 4*d, no cross-query reuse assumed"},
        }
        # ---- same workload, leaf-grouped GEMM formulation (K3g): every leaf read once for all the queries that chose it
        try:
            from mevi_b200.rerank import ClusterReranker

            t0 = time.perf_counter()
            rrg = ClusterReranker(None, index, mode="grouped", D_leaf=D_leaf)
            torch.cuda.synchronize()
            t_img = time.perf_counter() - t0
            gres = {}

            def rg():
                gres["out"] = rrg.rerank(Q, dec, topk=TOPK)  # includes the all-gather + merge when torch.distributed is up

            ms_g = timed(rg, 3)
            sg, ig, _ = gres["out"]
            ss, is_, _ = res["out"]
            fin = torch.isfinite(sg) & torch.isfinite(ss)
            out["rerank_grouped"] = {
                "metric": "rerank_queries_per_sec", "value": NQ_MARCO / (ms_g / 1e3), "unit": "queries/s", "ms_per_step": ms_g,
                "path": rrg.last_path, "corpus_docs": world * n, "tile_image_build_s": t_img,
                "speedup_vs_streaming_kernel": ms / ms_g,
                "formulation": "per leaf a [docs of the leaf] x [queries that chose it] fp16 tcgen05 GEMM (prefilter with a rigorous "
                               "margin) + exact fp32 re-score; thresholds bootstrapped from the exact top-k of a 2,048-row prefix; "
                               "falls back to the streaming kernel when the guarantee cannot be established",
                "parity_vs_streaming_kernel": None if world > 1 else {
                    "ids_identical_fraction": float((ig == is_).float().mean().item()),
                    "max_abs_score_diff": float((sg - ss)[fin].abs().max().item()) if bool(fin.any()) else 0.0,
                    "note": "ids may differ only where two documents tie in fp32 score"},
                "roofline": {"bound": "hbm", "achieved": gathered / (ms_g / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": gathered / (ms_g / 1e3) / 1e9 / hbm_peak,
                             "note": "same no-reuse byte count as the streaming kernel (sum_q candidates_q * 4*d) as the numerator: "
                                     "> 1 because a leaf's rows are read once for ~30 queries (SURVEY 8d, crossover note)"}}
            del rrg
        except Exception as e:
            out["rerank_grouped"] = {"error": repr(e)[:300]}
        del index, ql, dec, D_leaf
    except Exception as e:  # extras must never kill the headline line
        out["rerank"] = {"error": repr(e)[:300]}

    # ---- (iii) k-means iteration ----------------------------------------------------------
    try:
        C = cb[0].clone()
        buf = torch.empty(K_CENTS * D + K_CENTS, device=dev)
        assign = torch.empty(n, dtype=torch.int32, device=dev)

        def km():
            ctx.kmeans_step(X, C, buf, assign=assign, mode=args.mode)
            if world > 1:
                dist.all_reduce(buf)
            ctx.kmeans_update(buf, C)

        ms = timed(km, 3)
        bytes_ = n * D * 4
        # the two kernels of an iteration, each against the HBM roofline of ITS pass over the shard
        a2 = assign.view(n, 1)
        ms_assign = timed(lambda: ctx.rq_encode(X, C[None], metric="l2", mode=args.mode, codes=a2), 5)
        ms_accum = timed(lambda: ctx.accumulate_by_code(X, assign, K_CENTS, buf), 5)

        def roof(ms_k, nbytes, note):
            return {"bound": "hbm", "achieved": nbytes / (ms_k / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": nbytes / (ms_k / 1e3) / 1e9 / hbm_peak, "ms": ms_k, "note": note}

        out["kmeans_iteration"] = {
            "ms": ms, "docs_per_sec": world * n / (ms / 1e3), "allreduce_bytes": (K_CENTS * D + K_CENTS) * 4,
            "roofline": roof(ms, bytes_, "whole iteration against ONE pass over the shard (4*d bytes per doc); the iteration "
                                         "makes two passes (assign, then accumulate), see DESIGN.md for why they are not fused"),
            "kernels": {"assign (rq_encode, M=1)": roof(ms_assign, bytes_ + n * 4, "reads the shard once, writes int32 assignments"),
                        "accumulate_by_code": roof(ms_accum, bytes_ + n * 4, "reads the shard and the assignments once; timed as 5 "
                                                   "back-to-back launches, which runs slower than the same launch inside the "
                                                   "iteration (iteration ms - assign ms is the in-step cost)")},
            "accumulate_in_step_ms_estimate": ms - ms_assign}
        del assign, a2
    except Exception as e:
        out["kmeans_iteration"] = {"error": repr(e)[:300]}

    # ---- (v) flat IP on a bounded shard ----------------------------------------------------
    try:
        shard = min(n, args.flat_docs)
        g = torch.Generator(device=dev)
        g.manual_seed(4321)
        Q = torch.empty((NQ_MARCO, D), device=dev).normal_(generator=g)

        from mevi_b200.dist_utils import all_gather_stack

        def fl():
            sc, ids = ctx.flat_ip_topk(Q, X[:shard], TOPK, id_base=rank * shard, mode=args.mode)
            if world > 1:  # docs sharded: all-gather of per-shard top-k + merge (faiss_search.search under torch.distributed)
                ctx.topk_merge(all_gather_stack(sc).contiguous(), all_gather_stack(ids).contiguous())

        ms = timed(fl, 2)
        # parity of the tensor path with the direct fp32 search (both product kernels) on a slice of the queries
        nchk = 256
        s_t, i_t = ctx.flat_ip_topk(Q[:nchk], X[:shard], TOPK, mode=args.mode)
        s_e, i_e = ctx.flat_ip_topk(Q[:nchk], X[:shard], TOPK, mode="exact")
        parity = {"queries": nchk, "ids_identical_fraction": float((i_t == i_e).float().mean().item()),
                  "max_rel_score_diff": float(((s_t - s_e).abs() / s_e.abs().clamp_min(1e-6)).max().item()),
                  "note": "tensor-prefilter path vs fp32 CUDA-core path; ids may differ only at fp32 score ties"}
        flops = 2.0 * NQ_MARCO * shard * D
        ach = flops / (ms / 1e3) / 1e12
        out["flat_ip"] = {"ms": ms, "docs_per_gpu": shard, "corpus_docs": world * shard, "queries": NQ_MARCO, "topk": TOPK,
                          "queries_per_sec": NQ_MARCO / (ms / 1e3), "pairs_per_sec": world * NQ_MARCO * shard / (ms / 1e3),
                          "parity_vs_fp32_kernel": parity,
                          "roofline": {"bound": "tensor", "achieved": ach, "peak": bf16_peak, "unit": "TFLOP/s", "frac": ach / bf16_peak,
                                       "frac_of_tf32_equivalent_peak": ach / (bf16_peak / 2.0),
                                       "note": "FLOPs = 2*nq*N*d (algorithmic); peak = measured 16-bit cuBLAS burst rate (the kernel's MMAs "
                                               "are fp16 tcgen05, one pass); inputs and results are fp32, for which the dense rate would be "
                                               "half of that (TF32) - both fractions given; the time is the WHOLE call: fp32->fp16 image "
                                               "pass, GEMM with prefilter epilogue, compactions, exact fp32 re-score"}}
    except Exception as e:
        out["flat_ip"] = {"error": repr(e)[:300]}

    # ---- the callers / modes either side of the path, same measurement bar ------------------
    try:
        wid = {}
        hbm = lambda ms_k, nbytes: {"ms": ms_k, "GB/s": nbytes / (ms_k / 1e3) / 1e9, "frac_of_hbm_peak": nbytes / (ms_k / 1e3) / 1e9 / hbm_peak}
        # f4: pq/opq encode (pq.py:249-279): M sub-vectors x K centroids, one pass over the shard
        gq = torch.Generator(device=dev)
        gq.manual_seed(77)
        for Mp, Kp in ((4, 32), (24, 256)):
            cbp = torch.empty((Mp, Kp, D // Mp), device=dev).normal_(generator=gq)
            cp = torch.empty((n, Mp), dtype=torch.int32, device=dev)
            ms_pq = timed(lambda: ctx.pq_encode(X, cbp, metric="l2", codes=cp), 3)
            wid[f"pq_encode_M{Mp}_K{Kp}"] = dict(hbm(ms_pq, n * D * 4 + n * Mp * 4), docs_per_sec=world * n / (ms_pq / 1e3))
            del cbp, cp
        # f3: rq beam search on the device (pq.py:613-713), 100 beams per query, all queries in one call
        gq.manual_seed(4321)
        Qb = torch.empty((NQ_MARCO, D), device=dev).normal_(generator=gq)
        ms_beam = timed(lambda: ctx.rq_beam_search(Qb, cb, LEAVES, metric="l2", prod=True), 3)
        wid["rq_beam_search"] = {"ms": ms_beam, "queries_per_sec": NQ_MARCO / (ms_beam / 1e3), "beams": LEAVES,
                                 "note": "leaf producer of the re-rank (not in its timed region)"}
        # f1: inverted lists from codes (pq.py:236-242): sort by leaf key -> CSR + permutation
        ms_inv = timed(lambda: ctx.build_inverted_lists(codes, K_CENTS), 3)
        wid["build_inverted_lists"] = {"ms": ms_inv, "docs_per_sec": world * n / (ms_inv / 1e3)}
        # f1: one-time permutation of the corpus into leaf order (read + write of the shard)
        docids, _ = ctx.build_inverted_lists(codes, K_CENTS)
        ms_perm = timed(lambda: ctx.gather_rows(X, docids), 2)
        wid["leaf_order_permutation"] = hbm(ms_perm, 2 * n * D * 4)
        out["widened_rows"] = wid
    except Exception as e:
        out["widened_rows"] = {"error": repr(e)[:300]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference", "cpu-baseline-child"])
    ap.add_argument("--sample-file", type=str, default=None)
    ap.add_argument("--mode", type=str, default="auto", choices=["auto", "exact", "tensor"])
    ap.add_argument("--docs", type=int, default=N_MARCO, help="rows per GPU (default: MSMARCO 8,841,823)")
    ap.add_argument("--flat-docs", type=int, default=1 << 22)
    ap.add_argument("--ref-sample", type=int, default=65536)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if args.impl == "cpu-baseline-child":
        cpu_baseline_child(args)
        return
    if world != args.gpus:
        log(f"note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    our_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
