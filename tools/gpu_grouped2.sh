#!/bin/bash
# grouped re-rank at the bench shape: failed-query count and time per variant of the bootstrap / round plan
timeout 600 python - <<'PY'
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
from mevi_b200.pq import ProductQuantization
from mevi_b200.rerank import ClusterIndex, ClusterReranker, plan_grouped_rounds, plan_grouped_tile_rounds
ctx = mevi_b200.get_context(0)
dev = torch.device("cuda", 0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach().cuda()
n = 8841823
g = torch.Generator(device=dev); g.manual_seed(1234)
X = torch.empty((n, 768), device=dev)
for a in range(0, n, 1 << 20): X[a:a + (1 << 20)].normal_(generator=g)
codes = ctx.rq_encode(X, cb)
g.manual_seed(4321)
Q = torch.empty((6980, 768), device=dev).normal_(generator=g)
pq = ProductQuantization("rq", 4, 5, "l2", 768, "kmeans", "grad")
with torch.no_grad(): pq.codebook.copy_(cb.cpu())
dec = torch.cat([pq.beam_search(Q[a:a + 1024], 100) for a in range(0, 6980, 1024)])
index = ClusterIndex.from_codes(codes, 32)
D_leaf = ctx.gather_rows(X, index.leaf_docids)
del X
rr = ClusterReranker(None, index, mode="grouped", D_leaf=D_leaf)
ql = index.lookup(dec)
s_ref, i_ref, _ = ctx.cluster_rerank(Q, D_leaf, index.leaf_offsets, index.leaf_docids, ql, 100, leaf_ordered=True)
def run_tiles(bl):
    rr.PLAN, rr.BOOT_LEAVES = "tiles", bl
    for _ in range(2): out = rr.rerank(Q, dec, topk=100)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): out = rr.rerank(Q, dec, topk=100)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
    eq = float((out[1] == i_ref).float().mean().item())
    plan = plan_grouped_tile_rounds(rr._grouped["leaf_tile0"], ql, bl)
    print(f"tiles plan, boot leaves {str(bl):14s} {ms:7.2f} ms  path {rr.last_path:15s} failed {rr.last_failed_queries:5d} weak {rr.last_weak_queries:5d}  ids_eq {eq:.5f}  items/round {[int(p[0].numel()) for p in plan]} groups/round {[int(p[2].numel())//64 for p in plan]}", flush=True)
    gg = rr._grouped
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): plan = plan_grouped_tile_rounds(gg["leaf_tile0"], ql, bl)
    torch.cuda.synchronize(); print("   planning ms", round((time.perf_counter() - t0) / 5 * 1e3, 2), flush=True)
for bl in ((8, 63), (16, 63)):
    run_tiles(bl)
def run(tag, boot, rounds):
    rr.PLAN = "prefix"
    rr.BOOTSTRAP_ROWS, rr.ROUND_ROWS = boot, rounds
    for _ in range(2): out = rr.rerank(Q, dec, topk=100)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): out = rr.rerank(Q, dec, topk=100)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
    eq = float((out[1] == i_ref).float().mean().item())
    plan = plan_grouped_rounds(index.leaf_offsets, rr._grouped["leaf_tile0"], ql, rounds, boot)
    print(f"{tag:34s} {ms:7.2f} ms  path {rr.last_path:15s} failed {rr.last_failed_queries:5d} weak {rr.last_weak_queries:5d}  ids_eq {eq:.5f}  items/round {[int(p[0].numel()) for p in plan]} groups/round {[int(p[2].numel())//64 for p in plan]}", flush=True)
for boot, rounds in ((3072, (32768,)),):
    run(f"boot {boot} rounds {rounds}", boot, rounds)
# phase timing of the default plan
rr.BOOTSTRAP_ROWS, rr.ROUND_ROWS = 3072, (32768,)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): plan = plan_grouped_rounds(index.leaf_offsets, rr._grouped["leaf_tile0"], ql, rr.ROUND_ROWS, rr.BOOTSTRAP_ROWS)
torch.cuda.synchronize(); print("planning ms", (time.perf_counter() - t0) / 5 * 1e3)
gg = rr._grouped
for r, (it, ig, gq) in enumerate(plan):
    ctx.rerank_grouped_begin(Q, gg["absmax"], gg["maxnorm"], None)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): ctx.rerank_grouped_round(Q, gg["img"], gg["row0"], gg["nrows"], it, ig, gq, 100)
    torch.cuda.synchronize(); print(f"round {r}: {int(it.numel())} items, {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms (thresholds of a fresh call: upper bound)")
PY
