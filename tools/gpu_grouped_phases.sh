#!/bin/bash
# phase breakdown of one grouped re-rank call at the bench size
timeout 200 python /dev/stdin <<'PY'
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
from mevi_b200.pq import ProductQuantization
from mevi_b200.rerank import ClusterIndex, ClusterReranker, plan_grouped_rounds
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach()
n, nq, L, k = 8841823, 6980, 100, 100
g = torch.Generator(device="cuda"); g.manual_seed(1234)
X = torch.empty((n, 768), device="cuda")
for a in range(0, n, 1 << 20): X[a:a + (1 << 20)].normal_(generator=g)
codes = ctx.rq_encode(X, cb.cuda(), mode="auto")
g.manual_seed(4321)
Q = torch.empty((nq, 768), device="cuda").normal_(generator=g)
pq = ProductQuantization("rq", 4, 5, "l2", 768, "kmeans", "grad")
with torch.no_grad(): pq.codebook.copy_(cb)
dec = torch.cat([pq.beam_search(Q[a:a + 128], L) for a in range(0, nq, 128)])
index = ClusterIndex.from_codes(codes, 32)
rr = ClusterReranker(X, index, mode="grouped"); del X
gr = rr._grouped; off = index.leaf_offsets
for _ in range(2): rr.rerank(Q, dec, topk=k)
def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
torch.cuda.synchronize()
for rep in range(2):
    t = [ev()]
    ql = index.lookup(dec); t.append(ev())
    s0, _, _ = ctx.cluster_rerank_prefix(Q, rr.D, off, index.leaf_docids, ql, k, rr.BOOTSTRAP_ROWS); tau0 = s0[:, k - 1].contiguous(); t.append(ev())
    ctx.rerank_grouped_begin(Q, gr["absmax"], gr["maxnorm"], tau0); t.append(ev())
    plans = plan_grouped_rounds(off, gr["leaf_tile0"], ql, rr.ROUND_ROWS); t.append(ev())
    for it, ig, gq in plans:
        ctx.rerank_grouped_round(Q, gr["img"], gr["row0"], gr["nrows"], it, ig, gq, k); t.append(ev())
    sc, rows, fb = ctx.rerank_grouped_finish(Q, rr.D, k); t.append(ev())
    torch.cuda.synchronize()
    names = ["lookup", "bootstrap(prefix top-k)", "begin", "plan (torch)"] + [f"round {i} ({p[0].numel()} items, {p[2].numel()//64} groups)" for i, p in enumerate(plans)] + ["finish (rescore)"]
    if rep == 1:
        for nm, a, b in zip(names, t[:-1], t[1:]): print(f"{a.elapsed_time(b):8.3f} ms  {nm}")
        print(f"{t[0].elapsed_time(t[-1]):8.3f} ms  total; fell_back={fb}")
PY
