#!/bin/bash
# grouped re-rank at the bench shape, default (tile-sliced) plan: where do the milliseconds go?
timeout 600 python - <<'PY'
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
from mevi_b200.pq import ProductQuantization
from mevi_b200.rerank import ClusterIndex, ClusterReranker, plan_grouped_tile_rounds
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
gg = rr._grouped
def T(fn, reps=5):
    for _ in range(2): r = fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3, r
ms, ql = T(lambda: index.lookup(dec)); print(f"lookup                {ms:6.2f} ms")
ms, plan = T(lambda: plan_grouped_tile_rounds(gg["leaf_tile0"], ql, 63)); print(f"planning              {ms:6.2f} ms")
ms, _ = T(lambda: ctx.rerank_grouped_begin(Q, gg["absmax"], gg["maxnorm"], None)); print(f"begin (query image)   {ms:6.2f} ms")
def rounds(which):
    ctx.rerank_grouped_begin(Q, gg["absmax"], gg["maxnorm"], None)
    for r in which:
        it, ig, gq = plan[r]
        ctx.rerank_grouped_round(Q, gg["img"], gg["row0"], gg["nrows"], it, ig, gq, 100)
ms0, _ = T(lambda: rounds([0])); print(f"begin + round 0       {ms0:6.2f} ms  ({plan[0][0].numel()} items)")
ms1, _ = T(lambda: rounds([0, 1])); print(f"begin + rounds 0,1    {ms1:6.2f} ms  ({plan[1][0].numel()} items)")
def full():
    rounds([0, 1]); return ctx.rerank_grouped_finish(Q, D_leaf, 100)
ms2, _ = T(full); print(f"... + finish          {ms2:6.2f} ms")
ms3, out = T(lambda: rr.rerank(Q, dec, topk=100)); print(f"rr.rerank             {ms3:6.2f} ms  path {rr.last_path} weak {rr.last_weak_queries}")
# leaves: how many queries per chosen leaf, tiles per chosen leaf
lv = ql[ql >= 0].long(); ul, cnt = torch.unique(lv, return_counts=True)
sizes = (index.leaf_offsets[1:] - index.leaf_offsets[:-1])[ul]
print("chosen leaves", ul.numel(), "pairs", lv.numel(), "mean queries/leaf", float(cnt.float().mean()), "rows in chosen leaves", int(sizes.sum()),
      "row-weighted queries/leaf", float((cnt * sizes).sum() / sizes.sum()), "tiles", int(((sizes + 127) // 128).sum()))
PY
