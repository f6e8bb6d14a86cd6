#!/bin/bash
# grouped re-rank at the bench shape: device-side plan (csrc/rerank_plan.cu) and wide items (a tile meets up to MAXG query
# groups) against the torch plan with one group per item.  Prints ms per phase and per configuration.
timeout 900 python - <<'PY'
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
print("leaves in the index", index.n_leaves, "tiles", gg["row0"].numel())
def T(fn, reps=5):
    for _ in range(2): r = fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3, r
ms, ql = T(lambda: index.lookup(dec)); print(f"lookup                {ms:6.2f} ms")
ms, plan_t = T(lambda: plan_grouped_tile_rounds(gg["leaf_tile0"], ql, (8, 63))); print(f"torch plan            {ms:6.2f} ms")
off = index.leaf_offsets
def dplan(ms_, ml_):
    ncand, weak, rounds, nweak = ctx.rerank_grouped_plan(ql, off, gg["leaf_tile0"], (8, 63), 2048, 100, ms_, ml_)
    return [ctx.rerank_grouped_plan_fill(r, a, b, dev) for r, (a, b) in enumerate(rounds)], nweak
for ms_, ml_ in ((1, 1), (1, 2), (1, 4), (2, 4), (4, 4)):
    ms, (plan, nweak) = T(lambda: dplan(ms_, ml_))
    print(f"device plan maxg=({ms_},{ml_}) {ms:6.2f} ms  items {[p[0].numel() for p in plan]} groups {[p[2].numel() // 64 for p in plan]} weak {nweak}")
    def rounds(which):
        ctx.rerank_grouped_begin(Q, gg["absmax"], gg["maxnorm"], None)
        for r in which:
            it, ig, gq = plan[r]
            ctx.rerank_grouped_round(Q, gg["img"], gg["row0"], gg["nrows"], it, ig, gq, 100, ml_ if r == 2 else ms_)
    m0, _ = T(lambda: rounds([0])); m1, _ = T(lambda: rounds([0, 1])); m2, _ = T(lambda: rounds([0, 1, 2]))
    print(f"    begin+r0 {m0:6.2f}   +r1 {m1:6.2f}   +r2 {m2:6.2f} ms")
for planname, ms_, ml_ in (("tiles", 1, 1), ("device", 1, 1), ("device", 1, 2), ("device", 1, 4), ("device", 2, 4), ("device", 4, 4)):
    rr.PLAN, rr.MAXG_SAMPLE, rr.MAXG_LAST = planname, ms_, ml_
    ms3, out = T(lambda: rr.rerank(Q, dec, topk=100), reps=8)
    print(f"rr.rerank plan={planname} maxg=({ms_},{ml_})  {ms3:6.2f} ms  path {rr.last_path} weak {rr.last_weak_queries} failed {rr.last_failed_queries}")
    if planname == "tiles":
        ref = out
    else:
        same = float((out[1] == ref[1]).float().mean()); ds = float((out[0] - ref[0]).abs().max())
        print(f"    ids equal to the torch plan's: {same:.6f}  max |score diff| {ds:.3g}  ncand equal {bool((out[2] == ref[2]).all())}")
lv = ql[ql >= 0].long(); ul, cnt = torch.unique(lv, return_counts=True)
sizes = (index.leaf_offsets[1:] - index.leaf_offsets[:-1])[ul]
gpl = (cnt + 63) // 64
tiles = (sizes + 127) // 128
print("chosen leaves", ul.numel(), "pairs", lv.numel(), "tiles", int(tiles.sum()), "tile-weighted groups/leaf", float((gpl * tiles).sum() / tiles.sum()),
      "hist of groups per tile", torch.bincount(torch.repeat_interleave(gpl, tiles).clamp(max=9)).tolist())
PY
