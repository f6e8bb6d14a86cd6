#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_rerank.py tests/test_gpu_kmeans.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
timeout 300 python /dev/stdin <<'PY'
import os, sys, torch, time
sys.path.insert(0, os.getcwd())
import mevi_b200
from mevi_b200.pq import ProductQuantization
from mevi_b200.rerank import ClusterIndex
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach()
n, nq = 4000000, 1480
X = torch.randn((n, 768), device="cuda")
codes = ctx.rq_encode(X, cb.cuda(), mode="tensor")
pq = ProductQuantization("rq", 4, 5, "l2", 768, "kmeans", "grad")
with torch.no_grad(): pq.codebook.copy_(cb)
Q = torch.randn((nq, 768), device="cuda")
dec = torch.cat([pq.beam_search(Q[a:a+128], 100) for a in range(0, nq, 128)])
index = ClusterIndex.from_codes(codes, 32)
ql = index.lookup(dec)
t0 = time.time(); DL = ctx.gather_rows(X, index.leaf_docids); torch.cuda.synchronize(); print(f"permute {time.time()-t0:.3f}s")
for name, D, lo in (("doc-order gather", X, False), ("leaf-order stream", DL, True)):
    for _ in range(2): s, i, nc = ctx.cluster_rerank(Q, D, index.leaf_offsets, index.leaf_docids, ql, 100, leaf_ordered=lo)
    torch.cuda.synchronize(); t0 = time.time()
    s, i, nc = ctx.cluster_rerank(Q, D, index.leaf_offsets, index.leaf_docids, ql, 100, leaf_ordered=lo)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"{name}: {dt*1e3:.1f} ms, cand mean {nc.float().mean().item():.0f}, {nq/dt:.0f} q/s, gathered {nc.double().sum().item()*3072/dt/1e9:.0f} GB/s")
    if lo: s1, i1 = s, i
    else: s0, i0 = s, i
print("ids equal:", bool((i0 == i1).all()), "scores close:", bool(torch.allclose(s0, s1, rtol=1e-5, atol=1e-4)))
C = cb[0].cuda().clone(); buf = torch.empty(32*768+32, device="cuda"); a = torch.empty(n, dtype=torch.int32, device="cuda")
for _ in range(2): ctx.kmeans_step(X, C, buf, assign=a, mode="auto")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ctx.kmeans_step(X, C, buf, assign=a, mode="auto")
e1.record(); torch.cuda.synchronize()
print(f"kmeans_step {e0.elapsed_time(e1)/5:.3f} ms for {n} rows")
PY
