#!/bin/bash
# grouped re-rank at the bench shape: which threshold-sample sizes (BOOT_LEAVES) are best now that planning, compaction
# and the group images are cheap?
timeout 600 python - <<'PY'
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
from mevi_b200.pq import ProductQuantization
from mevi_b200.rerank import ClusterIndex, ClusterReranker
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
for boot in ((8, 100), (4, 100), (12, 100), (16, 100), (24, 100), (8, 32, 100), (4, 24, 100), (8, 63)):
    rr.BOOT_LEAVES = boot
    for _ in range(2): out = rr.rerank(Q, dec, topk=100)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(8): out = rr.rerank(Q, dec, topk=100)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 8 * 1e3
    print(f"BOOT_LEAVES {boot}: {ms:.2f} ms per call, path {rr.last_path}, weak {rr.last_weak_queries}, failed {rr.last_failed_queries}", flush=True)
PY
