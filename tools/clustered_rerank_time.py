import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
from mevi_b200.pq import ProductQuantization
from mevi_b200.rerank import ClusterIndex, ClusterReranker
from mevi_b200.trainer import train_rq_lloyd
ctx = mevi_b200.get_context(0)
dev = torch.device("cuda", 0)
n, D = 8841823, 768
gc = torch.Generator(device=dev); gc.manual_seed(99)
centers = torch.empty((4096, D), device=dev).normal_(generator=gc)
gl = torch.Generator(device=dev); gl.manual_seed(99 + 1000)
X = torch.empty((n, D), device=dev)
for a in range(0, n, 1 << 20):
    b = min(a + (1 << 20), n)
    lab = torch.randint(0, 4096, (b - a,), device=dev, generator=gl)
    X[a:b].normal_(generator=gl).mul_(0.3).add_(centers[lab])
cb, codes = train_rq_lloyd(X, M=4, K=32, seed=41, iters=10, tol=None, device_index=0, presharded=True)
g = torch.Generator(device=dev); g.manual_seed(4321)
Q = torch.empty((6980, D), device=dev).normal_(generator=g)
pq = ProductQuantization("rq", 4, 5, "l2", D, "kmeans", "grad")
with torch.no_grad(): pq.codebook.copy_(cb.cpu())
codes = ctx.rq_encode(X, cb)
dec = torch.cat([pq.beam_search(Q[a:a + 1024], 100) for a in range(0, 6980, 1024)])
index = ClusterIndex.from_codes(codes, 32)
D_leaf = ctx.gather_rows(X, index.leaf_docids); del X
rr = ClusterReranker(None, index, mode="grouped", D_leaf=D_leaf)
for _ in range(2): out = rr.rerank(Q, dec, topk=100)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): out = rr.rerank(Q, dec, topk=100)
torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
print(f"clustered corpus: {ms:.2f} ms per call, path {rr.last_path}, weak {rr.last_weak_queries}, failed {rr.last_failed_queries}, mean candidates {float(out[2].float().mean()):.0f}")
rs = ClusterReranker(None, index, mode="stream", D_leaf=D_leaf)
ref = rs.rerank(Q, dec, topk=100)
print("ids equal to the streaming kernel's:", float((ref[1] == out[1]).float().mean()))
