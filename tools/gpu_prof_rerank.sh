#!/bin/bash
set -u
mkdir -p gpurun_out
cat > /tmp/rr_driver.py <<'PY'
import os, sys, torch, time
sys.path.insert(0, os.getcwd())
import mevi_b200
from mevi_b200.pq import ProductQuantization
from mevi_b200.rerank import ClusterIndex
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach()
n = int(sys.argv[1]); nq = int(sys.argv[2])
X = torch.randn((n, 768), device="cuda")
codes = ctx.rq_encode(X, cb.cuda(), mode="tensor")
pq = ProductQuantization("rq", 4, 5, "l2", 768, "kmeans", "grad")
with torch.no_grad(): pq.codebook.copy_(cb)
Q = torch.randn((nq, 768), device="cuda")
dec = torch.cat([pq.beam_search(Q[a:a+128], 100) for a in range(0, nq, 128)])
index = ClusterIndex.from_codes(codes, 32)
ql = index.lookup(dec)
for _ in range(2):
    s, i, nc = ctx.cluster_rerank(Q, X, index.leaf_offsets, index.leaf_docids, ql, 100)
torch.cuda.synchronize()
t0 = time.time()
s, i, nc = ctx.cluster_rerank(Q, X, index.leaf_offsets, index.leaf_docids, ql, 100)
torch.cuda.synchronize()
dt = time.time() - t0
print(f"rerank nq={nq} n={n}: {dt*1e3:.1f} ms, cand mean {nc.float().mean().item():.0f}, gathered {nc.double().sum().item()*3072/dt/1e9:.0f} GB/s")
# k-means step
C = cb[0].cuda().clone(); buf = torch.empty(32*768+32, device="cuda"); a = torch.empty(n, dtype=torch.int32, device="cuda")
for _ in range(3): ctx.kmeans_step(X, C, buf, assign=a, mode="auto")
torch.cuda.synchronize()
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_rr.csv python /tmp/rr_driver.py 4000000 1480 > gpurun_out/rr_driver.txt 2>&1
cat gpurun_out/rr_driver.txt | tail -3
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_rr.csv')) if len(r)>5]
h=[i for i,r in enumerate(rows) if r[0]=='ID'][0]; H=rows[h]
for r in rows[h+1:]:
    n=r[H.index('Kernel Name')][:70]; v=r[H.index('Metric Value')]; u=r[H.index('Metric Unit')]
    if any(t in n for t in ['rerank','kmeans','topk_merge','rq_tensor_kernel<1>','rq_exact']): print(f"{v:>14s} {u}  {n}")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rerank_kernel|kmeans_accumulate" -c 2 -o gpurun_out/prof_rr python /tmp/rr_driver.py 4000000 1480 > /dev/null 2>&1
echo "ncu rc=$?"
