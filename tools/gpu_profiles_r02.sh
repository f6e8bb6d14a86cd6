#!/bin/bash
# ncu evidence for profiles/ (round 2): launch list of the bench command + full captures of the judged kernels
set -u
mkdir -p gpurun_out
echo "== launch list"; timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-nq > gpurun_out/bench_under_ncu_r02.json 2> gpurun_out/bench_under_ncu_r02.err; echo "rc=$?"
echo "== full capture: K1 default (rq_tensor4<4>) at bench size"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:rq_tensor4_kernel -s 3 -c 1 -f -o gpurun_out/prof_r02_k1 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2> gpurun_out/prof_r02_k1.err; echo "rc=$?"
echo "== full capture: K1 generation 6"; MEVI_RQ_KERNEL=6 timeout 900 ncu --set full --clock-control none --import-source on -k regex:rq_tensor6_kernel -s 3 -c 1 -f -o gpurun_out/prof_r02_k1v6 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2> gpurun_out/prof_r02_k1v6.err; echo "rc=$?"
cat > /tmp/others.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
from mevi_b200.pq import ProductQuantization
from mevi_b200.rerank import ClusterIndex, ClusterReranker
from mevi_b200.faiss_search import FlatIndex
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach()
n, nq = 3000000, 6980
X = torch.randn((n, 768), device="cuda")
codes = ctx.rq_encode(X, cb.cuda(), mode="tensor")
pq = ProductQuantization("rq", 4, 5, "l2", 768, "kmeans", "grad")
with torch.no_grad(): pq.codebook.copy_(cb)
Q = torch.randn((nq, 768), device="cuda")
dec = torch.cat([pq.beam_search(Q[a:a+1024], 100) for a in range(0, nq, 1024)])
index = ClusterIndex.from_codes(codes, 32)
rr = ClusterReranker(X, index, mode="grouped")
for _ in range(2): rr.rerank(Q, dec, topk=100)
print("grouped path:", rr.last_path, flush=True)
C = X[torch.randint(0, n, (32,), device="cuda")].clone(); buf = torch.empty(32*768+32, device="cuda")
a0 = torch.empty(n, dtype=torch.int32, device="cuda"); a1 = torch.empty_like(a0)
ctx.kmeans_step(X, C, buf, assign=a0, mode="auto")
for _ in range(2): ctx.kmeans_step_fused(X, C, a0, a1, buf)
cbp = torch.randn((4, 32, 192), device="cuda")
for _ in range(2): ctx.pq_encode(X, cbp)
ix = FlatIndex(768, mode="tensor"); ix.add(X)
for _ in range(2): ix.search_device(Q, 100)
torch.cuda.synchronize()
PY
echo "== full capture: grouped GEMM rounds, fused k-means, flat GEMM (persistent index), accumulate"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"grouped_gemm_kernel|rq_tensor4_kernel<1, 1>|kmeans_accumulate_kernel|flat_gemm_kernel|rerank_stream_kernel" -c 24 -f -o gpurun_out/prof_r02_others python /tmp/others.py > gpurun_out/prof_r02_others.log 2> gpurun_out/prof_r02_others.err; echo "rc=$?"
ls -la gpurun_out/prof_r02_*.ncu-rep
