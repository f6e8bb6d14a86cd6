#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full captures of the dominant kernels
set -u
mkdir -p gpurun_out
echo "== launch list"; timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err; echo "rc=$?"
echo "== full capture: rq_tensor4 (default K1 kernel) at bench size"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:rq_tensor4_kernel -s 3 -c 1 -f -o gpurun_out/prof_rq_encode python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2> gpurun_out/prof_rq.err; echo "rc=$?"
cat > /tmp/others.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
from mevi_b200.pq import ProductQuantization
from mevi_b200.rerank import ClusterIndex
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach()
n, nq = 2000000, 1480
X = torch.randn((n, 768), device="cuda")
codes = ctx.rq_encode(X, cb.cuda(), mode="tensor")
ex = ctx.rq_encode(X[:200000], cb.cuda(), mode="exact")
pq = ProductQuantization("rq", 4, 5, "l2", 768, "kmeans", "grad")
with torch.no_grad(): pq.codebook.copy_(cb)
Q = torch.randn((nq, 768), device="cuda")
dec = torch.cat([pq.beam_search(Q[a:a+128], 100) for a in range(0, nq, 128)])
index = ClusterIndex.from_codes(codes, 32)
ql = index.lookup(dec)
DL = ctx.gather_rows(X, index.leaf_docids)
for _ in range(2): ctx.cluster_rerank(Q, DL, index.leaf_offsets, index.leaf_docids, ql, 100, leaf_ordered=True)
C = cb[0].cuda().clone(); buf = torch.empty(32*768+32, device="cuda"); a = torch.empty(n, dtype=torch.int32, device="cuda")
for _ in range(2): ctx.kmeans_step(X, C, buf, assign=a, mode="auto")
Q2 = torch.randn((6980, 768), device="cuda")
for _ in range(2): ctx.flat_ip_topk(Q2, X, 100, mode="tensor")
torch.cuda.synchronize()
PY
echo "== full capture: other kernels"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rerank_stream_kernel|kmeans_accumulate_kernel|rq_exact_group_kernel|rq_tensor4_kernel<1>|to_fp16_image_kernel|flat_rescore_kernel|flat_tensor_compact_kernel" -c 16 -f -o gpurun_out/prof_others python /tmp/others.py > /dev/null 2> gpurun_out/prof_others.err; echo "rc=$?"
# the largest flat GEMM launch of a call (6th of 7; the second call = launches 7..13)
cat > /tmp/fl.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
Q = torch.randn((6980, 768), device="cuda"); D = torch.randn((1 << 22, 768), device="cuda")
for _ in range(2): ctx.flat_ip_topk(Q, D, 100, mode="tensor")
torch.cuda.synchronize()
PY
echo "== full capture: flat GEMM"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:flat_gemm_kernel -s 12 -c 1 -f -o gpurun_out/prof_flat_gemm python /tmp/fl.py > /dev/null 2> gpurun_out/prof_flat.err; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
