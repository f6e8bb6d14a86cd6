#!/bin/bash
# ncu evidence for profiles/ (round 2): launch list of the bench command + full captures of the judged kernels.
# Keep gpurun_out/ small (the merge back is capped at 64 MiB): one launch per kernel, no more than ~10 MB per report.
set -u
mkdir -p gpurun_out
echo "== launch list (bench, first 3000 launches)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-nq > gpurun_out/bench_under_ncu_r02.json 2> gpurun_out/bench_under_ncu_r02.err; echo "rc=$?"
echo "== full capture: K1 default (rq_tensor4<4>) at bench size"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_tensor4_kernel -s 3 -c 1 -f -o gpurun_out/prof_r02_k1 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2> gpurun_out/prof_r02_k1.err; echo "rc=$?"
cat > /tmp/others.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
from mevi_b200.pq import ProductQuantization
from mevi_b200.rerank import ClusterIndex, ClusterReranker
from mevi_b200.faiss_search import FlatIndex
what = sys.argv[1]
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach()
n, nq = 3000000, 6980
X = torch.randn((n, 768), device="cuda")
Q = torch.randn((nq, 768), device="cuda")
if what == "grouped":
    codes = ctx.rq_encode(X, cb.cuda(), mode="tensor")
    pq = ProductQuantization("rq", 4, 5, "l2", 768, "kmeans", "grad")
    with torch.no_grad(): pq.codebook.copy_(cb)
    dec = torch.cat([pq.beam_search(Q[a:a+1024], 100) for a in range(0, nq, 1024)])
    index = ClusterIndex.from_codes(codes, 32)
    rr = ClusterReranker(X, index, mode="grouped")
    for _ in range(2): rr.rerank(Q, dec, topk=100)
    print("grouped path:", rr.last_path, flush=True)
elif what == "kmeans":
    C = X[torch.randint(0, n, (32,), device="cuda")].clone(); buf = torch.empty(32*768+32, device="cuda")
    a0 = torch.empty(n, dtype=torch.int32, device="cuda"); a1 = torch.empty_like(a0)
    ctx.kmeans_step(X, C, buf, assign=a0, mode="auto")
    for _ in range(2): ctx.kmeans_step_fused(X, C, a0, a1, buf)
elif what == "flat":
    ix = FlatIndex(768, mode="tensor"); ix.add(X)
    for _ in range(2): ix.search_device(Q, 100)
elif what == "pq":
    cbp = torch.randn((24, 256, 32), device="cuda")
    for _ in range(2): ctx.pq_encode(X, cbp)
torch.cuda.synchronize()
PY
echo "== grouped GEMM rounds"; timeout 400 ncu --set full --clock-control none -k regex:grouped_gemm_kernel -s 3 -c 3 -f -o gpurun_out/prof_r02_grouped python /tmp/others.py grouped > gpurun_out/prof_r02_grouped.log 2>&1; echo "rc=$?"
echo "== fused k-means pass"; timeout 400 ncu --set full --clock-control none -k regex:rq_tensor4_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_kmfused python /tmp/others.py kmeans > gpurun_out/prof_r02_kmfused.log 2>&1; echo "rc=$?"
echo "== flat GEMM (persistent index)"; timeout 400 ncu --set full --clock-control none -k regex:flat_gemm_kernel -s 4 -c 2 -f -o gpurun_out/prof_r02_flat python /tmp/others.py flat > gpurun_out/prof_r02_flat.log 2>&1; echo "rc=$?"
echo "== PQ wide-codebook kernel"; timeout 400 ncu --set full --clock-control none -k regex:pq_tensor_kernel -s 1 -c 1 -f -o gpurun_out/prof_r02_pq256 python /tmp/others.py pq > gpurun_out/prof_r02_pq256.log 2>&1; echo "rc=$?"
rm -f gpurun_out/prof_kmfused.ncu-rep
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
