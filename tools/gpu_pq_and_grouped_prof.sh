#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_modes.py -m gpu -q -x -p no:cacheprovider -k "pq" 2>&1 | tail -3
timeout 120 python /dev/stdin <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
n = 8841823
X = torch.randn((n, 768), device="cuda")
for M, K in ((4, 32), (24, 256)):
    cb = torch.randn((M, K, 768 // M), device="cuda"); codes = torch.empty((n, M), dtype=torch.int32, device="cuda")
    ctx.pq_encode(X, cb, codes=codes)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(2): ctx.pq_encode(X, cb, codes=codes)
    b.record(); torch.cuda.synchronize(); ms = a.elapsed_time(b) / 2
    print(f"pq_encode M={M} K={K}: {ms:.1f} ms  {n/ms/1e3:.1f} M docs/s  {3*K*768*n/ms/1e9:.1f} TFLOP/s fp32", flush=True)
PY
cat > /tmp/gr.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
from mevi_b200.pq import ProductQuantization
from mevi_b200.rerank import ClusterIndex, ClusterReranker
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach()
n, nq = 2000000, 6980
X = torch.randn((n, 768), device="cuda")
codes = ctx.rq_encode(X, cb.cuda(), mode="auto")
Q = torch.randn((nq, 768), device="cuda")
pq = ProductQuantization("rq", 4, 5, "l2", 768, "kmeans", "grad")
with torch.no_grad(): pq.codebook.copy_(cb)
dec = torch.cat([pq.beam_search(Q[a:a+128], 100) for a in range(0, nq, 128)])
rr = ClusterReranker(X, ClusterIndex.from_codes(codes, 32), mode="grouped")
for _ in range(2): rr.rerank(Q, dec, topk=100)
torch.cuda.synchronize(); print(rr.last_path)
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm_kernel -s 2 -c 2 -f -o gpurun_out/prof_grouped python /tmp/gr.py > gpurun_out/prof_grouped.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/prof_grouped.log
