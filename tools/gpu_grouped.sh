#!/bin/bash
# K3g (leaf-grouped tensor-core re-rank): parity tests, then bench-size timing against the streaming kernel
timeout 150 python -m pytest tests/test_gpu_rerank.py -m gpu -q -x -p no:cacheprovider -k "grouped" 2>&1 | tail -15
[ "${1:-}" = "tests" ] && exit 0
timeout 240 python /dev/stdin <<'PY'
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
from mevi_b200.pq import ProductQuantization
from mevi_b200.rerank import ClusterIndex, ClusterReranker
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach()
n, nq, L, k = int(os.environ.get("GR_N", 8841823)), 6980, 100, 100
g = torch.Generator(device="cuda"); g.manual_seed(1234)
X = torch.empty((n, 768), device="cuda")
for a in range(0, n, 1 << 20): X[a:a + (1 << 20)].normal_(generator=g)
codes = ctx.rq_encode(X, cb.cuda(), mode="auto")
g.manual_seed(4321)
Q = torch.empty((nq, 768), device="cuda").normal_(generator=g)
pq = ProductQuantization("rq", 4, 5, "l2", 768, "kmeans", "grad")
with torch.no_grad(): pq.codebook.copy_(cb)
dec = torch.cat([pq.beam_search(Q[a:a + 128], L) for a in range(0, nq, 128)])
index = ClusterIndex.from_codes(codes, 32)
t0 = time.time(); rr = ClusterReranker(X, index, mode="grouped"); torch.cuda.synchronize()
print(f"index: leaf-ordered copy + fp16 tile image in {time.time()-t0:.2f} s; tiles {rr._grouped['row0'].numel()} absmax {rr._grouped['absmax']:.3f} maxnorm {rr._grouped['maxnorm']:.2f}", flush=True)
del X
def run(mode):
    rr.mode = mode
    saved = rr._grouped
    if mode == "stream": rr._grouped = None
    for _ in range(2): out = rr.rerank(Q, dec, topk=k)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(3): out = rr.rerank(Q, dec, topk=k)
    b.record(); torch.cuda.synchronize()
    rr._grouped = saved
    return a.elapsed_time(b) / 3, out
ms_g, (sg, ig, ng) = run("grouped"); print(f"grouped: {ms_g:.2f} ms  {nq/ms_g*1e3:.0f} queries/s  path={rr.last_path}", flush=True)
ms_s, (ss, is_, ns) = run("stream"); print(f"stream : {ms_s:.2f} ms  {nq/ms_s*1e3:.0f} queries/s  path={rr.last_path}", flush=True)
print("ids identical:", float((ig == is_).float().mean()), " max |score diff|:", float((sg - ss).abs().max()), " ncand equal:", bool((ng == ns).all()))
PY
