#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kmeans.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5
cat > /tmp/km_t.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach()
n = 4000000
X = torch.randn((n, 768), device="cuda")
C = cb[0].cuda().clone(); buf = torch.empty(32*768+32, device="cuda"); a = torch.empty(n, dtype=torch.int32, device="cuda")
for mode in ("auto",):
    for _ in range(2): ctx.kmeans_step(X, C, buf, assign=a, mode=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ctx.kmeans_step(X, C, buf, assign=a, mode=mode)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/5
    print(f"kmeans_step[{mode}] {ms:.3f} ms for {n} rows: {n*3072/ms/1e6:.0f} GB/s (one-pass bytes)")
PY
timeout 300 python /tmp/km_t.py
