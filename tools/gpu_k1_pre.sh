#!/bin/bash
# K1 v4: early accumulator release variants (MEVI_RQ_EARLY = levels decided from registers after the release)
for pre in 0 2 3 4; do
MEVI_RQ_EARLY=$pre timeout 90 python /dev/stdin <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach().cuda()
n = 8841823
X = torch.randn((n, 768), device="cuda")
codes = torch.empty((n, 4), dtype=torch.int32, device="cuda")
for _ in range(3): ctx.rq_encode(X, cb, mode="auto", codes=codes)
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); s.record()
for _ in range(10): ctx.rq_encode(X, cb, mode="auto", codes=codes)
t.record(); torch.cuda.synchronize(); ms = s.elapsed_time(t) / 10
e = ctx.rq_encode(X[:300000], cb, mode="exact")
print(f"PRE={os.environ['MEVI_RQ_EARLY']}: rq_encode {ms:.3f} ms  {n/ms/1e6:.3f} G docs/s  {n*3088/ms/1e6:.0f} GB/s  mismatch vs exact {int((codes[:300000] != e).any(1).sum())}", flush=True)
PY
done
