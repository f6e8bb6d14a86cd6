#!/bin/bash
# quick K1 check: tensor path vs direct fp32 kernel on 1M rows, then time at the bench size (tight timeouts: a hang must not eat the budget)
MEVI_RQ_EARLY=${MEVI_RQ_EARLY:-1} timeout 100 python /dev/stdin <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach().cuda()
X = torch.randn((1 << 20, 768), device="cuda")
a = ctx.rq_encode(X, cb, mode="auto"); torch.cuda.synchronize(); print("auto ok", flush=True)
e = ctx.rq_encode(X, cb, mode="exact")
print("mismatch rows vs exact:", int((a != e).any(1).sum()), flush=True)
for M in (1, 2, 3):
    a = ctx.rq_encode(X, cb[:M].contiguous(), mode="auto"); e = ctx.rq_encode(X[:200000], cb[:M].contiguous(), mode="exact")
    print("M", M, "mismatch", int((a[:200000] != e).any(1).sum()), flush=True)
n = 8841823
X = torch.randn((n, 768), device="cuda")
codes = torch.empty((n, 4), dtype=torch.int32, device="cuda")
for _ in range(3): ctx.rq_encode(X, cb, mode="auto", codes=codes)
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); s.record()
for _ in range(10): ctx.rq_encode(X, cb, mode="auto", codes=codes)
t.record(); torch.cuda.synchronize(); ms = s.elapsed_time(t) / 10
print(f"rq_encode {ms:.3f} ms  {n/ms/1e6:.3f} G docs/s  {n*3088/ms/1e6:.0f} GB/s", flush=True)
PY
echo "rc=$?"
