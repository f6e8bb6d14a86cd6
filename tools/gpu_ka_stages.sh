#!/bin/bash
for st in 2 3 4 5 6 8; do
MEVI_KA_STAGES=$st timeout 120 python /dev/stdin <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
n, d, K = 8841823, 768, 32
X = torch.randn((n, d), device="cuda")
assign = torch.randint(0, K, (n,), device="cuda", dtype=torch.int32)
assign[torch.rand(n, device="cuda") < 0.8] = 3   # skew like the reference-trained codebooks
buf = torch.empty(K * d + K, device="cuda")
for _ in range(2): ctx.accumulate_by_code(X, assign, K, buf)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(5): ctx.accumulate_by_code(X, assign, K, buf)
b.record(); torch.cuda.synchronize(); ms = a.elapsed_time(b) / 5
ref = torch.zeros(K, d, device="cuda", dtype=torch.float64).index_add_(0, assign[:1000000].long(), X[:1000000].double())
got = ctx.accumulate_by_code(X[:1000000], assign[:1000000], K)
err = ((got[:K*d].view(K, d).double() - ref).abs().max() / ref.abs().max()).item()
print(f"stages={os.environ['MEVI_KA_STAGES']}: {ms:.3f} ms  {n*d*4/ms/1e6:.0f} GB/s  rel err {err:.2e}")
PY
done
