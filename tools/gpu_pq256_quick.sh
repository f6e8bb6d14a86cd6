#!/bin/bash
# first contact with the wide-codebook PQ kernel: small sizes under a short timeout (a protocol bug must not eat GPU minutes)
timeout 90 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
torch.manual_seed(3)
for n in (4096, 20000, 200003):
    X = torch.randn((n, 768), device="cuda"); cb = torch.randn((24, 256, 32), device="cuda")
    os.environ["MEVI_PQ_TENSOR"] = "1"; a = ctx.pq_encode(X, cb); torch.cuda.synchronize(); print("tensor ran", flush=True); ctx.check()
    os.environ["MEVI_PQ_TENSOR"] = "0"; e = ctx.pq_encode(X, cb); ctx.check()
    print(f"n {n}: mismatching codes {int((a != e).sum())} of {a.numel()}", flush=True)
PY
echo "rc=$?"
