#!/bin/bash
timeout 60 python /dev/stdin <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
n, d, M, K = 4_000_000, 768, 4, 32
g = torch.Generator(device="cuda"); g.manual_seed(5)
X = torch.empty((n, d), device="cuda").normal_(generator=g)
cb = torch.empty((M, K, d // M), device="cuda").normal_(generator=g)
padded = torch.zeros((M, K, d), device="cuda")
for j in range(M): padded[j, :, j * (d // M):(j + 1) * (d // M)] = cb[j]
def t(fn):
    fn(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); fn(); fn(); b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / 2
c_pq = ctx.pq_encode(X, cb); c_rq, st = ctx.rq_encode(X, padded, mode="auto", return_stats=True)
bad = (c_pq != c_rq).any(1)
print(f"rows differing: {int(bad.sum())} of {n} (flagged to the exact kernel: {int(st[0])})")
if bad.any():
    idx = bad.nonzero().flatten()[:2000]; x = X[idx].double().view(-1, M, d // M)
    dist = ((x[:, :, None, :] - cb.double()[None]) ** 2).sum(-1)               # [r, M, K]
    da = dist.gather(2, c_pq[idx].long().unsqueeze(-1)).squeeze(-1); db = dist.gather(2, c_rq[idx].long().unsqueeze(-1)).squeeze(-1)
    print("max relative float64 distance gap at the differing entries:", float(((da - db).abs() / (torch.maximum(da, db) + (x ** 2).sum(-1))).max()))
print(f"pq_encode kernel {t(lambda: ctx.pq_encode(X, cb)):.2f} ms; rq tensor kernel on the padded codebook {t(lambda: ctx.rq_encode(X, padded, mode='auto')):.2f} ms")
PY
