#!/bin/bash
timeout 600 python - <<'PY'
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
dev = torch.device("cuda", 0)
n, d, K = 8841823, 768, 32
g = torch.Generator(device=dev); g.manual_seed(1234)
X = torch.empty((n, d), device=dev)
for a in range(0, n, 1 << 20): X[a:a + (1 << 20)].normal_(generator=g)
C = X[torch.randint(0, n, (K,), device=dev, generator=g)].clone()
buf = torch.empty(K * d + K, device=dev); a0 = torch.empty(n, dtype=torch.int32, device=dev); a1 = torch.empty_like(a0)
def T(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3, r
ctx.kmeans_step(X, C, buf, assign=a0, mode="tensor"); ctx.kmeans_update(buf, C)
for it in range(8):
    t_f, _ = T(lambda: ctx.kmeans_step_fused(X, C, a0, a1, buf))
    t_nz, ch = T(lambda: torch.nonzero(a1 != a0).squeeze(1))
    nc = ch.numel()
    t_g, moved = T(lambda: ctx.gather_rows(X, ch.to(torch.int32)))
    t_a, (plus, minus) = T(lambda: (ctx.accumulate_by_code(moved, a1[ch].contiguous(), K), ctx.accumulate_by_code(moved, a0[ch].contiguous(), K)))
    t_u, _ = T(lambda: (buf.add_(plus).sub_(minus), ctx.kmeans_update(buf, C)))
    cnts = torch.bincount(a1.long(), minlength=K)
    print(f"it {it}: fused {t_f:6.2f} ms | changed {nc:8d} ({nc/n*100:5.1f}%) nonzero {t_nz:5.2f} gather {t_g:5.2f} 2x accumulate {t_a:5.2f} update {t_u:5.2f} | cluster sizes min {int(cnts.min())} max {int(cnts.max())}", flush=True)
    a0, a1 = a1, a0
PY
