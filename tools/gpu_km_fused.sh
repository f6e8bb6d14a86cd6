#!/bin/bash
# fused (one-pass) k-means iteration: parity tests, then timing at the bench size against the two-pass iteration
timeout 600 python -m pytest tests/test_gpu_kmeans.py -m gpu -x -q 2>&1 | tail -15
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
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    t.record(); torch.cuda.synchronize(); return s.elapsed_time(t) / reps
ms2 = timed(lambda: ctx.kmeans_step(X, C, buf, assign=a0, mode="tensor"))
ctx.kmeans_step(X, C, buf, assign=a0, mode="tensor")
msf = timed(lambda: ctx.kmeans_step_fused(X, C, a0, a1, buf))
ms_assign = timed(lambda: ctx.rq_encode(X, C[None], mode="tensor", codes=a1.view(n, 1)))
print(f"two-pass step {ms2:.3f} ms | fused pass {msf:.3f} ms ({n*3072/msf/1e6:.0f} GB/s, frac {n*3072/msf/1e6/6541.8:.3f}) | assign alone {ms_assign:.3f} ms", flush=True)
ctx.check()
from mevi_b200 import trainer
for fused in (True, False):
    trainer.FUSED_LLOYD = fused
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cb, codes = trainer.train_rq_lloyd(X, M=2, K=32, seed=41, iters=25, tol=None, device_index=0, presharded=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    info = trainer.train_rq_lloyd.last_info
    print(f"train fused={fused}: {dt:.2f} s, levels:", [(round(l['loop_ms_per_iter'], 2), l.get('fused_iters'), l.get('changed_rows'), round(l['mse'], 5)) for l in info['levels']], flush=True)
PY
