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
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    t.record(); torch.cuda.synchronize(); return s.elapsed_time(t) / reps
ctx.kmeans_step(X, C, buf, assign=a0, mode="tensor")
for dbg, name in ((0, "full"), (128, "accumulators idle"), (256, "no sort (one bucket)"), (384, "no sort, accumulators idle"), (2, "no MMA"), (130, "no MMA, accumulators idle")):
    os.environ["MEVI_RQ_DEBUG"] = str(dbg)
    ms = timed(lambda: ctx.kmeans_step_fused(X, C, a0, a1, buf))
    print(f"debug {dbg:4d} ({name:28s}) {ms:7.3f} ms  frac {n*3072/ms/1e6/6541.8:.3f}", flush=True)
os.environ["MEVI_RQ_DEBUG"] = "0"
ctx.check()
PY
