#!/bin/bash
# incremental Lloyd iteration (mevi_kmeans_step_delta) at 8,841,823 x 768, K = 32: ms per iteration over 12 iterations
# from k-means++-like seeds, fraction of rows moved, against the fused and the two-pass iteration; full training time.
timeout 900 python - <<'PY'
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
from mevi_b200 import trainer
ctx = mevi_b200.get_context(0)
dev = torch.device("cuda", 0)
n, d, K = 8841823, 768, 32
g = torch.Generator(device=dev); g.manual_seed(1234)
X = torch.empty((n, d), device=dev)
for a in range(0, n, 1 << 20): X[a:a + (1 << 20)].normal_(generator=g)
g.manual_seed(41)
C = X[torch.randint(0, n, (K,), device=dev, generator=g)].clone()
buf = torch.empty(K * d + K, device=dev)
a, b = (torch.empty(n, dtype=torch.int32, device=dev) for _ in range(2))
master = torch.empty(K * d + K, dtype=torch.float64, device=dev)
nchg = torch.zeros(1, dtype=torch.int32, device=dev)
ctx.kmeans_step(X, C, buf, assign=a); master.copy_(buf); ctx.kmeans_update(buf, C)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(14)]
moved = []
for it in range(13):
    ev[it].record()
    ctx.kmeans_step_delta(X, C, a, b, master, buf, n_changed=nchg)
    ctx.kmeans_update(buf, C)
    a, b = b, a
    moved.append(nchg.clone())
ev[13].record(); torch.cuda.synchronize()
print("delta iterations: ms", [round(ev[i].elapsed_time(ev[i + 1]), 2) for i in range(13)])
print("moved fraction   ", [round(int(m.item()) / n, 4) for m in moved])
# drift check: running sums against a fresh accumulation under the final assignment
fresh = ctx.accumulate_by_code(X, a, K)
rel = float((buf[: K * d] - fresh[: K * d]).abs().max() / fresh[: K * d].abs().max())
print("running sums vs fresh accumulation: max rel diff", rel, "counts equal", bool(torch.equal(buf[K * d:], fresh[K * d:])))
for how in ("delta", "delta", "fused", "twopass"):
    trainer.LLOYD_ITERATION = how
    trainer.TRACE = how == "delta"
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cb, _ = trainer.train_rq_lloyd(X, M=4, K=32, seed=41, iters=25, tol=None, device_index=0, presharded=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    info = trainer.train_rq_lloyd.last_info
    print(how, f"train 4 x 25 iterations {dt:.3f} s, loop ms/iter", [round(l["loop_ms_per_iter"], 2) for l in info["levels"]],
          "mse", round(info["levels"][-1]["mse"], 5), "moved rows", [l["changed_rows"] for l in info["levels"]])
    if "trace" in info: print("   trace (ms):", info["trace"][:9], "...")
PY
