#!/bin/bash
# flat IP: parity tests + whole-call time for each cluster size of the GEMM
for cs in 2 4 1; do
  echo "=== MEVI_FLAT_CLUSTER=$cs"
  MEVI_FLAT_CLUSTER=$cs timeout 600 python -m pytest tests/test_gpu_flat_ip.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3
  MEVI_FLAT_CLUSTER=$cs timeout 300 python /dev/stdin <<'PY'
import os, sys, torch, time
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
nq, n, d = 6980, 1 << 22, 768
Q = torch.randn((nq, d), device="cuda"); D = torch.randn((n, d), device="cuda")
for _ in range(2): s, i = ctx.flat_ip_topk(Q, D, 100, mode="tensor")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(3): s, i = ctx.flat_ip_topk(Q, D, 100, mode="tensor")
b.record(); torch.cuda.synchronize(); dt = a.elapsed_time(b) / 3e3
print(f"flat[tensor] {dt*1e3:.1f} ms  {2*nq*n*d/dt/1e12:.1f} TFLOP/s  {nq/dt:.0f} q/s at {n} docs")
se, ie = ctx.flat_ip_topk(Q[:512], D[:1 << 20], 100, mode="exact")
st, it = ctx.flat_ip_topk(Q[:512], D[:1 << 20], 100, mode="tensor")
print("ids equal:", float((it == ie).float().mean()), "max score diff", float((st - se).abs().max()))
PY
done
