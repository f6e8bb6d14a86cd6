#!/bin/bash
cat > /tmp/kf.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
dev = torch.device("cuda", 0)
n, d, K = 3000000, 768, 32
X = torch.randn((n, d), device=dev)
C = X[torch.randint(0, n, (K,), device=dev)].clone()
buf = torch.empty(K * d + K, device=dev); a0 = torch.empty(n, dtype=torch.int32, device=dev); a1 = torch.empty_like(a0)
ctx.kmeans_step(X, C, buf, assign=a0, mode="tensor")
for _ in range(2): ctx.kmeans_step_fused(X, C, a0, a1, buf)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_tensor4_kernel -s 2 -c 1 -f -o gpurun_out/prof_kmfused python /tmp/kf.py > /dev/null 2> gpurun_out/prof_kmfused.err; echo "rc=$?"
ls -la gpurun_out/prof_kmfused.ncu-rep
