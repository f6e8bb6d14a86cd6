#!/bin/bash
# ncu --set full capture of the wide-codebook PQ kernel (pq_tensor_kernel) on 2,000,000 x 768, 24 x 256
cat > /tmp/pq256.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
torch.manual_seed(3)
X = torch.randn((2000000, 768), device="cuda"); cb = torch.randn((24, 256, 32), device="cuda")
for _ in range(2): ctx.pq_encode(X, cb)
torch.cuda.synchronize(); ctx.check()
PY
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pq_tensor_kernel -s 1 -c 1 -f -o gpurun_out/prof_r02_pq256 python /tmp/pq256.py > gpurun_out/prof_r02_pq256.log 2>&1; echo "rc=$?"
ls -la gpurun_out/prof_r02_pq256.ncu-rep
