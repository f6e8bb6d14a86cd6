#!/bin/bash
# source-level ncu capture of the opt-in v4 (TMEM operand) encode kernel + its ablations
set -u
mkdir -p gpurun_out
cat > /tmp/prof_driver.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach().cuda()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
X = torch.randn((n, 768), device="cuda")
for _ in range(3):
    ctx.rq_encode(X, cb, mode="tensor")
torch.cuda.synchronize()
PY
export MEVI_RQ_KERNEL=${MEVI_RQ_KERNEL:-4}

rm -f gpurun_out/prof_v4.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_tensor._kernel -s 2 -c 1 -o gpurun_out/prof_v4 python /tmp/prof_driver.py 2000000 > gpurun_out/prof_v4.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/prof_v4.log
