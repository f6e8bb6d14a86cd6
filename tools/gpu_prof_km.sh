#!/bin/bash
cat > /tmp/km_driver.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach()
n = 2000000
X = torch.randn((n, 768), device="cuda")
C = cb[0].cuda().clone(); buf = torch.empty(32*768+32, device="cuda"); a = torch.empty(n, dtype=torch.int32, device="cuda")
for _ in range(3): ctx.kmeans_step(X, C, buf, assign=a, mode="auto")
torch.cuda.synchronize()
print(torch.bincount(a.long(), minlength=32).tolist())
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"kmeans_accumulate" -s 1 -c 1 -o gpurun_out/prof_km python /tmp/km_driver.py > gpurun_out/km_driver.txt 2>&1
tail -2 gpurun_out/km_driver.txt
