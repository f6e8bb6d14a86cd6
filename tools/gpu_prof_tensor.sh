#!/bin/bash
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
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tensor.csv python /tmp/prof_driver.py 8841823 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_tensor.csv')) if len(r)>5]
h=[i for i,r in enumerate(rows) if r[0]=='ID'][0]; H=rows[h]
for r in rows[h+1:]:
    n=r[H.index('Kernel Name')][:60]; v=r[H.index('Metric Value')]; u=r[H.index('Metric Unit')]
    if 'distribution' in n: continue
    print(f"{v:>14s} {u}  {n}")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_tensor3_kernel -s 2 -c 1 -o gpurun_out/prof_tensor python /tmp/prof_driver.py 2000000 > gpurun_out/prof_tensor.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/prof_tensor.log
