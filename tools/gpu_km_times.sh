#!/bin/bash
cat > /tmp/km_t2.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach()
n = 4000000
X = torch.randn((n, 768), device="cuda")
C = cb[0].cuda().clone(); buf = torch.empty(32*768+32, device="cuda"); a = torch.empty(n, dtype=torch.int32, device="cuda")
for _ in range(3): ctx.kmeans_step(X, C, buf, assign=a, mode="auto")
torch.cuda.synchronize()
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none --csv --log-file gpurun_out/launches_km.csv python /tmp/km_t2.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_km.csv')) if len(r)>5]
h=[i for i,r in enumerate(rows) if r[0]=='ID'][0]; H=rows[h]
for r in rows[h+1:][-24:]:
    n=r[H.index('Kernel Name')][:60]; v=r[H.index('Metric Value')]; u=r[H.index('Metric Unit')]; m=r[H.index('Metric Name')]
    if 'distribution' in n: continue
    print(f"{v:>14s} {u:6s} {m[:24]:24s} {n}")
PY
