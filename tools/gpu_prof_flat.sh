#!/bin/bash
# per-kernel durations of one flat_ip_topk call + ncu --set full of its largest flat_gemm launch (6th of a call)
set -u
mkdir -p gpurun_out
CS=${1:-2}
cat > /tmp/fl.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
nq, n, d = 6980, 1 << 22, 768
Q = torch.randn((nq, d), device="cuda"); D = torch.randn((n, d), device="cuda")
for _ in range(2): ctx.flat_ip_topk(Q, D, 100, mode="tensor")
torch.cuda.synchronize()
PY
MEVI_FLAT_CLUSTER=$CS timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_flat.csv python /tmp/fl.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_flat.csv')) if len(r)>5]
h=[i for i,r in enumerate(rows) if r[0]=='ID'][0]; H=rows[h]
data=[r for r in rows[h+1:] if 'distribution' not in r[H.index('Kernel Name')]]
half=len(data)//2
for r in data[half:]:
    v=float(r[H.index('Metric Value')].replace(',','')); u=r[H.index('Metric Unit')]
    v*={'ns':1e-6,'us':1e-3,'ms':1.0}.get(u,1e-6)
    print(f"{v:10.3f} ms  {r[H.index('Kernel Name')][:70]}")
PY
MEVI_FLAT_CLUSTER=$CS timeout 280 ncu --set full --clock-control none --import-source on -k regex:flat_gemm_kernel -s 12 -c 1 \
   -f -o gpurun_out/prof_flat_cs$CS python /tmp/fl.py > gpurun_out/prof_flat_cs$CS.log 2>&1
echo "cs=$CS ncu rc=$?"
ncu -i gpurun_out/prof_flat_cs$CS.ncu-rep --page raw --csv > gpurun_out/prof_flat_cs$CS.raw.csv 2>/dev/null
