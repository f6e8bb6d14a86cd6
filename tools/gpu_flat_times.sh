#!/bin/bash
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
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_flat.csv python /tmp/fl.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_flat.csv')) if len(r)>5]
h=[i for i,r in enumerate(rows) if r[0]=='ID'][0]; H=rows[h]
agg=collections.OrderedDict()
data=rows[h+1:]
half=len(data)//2
for r in data[half:]:
    n=r[H.index('Kernel Name')][:60]; v=float(r[H.index('Metric Value')].replace(',','')); u=r[H.index('Metric Unit')]
    if 'distribution' in n: continue
    if u=='us': v*=1e3
    if u=='ms': v*=1e6
    agg.setdefault(n,[0,0.0]); agg[n][0]+=1; agg[n][1]+=v
for k,v in agg.items(): print(f"{v[1]/1e6:10.3f} ms {v[0]:3d}x  {k}")
for r in data[half:]:
    n=r[H.index('Kernel Name')]
    if 'gemm' in n: print('gemm launch', r[H.index('Metric Value')], r[H.index('Metric Unit')], r[H.index('Grid Size')] if 'Grid Size' in H else '')
PY
