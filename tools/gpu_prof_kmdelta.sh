#!/bin/bash
# per-kernel time / DRAM bytes of incremental Lloyd iterations (mevi_kmeans_step_delta) at 8,841,823 x 768, K = 32
set -u
mkdir -p gpurun_out
cat > /tmp/km_delta_prof.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
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
for it in range(8):
    ctx.kmeans_step_delta(X, C, a, b, master, buf, n_changed=nchg)
    ctx.kmeans_update(buf, C)
    a, b = b, a
    print("iteration", it, "moved", int(nchg.item()), flush=True)
torch.cuda.synchronize()
PY
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpc__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
timeout 900 ncu --metrics $M --clock-control none -k regex:'rq_tensor4_kernel|kmeans_delta|DeviceSelect|DeviceCompact|kmeans_update|rq_exact' --csv --log-file gpurun_out/km_delta_kernels.csv python /tmp/km_delta_prof.py > gpurun_out/km_delta_prof.log 2>&1
echo "rc=$?"
