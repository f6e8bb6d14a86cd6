#!/bin/bash
# per-kernel times / DRAM bytes / L2 hit rates of ONE grouped re-rank call at the bench shape (8,841,823 x 768, 6,980 x 100)
set -u
mkdir -p gpurun_out
cat > /tmp/grouped_full.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
from mevi_b200.pq import ProductQuantization
from mevi_b200.rerank import ClusterIndex, ClusterReranker
ctx = mevi_b200.get_context(0)
dev = torch.device("cuda", 0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach().cuda()
n = 8841823
g = torch.Generator(device=dev); g.manual_seed(1234)
X = torch.empty((n, 768), device=dev)
for a in range(0, n, 1 << 20): X[a:a + (1 << 20)].normal_(generator=g)
codes = ctx.rq_encode(X, cb)
g.manual_seed(4321)
Q = torch.empty((6980, 768), device=dev).normal_(generator=g)
pq = ProductQuantization("rq", 4, 5, "l2", 768, "kmeans", "grad")
with torch.no_grad(): pq.codebook.copy_(cb.cpu())
dec = torch.cat([pq.beam_search(Q[a:a + 1024], 100) for a in range(0, 6980, 1024)])
index = ClusterIndex.from_codes(codes, 32)
D_leaf = ctx.gather_rows(X, index.leaf_docids)
del X
rr = ClusterReranker(None, index, mode="grouped", D_leaf=D_leaf)
for _ in range(int(sys.argv[1])): rr.rerank(Q, dec, topk=100)
torch.cuda.synchronize()
print("path", rr.last_path, flush=True)
PY
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpc__cycles_elapsed.avg.per_second"
# the third call (two warm-up calls are skipped by counting the grouped call's own kernels: 3 rounds x (image, gemm, compact) ...)
timeout 900 ncu --metrics $M --clock-control none -k regex:'grouped_gemm|to_fp16_image|flat_tensor_compact|flat_rescore|plan_|rerank_stream|gr_|flat_absmax|flat_margin' --csv --log-file gpurun_out/grouped_call_kernels.csv python /tmp/grouped_full.py 2 > gpurun_out/grouped_call.log 2>&1
echo "rc=$?"
