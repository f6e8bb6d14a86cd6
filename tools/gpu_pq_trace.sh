#!/bin/bash
# pipeline trace of the wide-codebook PQ kernel (CTA 0, sub-vector steps 40-51): full kernel and handshakes only
timeout 200 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
X = torch.randn((2000000, 768), device="cuda"); cb = torch.randn((24, 256, 32), device="cuda")
ctx.pq_encode(X, cb); torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
os.environ["MEVI_PQ_TRACE"] = "gpurun_out/pq_trace_full.bin"; ctx.pq_encode(X, cb); torch.cuda.synchronize()
os.environ["MEVI_RQ_DEBUG"] = "30"; os.environ["MEVI_PQ_TRACE"] = "gpurun_out/pq_trace_skel.bin"; ctx.pq_encode(X, cb); torch.cuda.synchronize()
PY
python tools/pq_trace.py gpurun_out/pq_trace_full.bin > gpurun_out/pq_trace_full.txt; python tools/pq_trace.py gpurun_out/pq_trace_skel.bin > gpurun_out/pq_trace_skel.txt; wc -l gpurun_out/pq_trace_*.txt
