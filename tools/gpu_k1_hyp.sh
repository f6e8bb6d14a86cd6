#!/bin/bash
# K1 hypothesis test: how much of the M=4 kernel's time is the tensor work (power / pipe)?  debug 32 = hi.hi MMA only.
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv
(nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 100 > gpurun_out/k1_hyp_clocks.txt) &
SMI=$!
timeout 300 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach().cuda()
n = 8841823
X = torch.randn((n, 768), device="cuda")
codes = torch.empty((n, 4), dtype=torch.int32, device="cuda")
def run(tag, M=4, reps=20):
    c = cb[:M].contiguous(); co = codes if M == 4 else torch.empty((n, M), dtype=torch.int32, device="cuda")
    for _ in range(3): ctx.rq_encode(X, c, mode="tensor", codes=co)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): ctx.rq_encode(X, c, mode="tensor", codes=co)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"{tag:44s} {ms:7.3f} ms  {n*3088/ms/1e6:7.0f} GB/s  frac {n*3088/ms/1e6/6541.8:.3f}", flush=True)
for dbg, name in ((0, "full"), (32, "hi.hi MMA only"), (2, "no MMA"), (4, "no epilogue math"), (36, "hi.hi only, no epilogue math"), (0, "full again")):
    os.environ["MEVI_RQ_DEBUG"] = str(dbg)
    run(f"M=4 debug={dbg} ({name})")
os.environ["MEVI_RQ_DEBUG"] = "0"
for M in (1, 2, 3):
    run(f"M={M} full", M=M)
PY
kill $SMI
sort gpurun_out/k1_hyp_clocks.txt | uniq -c | sort -k1nr | head -12
