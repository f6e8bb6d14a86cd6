#!/bin/bash
# K1 generation 6 (hi.hi prefilter; undecided rows dumped for rq_refine6_kernel, or MEVI_RQ_REFINE=inline): parity vs the direct fp32 kernel, then timing at the bench size
timeout 600 python - <<'PY'
import os, sys, torch
os.environ['MEVI_RQ_KERNEL'] = '6'
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach().cuda()
torch.manual_seed(1)
X = torch.randn((1 << 20, 768), device="cuda")
a, st = ctx.rq_encode(X, cb, mode="tensor", return_stats=True); ctx.check(); print("v6 ran; stats", st.tolist(), flush=True)
e = ctx.rq_encode(X, cb, mode="exact")
print("M=4 mismatch rows vs exact:", int((a != e).any(1).sum()), "of", X.shape[0], flush=True)
for M in (1, 2, 3):
    c = cb[:M].contiguous()
    a, st = ctx.rq_encode(X, c, mode="tensor", return_stats=True); e = ctx.rq_encode(X[:200000], c, mode="exact")
    print("M", M, "mismatch", int((a[:200000] != e).any(1).sum()), "stats", st.tolist()[:3], flush=True)
# scaled / shifted data, ragged sizes
for scale, shift, n in ((1e-3, 0.0, 100003), (37.0, 0.5, 77777), (1.0, 0.0, 4097)):
    Y = (X[:n] * scale + shift).contiguous(); c = (cb * scale).contiguous()
    a, st = ctx.rq_encode(Y, c, mode="tensor", return_stats=True); e = ctx.rq_encode(Y, c, mode="exact")
    print(f"scale {scale} shift {shift} n {n}: mismatch", int((a != e).any(1).sum()), "stats", st.tolist()[:3], flush=True)
n = 8841823
X = torch.randn((n, 768), device="cuda")
codes = torch.empty((n, 4), dtype=torch.int32, device="cuda")
def run(tag, reps=20):
    for _ in range(3): ctx.rq_encode(X, cb, mode="tensor", codes=codes)
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): ctx.rq_encode(X, cb, mode="tensor", codes=codes)
    t.record(); torch.cuda.synchronize(); ms = s.elapsed_time(t) / reps
    print(f"{tag:40s} {ms:7.3f} ms  {n*3088/ms/1e6:7.0f} GB/s  frac {n*3088/ms/1e6/6541.8:.3f}", flush=True)
run("v6 default")
_, st = ctx.rq_encode(X, cb, mode="tensor", return_stats=True); print("stats at 8.84M:", st.tolist()[:3], flush=True)
e = ctx.rq_encode(X[:300000], cb, mode="exact"); print("mismatch vs exact on 300k:", int((codes[:300000] != e).any(1).sum()), flush=True)
for dbg, name in ((256, "refine kernel not launched"), (4, "no epilogue math")):
    os.environ["MEVI_RQ_DEBUG"] = str(dbg); run(f"v6 debug={dbg} ({name})", reps=10)
os.environ["MEVI_RQ_DEBUG"] = "0"
os.environ["MEVI_RQ_REFINE"] = "inline"; run("v6 inline refinement", reps=5); del os.environ["MEVI_RQ_REFINE"]
os.environ["MEVI_RQ_KERNEL"] = "4"; run("v4 (split fp16)", reps=10); os.environ["MEVI_RQ_KERNEL"] = "6"
run("v6 again")
ctx.check()
PY
echo "rc=$?"

MEVI_RQ_KERNEL=6 timeout 600 python -m pytest tests/test_gpu_rq_encode.py -m gpu -x -q 2>&1 | tail -3
