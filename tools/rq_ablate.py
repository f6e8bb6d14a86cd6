import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach().cuda()
n = 4_000_000
X = torch.randn((n, 768), device="cuda")
def run(tag):
    for _ in range(2): ctx.rq_encode(X, cb, mode="tensor")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): ctx.rq_encode(X, cb, mode="tensor")
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"{tag:40s} {ms:7.3f} ms  {n*3072/ms/1e6:7.0f} GB/s", flush=True)
names = {0: "full", 1: "no B copies", 2: "no MMA", 4: "no epilogue math", 8: "no convert/STS", 3: "no B, no MMA", 7: "no B/MMA/epi", 15: "loads only", 12: "no epi, no convert", 6: "no MMA, no epi", 16: "codebook copied twice", 20: "codebook twice, no epi"}
order = [int(k) for k in sys.argv[1].split(",")] if len(sys.argv) > 1 else list(names)
for k in order:
    v = names.get(k, "?")
    os.environ["MEVI_RQ_DEBUG"] = str(k)
    run(f"debug={k} ({v})")
