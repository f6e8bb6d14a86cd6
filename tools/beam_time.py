import os, sys, torch
sys.path.insert(0, os.getcwd())
import mevi_b200
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach().cuda()
g = torch.Generator(device="cuda"); g.manual_seed(4321)
Q = torch.empty((6980, 768), device="cuda").normal_(generator=g)
for nb in (100, 10, 128, 500):
    for _ in range(2): out = ctx.rq_beam_search(Q, cb, nb, metric="l2", prod=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): out = ctx.rq_beam_search(Q, cb, nb, metric="l2", prod=True)
    b.record(); torch.cuda.synchronize()
    print(f"beam search 6980 queries x {nb} beams: {a.elapsed_time(b) / 5:.2f} ms", flush=True)
