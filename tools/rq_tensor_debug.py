"""GPU diagnostic for the tcgen05 RQ path: parity vs the exact kernel and the oracle, flag rates, timing."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import mevi_b200
from oracle import oracle

ctx = mevi_b200.get_context(0)
cb = torch.load(os.path.join(ROOT, "tests/golden/gauss768/codebook.pt"), map_location="cpu", weights_only=False).detach().numpy()
cbd = torch.from_numpy(cb).cuda()
for n in (128, 1000, 4096, 50000):
    rs = np.random.RandomState(n)
    X = rs.standard_normal((n, 768)).astype(np.float32)
    Xd = torch.from_numpy(X).cuda()
    ex = ctx.rq_encode(Xd, cbd, mode="exact").cpu().numpy()
    t0 = time.time()
    te, stats = ctx.rq_encode(Xd, cbd, mode="tensor", return_stats=True)
    torch.cuda.synchronize()
    te = te.cpu().numpy()
    mism = (ex != te).any(1)
    rep = oracle.classify_code_mismatches(X, cb, ex, te)
    print(f"n={n}: tensor vs exact mismatching rows {int(mism.sum())} (ties {rep['n_ties']}, hard {rep['n_hard']}, worst rel gap {rep['worst_rel_gap']:.2e}); "
          f"flagged {int(stats[0])}/{int(stats[1])}; first-level agreement {(ex[:,0]==te[:,0]).mean():.4f}; {time.time()-t0:.3f}s", flush=True)
    if rep["n_hard"]:
        r = rep["hard_rows"][0]
        print("  first hard row", r, "exact", ex[r], "tensor", te[r])
# k-means style M=1
C = cb[0:1].copy()
X = np.random.RandomState(5).standard_normal((20000, 768)).astype(np.float32)
a1 = ctx.rq_encode(torch.from_numpy(X).cuda(), torch.from_numpy(C).cuda(), mode="exact").cpu().numpy()
a2 = ctx.rq_encode(torch.from_numpy(X).cuda(), torch.from_numpy(C).cuda(), mode="tensor").cpu().numpy()
print("M=1 K=32: mismatches", int((a1 != a2).sum()))
# timing at 2M rows
n = 2_000_000
Xd = torch.randn((n, 768), device="cuda")
for mode in ("tensor", "exact"):
    for _ in range(2): ctx.rq_encode(Xd, cbd, mode=mode)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): _, st = ctx.rq_encode(Xd, cbd, mode=mode, return_stats=True)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"{mode}: {ms:.3f} ms for {n} rows -> {n/ms/1e3:.1f} M docs/s, {n*3072/ms/1e6:.0f} GB/s, flagged {int(st[0])}")
