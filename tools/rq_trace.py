"""Pipeline timeline of the tensor-path encode kernel (debug aid).

  MEVI_RQ_KERNEL=4 MEVI_RQ_TRACE=gpurun_out/trace.bin python tools/rq_trace.py run    # on the GPU box
  python tools/rq_trace.py show gpurun_out/trace.bin                                    # anywhere

The kernel (CTA 0, lane 0 of every warp) records clock64 at the hand-over points of the ring pipeline for a few
tile iterations; `show` prints, per chunk, when each role got what it was waiting for.
"""
import os, sys, struct
import numpy as np

SLOTS, WARPS = 512, 20
ROLE = {0: "tma", 1: "mma0", 2: "mma1", 3: "bprod"}
ROLE.update({w: f"conv{w-4}" for w in range(4, 12)})
ROLE.update({w: f"epi{w-12}" for w in range(12, 20)})


def run():
    import torch
    sys.path.insert(0, os.getcwd())
    import mevi_b200
    ctx = mevi_b200.get_context(0)
    cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach().cuda()
    X = torch.randn((2_000_000, 768), device="cuda")
    path = os.environ.pop("MEVI_RQ_TRACE")
    for _ in range(2):
        ctx.rq_encode(X, cb, mode="tensor")
    torch.cuda.synchronize()
    os.environ["MEVI_RQ_TRACE"] = path
    ctx.rq_encode(X, cb, mode="tensor")
    torch.cuda.synchronize()
    show(path)


def load(path):
    allw = np.fromfile(path, dtype=np.uint64)
    if allw.size >= WARPS * SLOTS + 4:
        c0, t0, c1, t1 = (int(v) for v in allw[WARPS * SLOTS:WARPS * SLOTS + 4])
        if t1 > t0:
            print(f"CTA 0 ran {c1 - c0} SM cycles in {(t1 - t0) / 1e3:.1f} us -> SM clock {(c1 - c0) / (t1 - t0) * 1e3:.0f} MHz")
    raw = allw[:WARPS * SLOTS].reshape(WARPS, SLOTS)
    ev = []
    for w in range(WARPS):
        for v in raw[w]:
            v = int(v)
            if v == 0:
                continue
            ev.append((v >> 16, w, (v >> 12) & 15, (v >> 8) & 15, v & 255))
    return sorted(ev)


def show(path):
    ev = load(path)
    if not ev:
        print("empty trace"); return
    v5 = os.environ.get("MEVI_RQ_KERNEL", "3") == "5" or "v5" in path
    t0 = ev[0][0]
    by = {}
    for t, w, e, it, c in ev:
        by.setdefault((w, e, it, c), t - t0)
    its = sorted({it for _, _, _, it, c in ev})
    for it in its:
        chunks = sorted({c for _, _, _, i, c in ev if i == it and c != 255})
        g = lambda w, e, c: by.get((w, e, it, c), -1)
        if v5:
            ew = range(12, 16) if it % 2 == 0 else range(16, 20)
            print(f"-- tile iteration {it}: epilogue start {g(ew[0],0,255)} drained min {min(g(w,1,255) for w in ew)} max {max(g(w,1,255) for w in ew)}")
            print("   chunk: tma slot free | conv: X landed, A free, published | mma: ready, issued | bprod slot free")
            prev = None
            for c in chunks:
                cw = 4 if g(4, 0, c) >= 0 else 8
                mw = 1 if g(1, 1, c) >= 0 else 2
                row = [g(0, 0, c), g(cw, 0, c), g(cw, 1, c), g(cw, 2, c), g(mw, 1, c), g(mw, 2, c), g(3, 0, c)]
                d = "" if prev is None else f"  d(issued) {row[5]-prev}"
                prev = row[5]
                print(f"   c{c:2d}: tma {row[0]:7d} | conv(w{cw}) {row[1]:7d} {row[2]:7d} {row[3]:7d} | mma(w{mw}) {row[4]:7d} {row[5]:7d} | b {row[6]:7d}{d}")
            continue
        print(f"-- tile iteration {it}: acc free mma0 {g(1,3,255)} mma1 {g(2,3,255)}; epilogue start {g(12,0,255)} end min {min(g(w,1,255) for w in range(12,20))} max {max(g(w,1,255) for w in range(12,20))}")
        prev = None
        for c in chunks:
            row = [g(0, 0, c), g(4, 0, c), g(4, 1, c), g(4, 2, c), g(11, 2, c), g(1, 0, c), g(1, 1, c), g(1, 2, c), g(2, 2, c), g(3, 0, c)]
            d = "" if prev is None else f"  d(mma0 issued) {row[7]-prev}"
            prev = row[7]
            print(f"   c{c:2d}: tma {row[0]:7d} | X {row[1]:7d} | Afree {row[2]:7d} | conv {row[3]:7d} {row[4]:7d} | mma0 {row[5]:7d} {row[6]:7d} {row[7]:7d} | mma1 {row[8]:7d} | b {row[9]:7d}{d}")


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run()
    else:
        show(sys.argv[2])
