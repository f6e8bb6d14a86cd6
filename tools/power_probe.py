"""SM clock / board power / throttle reasons while the encode kernel runs back to back (NVML, 20 ms samples).

  MEVI_RQ_KERNEL=4 python tools/power_probe.py [seconds] [MEVI_RQ_DEBUG values, comma separated]
"""
import os, sys, threading, time, statistics
import torch
sys.path.insert(0, os.getcwd())
import mevi_b200
import pynvml

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "2"]
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
ctx = mevi_b200.get_context(0)
cb = torch.load("tests/golden/gauss768/codebook.pt", map_location="cpu", weights_only=False).detach().cuda()
n = 4_000_000
X = torch.randn((n, 768), device="cuda")
print(f"power limit {pynvml.nvmlDeviceGetPowerManagementLimit(h) / 1000:.0f} W, max SM clock {pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)} MHz")
for mode in modes:
    os.environ["MEVI_RQ_DEBUG"] = mode
    samples, stop = [], False
    def poll():
        while not stop:
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                            pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
            time.sleep(0.02)
    for _ in range(3): ctx.rq_encode(X, cb, mode="tensor")
    torch.cuda.synchronize()
    t = threading.Thread(target=poll); t.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); iters = 0
    a.record()
    while time.time() - t0 < secs:
        for _ in range(20): ctx.rq_encode(X, cb, mode="tensor")
        iters += 20
        torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    stop = True; t.join()
    ms = a.elapsed_time(b) / iters
    late = samples[len(samples) // 3:]
    reasons = 0
    for s in late: reasons |= s[2]
    print(f"debug={mode}: {ms:.3f} ms/launch ({n*3072/ms/1e6:.0f} GB/s) | SM clock median {statistics.median(s[0] for s in late):.0f} MHz, "
          f"power median {statistics.median(s[1] for s in late):.0f} W max {max(s[1] for s in late):.0f} W, event reasons 0x{reasons:x} ({len(late)} samples)")
