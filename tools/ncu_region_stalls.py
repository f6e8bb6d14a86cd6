"""Samples per SASS line bucket with stall-reason totals.  usage: ncu_region_stalls.py report.ncu-rep [bucket]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 200
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = [i for i, r in enumerate(rows) if 'Source' in r][0]
H = rows[h]
col = H.index('Warp Stall Sampling (All Samples)'); ex = H.index('Instructions Executed')
stalls = [i for i, x in enumerate(H) if x.startswith('stall_') and 'Not Issued' not in x]
buckets = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for idx, r in enumerate(rows[h + 1:]):
    try: v = int(r[col])
    except Exception: continue
    b = buckets[idx // B]
    b[0] += v; b[1] += int(r[ex] or 0)
    for i in stalls: b[2][H[i][6:]] += int(r[i] or 0)
for k in sorted(buckets):
    s, e, c = buckets[k]
    if s < 50: continue
    print(f"lines {k*B:5d}-{k*B+B-1:5d} samples {s:7d} warp-instr {e:11d}  " + ", ".join(f"{n}={v}" for n, v in c.most_common(5)))
