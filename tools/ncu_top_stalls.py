"""Top stall sites of an ncu --import-source report (SASS view).  usage: ncu_top_stalls.py report.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = [i for i, r in enumerate(rows) if 'Source' in r][0]
H = rows[h]
col = H.index('Warp Stall Sampling (All Samples)'); ex = H.index('Instructions Executed'); si = H.index('Source')
stalls = [i for i, x in enumerate(H) if x.startswith('stall_') and 'Not Issued' not in x]
L = []; tot = 0
for idx, r in enumerate(rows[h + 1:]):
    try: v = int(r[col])
    except Exception: continue
    tot += v
    top = sorted(((int(r[i] or 0), H[i][6:]) for i in stalls), reverse=True)[:2]
    L.append((v, idx, r[ex], r[si].strip()[:70], top))
print("total samples", tot)
for v, idx, e, s, top in sorted(L, reverse=True)[:N]:
    print(f"{v:7d} line {idx:5d} exec {e:>9s}  {s:70s} {top}")
