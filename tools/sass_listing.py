"""profiles/r02_sass_tensor_instructions.txt: per kernel of libmevi_b200.so, how many Blackwell tensor / TMA instructions
its SASS holds (cuobjdump -sass; works without a GPU).  UTCHMMA = tcgen05.mma kind::f16, LDTM / STTM = tcgen05.ld / st,
UTMALDG = cp.async.bulk.tensor (TMA tiled load), UBLKCP = cp.async.bulk, SYNCS = mbarrier, UTCBAR = tcgen05.commit."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "mevi_b200", "libmevi_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCIMMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTCBAR|UTCCP|HMMA|UTMAPF|UTMACCTL)\b")
counts, order, cur = collections.defaultdict(collections.Counter), [], None
excerpt = {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
        order.append(cur)
        continue
    if cur is None:
        continue
    mm = pat.search(line)
    if mm:
        counts[cur][mm.group(1)] += 1
        if mm.group(1) in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP") and (cur, mm.group(1)) not in excerpt:
            excerpt[(cur, mm.group(1))] = re.sub(r"/\*[0-9a-fx]+\*/", "", line).strip()[:110]
out = os.path.join(ROOT, "profiles", "r02_sass_tensor_instructions.txt")
with open(out, "w") as fw:
    fw.write("# cuobjdump -sass mevi_b200/libmevi_b200.so (sm_100a): Blackwell tensor / TMA instructions per kernel (tools/sass_listing.py)\n")
    fw.write("# UTCHMMA = tcgen05.mma kind::f16 | LDTM/STTM = tcgen05.ld/st | UTMALDG = TMA tiled load | UBLKCP = cp.async.bulk | UTCBAR = tcgen05.commit\n\n")
    tot = collections.Counter()
    for k in order:
        if not counts[k]:
            continue
        tot.update(counts[k])
        fw.write(f"{k}\n    " + "  ".join(f"{n}={c}" for n, c in sorted(counts[k].items())) + "\n")
    fw.write("\nTOTAL  " + "  ".join(f"{n}={c}" for n, c in sorted(tot.items())) + "\n\n# first occurrence of each (kernel, mnemonic):\n")
    for (k, mn), l in excerpt.items():
        fw.write(f"{k[:60]:60s} {l}\n")
print(open(out).read()[:3000])
