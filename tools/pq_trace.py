"""Print the per-warp timeline of a MEVI_PQ_TRACE dump (pq_tensor.cuh): clocks relative to the first event."""
import struct, sys
SLOTS, WARPS = 128, 16
raw = open(sys.argv[1], "rb").read()
v = struct.unpack(f"<{SLOTS * WARPS}Q", raw)
role = {0: "tma", 1: "mma0", 2: "mma1", 3: "bload", 4: "conv0", 5: "conv1", 6: "conv2", 7: "conv3", 8: "epiA0", 9: "epiA1", 10: "epiA2", 11: "epiA3",
        12: "epiB0", 13: "epiB1", 14: "epiB2", 15: "epiB3"}
names = {"tma": ["x_empty ok"], "mma": ["acc_empty ok", "a_full ok", "committed"], "conv": ["x_full ok", "a_empty ok", "published"],
         "epi": ["acc_full h0", "released h0", "acc_full h1", "released h1", "sub done"]}
ev = []
for w in range(WARPS):
    for s in range(SLOTS):
        x = v[w * SLOTS + s]
        if x: ev.append((x >> 16, w, (x >> 12) & 15, x & 4095))
t0 = min(e[0] for e in ev)
for w in sorted(set(e[1] for e in ev)):
    if w in (5, 6, 7, 9, 10, 11, 13, 14, 15): continue
    r = role[w]; nm = names["".join(c for c in r if c.isalpha())[:4].rstrip("AB")] if not r.startswith("epi") else names["epi"]
    print(f"--- warp {w} ({r})")
    last = None
    for t, _, e, q in sorted(x for x in ev if x[1] == w):
        print(f"  q {q:3d}  {nm[e] if e < len(nm) else e:14s} t={t - t0:7d}" + (f"  (+{t - last})" if last is not None else ""))
        last = t
