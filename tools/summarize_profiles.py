"""Turn the ncu reports brought back in gpurun_out/ into the small tracked summaries under profiles/."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")

def ncu_csv(rep, page):
    txt = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(txt.splitlines()))

def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H = rows[h]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        v = float(r[H.index("Metric Value")].replace(",", "")); u = r[H.index("Metric Unit")]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        name = r[H.index("Kernel Name")]
        bare = name.replace("void ", "")
        mine = bare.startswith("<unnamed>::") and not bare.startswith("<unnamed>::elementwise")
        key = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:70]
        agg[("mevi_b200" if mine else "torch/cub") + "  " + key][0] += 1
        agg[("mevi_b200" if mine else "torch/cub") + "  " + key][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as fw:
        fw.write("# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised launches: compare SHARES)\n")
        fw.write("# command: python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-nq   (5 encode passes, e2e passes, the other configs)\n")
        fw.write(f"# total device time {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            fw.write(f"{v[1]:12.3f} ms {v[0]:6d}x {100 * v[1] / tot:6.2f}%  {k}\n")

KEYS = ["gpu__time_duration.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpc__cycles_elapsed.avg.per_second",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum"]

def kernels(rep, fw, seen, all_launches=False):
    rows = ncu_csv(rep, "raw")
    H, U = rows[0], rows[1]
    nth = collections.Counter()
    for r in rows[2:]:
        name = r[H.index("Kernel Name")]
        short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "").replace("(int)", "").replace("(bool)", "")
        nth[short] += 1
        if short in seen and not (all_launches and nth[short] <= 3): continue
        seen.add(short)
        fw.write(f"\n## {short}" + (f" (launch {nth[short]})" if all_launches else "") + "\n")
        for k in KEYS:
            if k in H: fw.write(f"  {k:78s} {r[H.index(k)]:>18s} {U[H.index(k)]}\n")
    return rows

os.makedirs(OUT, exist_ok=True)
launches(os.path.join(G, "launches_bench_r02.csv"), os.path.join(OUT, "r02_launches_bench.txt"))
seen = set()
with open(os.path.join(OUT, "r02_ncu_kernels.md"), "w") as fw:
    fw.write("# ncu --set full --clock-control none  (one launch per kernel; B200, round 2; tools/gpu_profiles_r02.sh)\n")
    fw.write("rq_tensor4_kernel<4, false> (the default K1 kernel) captured inside `bench.py` at the bench size (8,841,823 x 768); the others on "
             "a 3,000,000 x 768 corpus: grouped_gemm_kernel = the three rounds of one leaf-grouped re-rank call (6,980 queries x 100 leaves), "
             "rq_tensor4_kernel<1, true> = the one-pass k-means kernel, flat_gemm_kernel = two chunks of a 6,980-query search of a persistent "
             "index, pq_tensor_kernel<32> = the wide-codebook PQ encode (24 x 256).\n")
    rows = kernels(os.path.join(G, "prof_r02_k1.ncu-rep"), fw, seen)
    H = rows[0]; r = rows[2]
    rd = float(r[H.index("dram__bytes_read.sum")]); wr = float(r[H.index("dram__bytes_write.sum")])
    ur, uw = rows[1][H.index("dram__bytes_read.sum")], rows[1][H.index("dram__bytes_write.sum")]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic = rd * mult[ur] + wr * mult[uw]
    for rep in ("prof_r02_kmfused.ncu-rep", "prof_r02_grouped.ncu-rep", "prof_r02_flat.ncu-rep", "prof_r02_pq256.ncu-rep"):
        if os.path.isfile(os.path.join(G, rep)):
            seen.discard("rq_tensor4_kernel")
            kernels(os.path.join(G, rep), fw, seen, all_launches=rep in ("prof_r02_grouped.ncu-rep", "prof_r02_flat.ncu-rep"))
json.dump({"rq_encode_dram_bytes_per_launch": traffic, "source": "profiles/r02_ncu_kernels.md (dram__bytes_read.sum + dram__bytes_write.sum of rq_tensor4_kernel<4, false>, one launch at the bench size)"},
          open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
print("traffic", traffic)
