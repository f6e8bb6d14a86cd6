"""Per-kernel table of ONE grouped re-rank call from gpurun_out/grouped_call_kernels.csv (tools/gpu_prof_grouped.sh)."""
import collections
import csv
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/grouped_call_kernels.csv"
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
data = collections.OrderedDict()
for row in csv.DictReader(lines):
    data.setdefault((row["ID"], row["Kernel Name"][:44]), {})[row["Metric Name"]] = row["Metric Value"]
ks = list(data.items())
ks = ks[len(ks) // 2:]  # the second (warm) call
tot = 0.0
for (i, name), m in ks:
    g = lambda k: float(m.get(k, "0").replace(",", ""))
    t = g("gpu__time_duration.sum") / 1e6
    tot += t
    print(i, name.ljust(44), f"{t:6.3f} ms  dram rd {g('dram__bytes_read.sum') / 1e9:6.2f} GB wr {g('dram__bytes_write.sum') / 1e9:5.2f} GB"
          f"  L2 hit {g('lts__t_sector_hit_rate.pct'):3.0f}%  L2 bytes {g('lts__t_bytes.sum') / 1e9:5.1f} GB"
          f"  tensor {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):3.0f}%")
print(f"sum of kernel times {tot:.2f} ms (cold-cache, serialised under ncu)")
