"""DRAM traffic of one sample() pass from two ncu launch lists (dram__bytes_read.sum, dram__bytes_write.sum per kernel).

    python tools/traffic_from_ncu.py cfg2 tf32 gpurun_out/traffic_T2.csv gpurun_out/traffic_T3.csv 64 4096 >> profiles/r02_traffic.json

The two CSVs are `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` captures of
tools/profile_step.py with 2 and 3 timesteps (1 and 2 ADPM2 iterations, graphs off): their difference is one iteration, the rest is
the one-time work of a call (time tables, conditioning K/V).  Scaled to `timesteps` it is the traffic of a full pass."""
import collections
import csv
import json
import sys


def totals(path):
    lines = [l for l in open(path, newline="") if l.startswith('"')]
    by_metric = collections.Counter()
    by_kernel = collections.defaultdict(lambda: collections.Counter())
    n = 0
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        u = r.get("Metric Unit", "")
        name = r["Metric Name"]
        if name.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        elif name == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
            n += 1
        by_metric[name] += v
        by_kernel[r["Kernel Name"].split("(")[0]][name] += v
    return by_metric, by_kernel, n


name, prec, p2, p3, T, B = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
m2, k2, n2 = totals(p2)
m3, k3, n3 = totals(p3)
rd = lambda m: m["dram__bytes_read.sum"]
wr = lambda m: m["dram__bytes_write.sum"]
it_r, it_w = rd(m3) - rd(m2), wr(m3) - wr(m2)
once_r, once_w = rd(m2) - it_r, wr(m2) - it_w
total = once_r + once_w + (T - 1) * (it_r + it_w)
top = sorted(((k3[k]["dram__bytes_read.sum"] + k3[k]["dram__bytes_write.sum"] - k2[k]["dram__bytes_read.sum"] - k2[k]["dram__bytes_write.sum"], k)
              for k in k3), reverse=True)[:6]
print(json.dumps({f"{name}_{prec}": {
    "dram_bytes_per_sample_pass": total, "batch": B, "timesteps": T,
    "per_iteration_read": it_r, "per_iteration_write": it_w, "one_time_read": once_r, "one_time_write": once_w,
    "launches_per_iteration": n3 - n2, "kernel_time_per_iteration_us_cold": m3["gpu__time_duration.sum"] - m2["gpu__time_duration.sum"],
    "top_kernels_per_iteration_bytes": {k: v for v, k in top},
    "method": "ncu dram__bytes_read.sum + dram__bytes_write.sum of every kernel, T=3 minus T=2 captures of tools/profile_step.py (graphs off)"}}))
