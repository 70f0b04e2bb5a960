"""Key metrics of every launch in ncu --set full reports, one CSV row per launch (runs on the GPU box so that only text travels back).

    python tools/ncu_summary.py out.csv a.ncu-rep b.ncu-rep ...
"""
import csv
import io
import os
import subprocess
import sys

COLS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "derived__lts__lts2xbar_bytes.sum.per_second", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]

out = csv.writer(open(sys.argv[1], "w", newline=""))
out.writerow(["capture"] + COLS + ["units"])
for path in sys.argv[2:]:
    if not os.path.exists(path) or os.path.getsize(path) == 0:
        continue
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(c) if c in hdr else -1 for c in COLS]
    u = ";".join(f"{c.split('.')[0]}={units[i]}" for c, i in zip(COLS, idx) if i >= 0 and units[i])
    for r in rows[2:]:
        out.writerow([os.path.basename(path).replace(".ncu-rep", "")] + [(r[i][:110] if i >= 0 else "") for i in idx] + [u])
