"""Condenses an .ncu-rep (ncu --set full) into the handful of per-launch metrics the roofline discussion uses.
Usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
        "smsp__cycles_active.avg", "launch__occupancy_limit_registers", "sm__cycles_elapsed.max"]

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
print(f"# ncu summary of `{rep}` (per launch; --set full --clock-control none)\n")
cols = [k for k in KEYS if k in idx]
print("| kernel | " + " | ".join(c.replace("avg.pct_of_peak_sustained_", "%").replace(".sum", "") for c in cols) + " |")
print("|---|" + "---|" * len(cols))
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[idx["Kernel Name"]].split("(")[0][-48:]
    vals = []
    for c in cols:
        v, u = r[idx[c]], units[idx[c]]
        vals.append(f"{v} {u}".strip())
    print(f"| {name} | " + " | ".join(vals) + " |")
