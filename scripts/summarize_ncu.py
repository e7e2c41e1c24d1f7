#!/usr/bin/env python
"""Summarises gpurun_out/launches_<tag>.csv (ncu gpu__time_duration per launch) and
gpurun_out/full_<tag>.ncu-rep (one --set full capture) into profiles/<tag>_*.  Run locally."""
import csv
import io
import subprocess
import sys
from collections import defaultdict

tag = sys.argv[1]
rows = []
with open(f"gpurun_out/launches_{tag}.csv", newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(io.StringIO("".join(lines))):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        rows.append((r["Kernel Name"].split("(")[0], us))
tot = sum(u for _, u in rows)
agg = defaultdict(lambda: [0, 0.0])
for k, u in rows:
    agg[k][0] += 1
    agg[k][1] += u
out = [f"# ncu launch list summary: {tag}", "",
       "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -k <our kernels> python bench.py --steps 1 --warmup 1 --pairs 2 --cpu-pairs 0`",
       "(cold-cache, serialised launches: compare SHARES, not absolutes)", "",
       f"total launches {len(rows)}, total device time {tot / 1000:.3f} ms", "",
       "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
for k, (n, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {k} | {n} | {u:.1f} | {100 * u / tot:.1f}% | {u / n:.2f} |")
open(f"profiles/{tag}_launches.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:22]))

try:
    raw = subprocess.run(["ncu", "-i", f"gpurun_out/full_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    keep = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
            "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
            "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct")
    rd = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    lines = [f"# ncu --set full capture: {tag}", ""]
    for d in data:
        lines.append(f"## {d[hdr.index('Kernel Name')][:80]}")
        for name in keep[1:]:
            if name in hdr:
                i = hdr.index(name)
                lines.append(f"- {name}: {d[i]} {units[i]}")
        lines.append("")
    open(f"profiles/{tag}_full.md", "w").write("\n".join(lines))
    print("\n".join(lines[:40]))
except Exception as e:  # noqa: BLE001
    print("full capture summary failed:", e)
