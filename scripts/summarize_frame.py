#!/usr/bin/env python
"""Summarises gpurun_out/frame_<tag>.ncu-rep (ncu --set full of every kernel of a few main frames, scripts/profile_r2.sh)
into profiles/<tag>_frame.md and profiles/kernels_r2.json.  Run locally (ncu reads the report without a GPU)."""
import csv
import io
import json
import subprocess
import sys

tag = sys.argv[1]
W, H, S = 1920, 1080, 1
raw = subprocess.run(["ncu", "-i", f"gpurun_out/frame_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rd = list(csv.reader(io.StringIO(raw)))
hdr, data = rd[0], rd[2:]
col = {n: i for i, n in enumerate(hdr)}


def f(d, name):
    try:
        return float(d[col[name]].replace(",", ""))
    except Exception:
        return 0.0


names = [d[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "") for d in data]
# one complete main frame: from the tri_setup that follows a normals_finish to the next normals_finish
ends = [i for i, n in enumerate(names) if n.startswith("normals_finish")]
assert len(ends) >= 2, names
lo, hi = ends[0] + 1, ends[1] + 1
frame = list(range(lo, hi))
units = rd[1]
dur_unit = units[col["gpu__time_duration.sum"]]
scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(dur_unit, 1e-3)
rows, tot = [], {"us": 0.0, "inst": 0.0, "dram": 0.0}
for i in frame:
    d = data[i]
    us = f(d, "gpu__time_duration.sum") * scale
    inst = f(d, "smsp__inst_executed.sum")
    rb, wb = f(d, "dram__bytes_read.sum"), f(d, "dram__bytes_write.sum")
    bu = units[col["dram__bytes_read.sum"]]
    bs = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(bu, 1.0)
    cyc = f(d, "smsp__cycles_active.avg")
    rows.append({"kernel": names[i], "us": us, "warp_inst": inst, "dram_bytes": (rb + wb) * bs,
                 "issue_pct": 100.0 * inst / (148 * 4 * cyc) if cyc else 0.0,
                 "fp64_pct": f(d, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                 "fma_pct": f(d, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                 "xu_pct": f(d, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                 "alu_pct": f(d, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                 "dram_pct": f(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                 "warps_pct": f(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                 "regs": f(d, "launch__registers_per_thread")})
    tot["us"] += us
    tot["inst"] += inst
    tot["dram"] += (rb + wb) * bs
agg = {}
for r in rows:
    a = agg.setdefault(r["kernel"], {"n": 0, "us": 0.0, "warp_inst": 0.0, "dram_bytes": 0.0, "issue_w": 0.0, "fp64_w": 0.0, "xu_w": 0.0, "fma_w": 0.0, "alu_w": 0.0, "regs": r["regs"], "warps_w": 0.0, "dram_w": 0.0})
    a["n"] += 1
    for k_ in ("us", "warp_inst", "dram_bytes"):
        a[k_] += r[k_]
    for k_, s_ in (("issue_w", "issue_pct"), ("fp64_w", "fp64_pct"), ("xu_w", "xu_pct"), ("fma_w", "fma_pct"), ("alu_w", "alu_pct"), ("warps_w", "warps_pct"), ("dram_w", "dram_pct")):
        a[k_] += r[s_] * r["us"]
out = [f"# ncu (speed-of-light, compute / memory workload, launch, occupancy sections), every kernel of ONE main frame ({W}x{H}, S = {S}): {tag}", "",
       "Command: `scripts/profile_r2.sh` (bench.py --pairs 1 --contexts 1 under `ncu --clock-control none`); per-launch times are",
       "cold-cache and serialised: compare SHARES.  issue = warp instructions / (148 SMs x 4 schedulers x active cycles).", "",
       f"frame total: {tot['us']:.1f} us over {len(rows)} launches, {tot['inst'] / 1e6:.1f} M warp instructions, {tot['dram'] / 1e6:.1f} MB DRAM traffic "
       f"(algorithmic {54 * W * H / 1e6:.0f} MB)", "",
       "| kernel | launches | us | share | M warp-inst | issue % | FMA pipe % | ALU % | FP64 pipe % | XU % | DRAM % | warps active % | regs | DRAM MB |",
       "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
for k_, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    u = a["us"]
    out.append(f"| {k_[:60]} | {a['n']} | {u:.1f} | {100 * u / tot['us']:.1f}% | {a['warp_inst'] / 1e6:.2f} | {a['issue_w'] / u:.0f} | {a['fma_w'] / u:.0f} | {a['alu_w'] / u:.0f} | "
               f"{a['fp64_w'] / u:.0f} | {a['xu_w'] / u:.0f} | {a['dram_w'] / u:.1f} | {a['warps_w'] / u:.0f} | {a['regs']:.0f} | {a['dram_bytes'] / 1e6:.1f} |")
open(f"profiles/{tag}_frame.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
big = sorted(agg.items(), key=lambda kv: -kv[1]["us"])[:5]
js = {"shape": [W, H, S], "source": f"profiles/{tag}_frame.md (ncu, every kernel of one main frame)", "warp_inst_per_main_frame": tot["inst"],
      "dram_bytes_per_main_frame": tot["dram"], "us_per_main_frame_serialised": tot["us"],
      "pipes_pct_of_peak": {k_[:40]: {"share_of_frame_pct": round(100 * a["us"] / tot["us"], 1), "issue": round(a["issue_w"] / a["us"]), "fma": round(a["fma_w"] / a["us"]),
                                       "fp64": round(a["fp64_w"] / a["us"]), "xu": round(a["xu_w"] / a["us"]), "dram": round(a["dram_w"] / a["us"], 1)} for k_, a in big}}
json.dump(js, open("profiles/kernels_r2.json", "w"), indent=1)
