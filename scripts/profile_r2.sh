#!/bin/bash
# Runs on the GPU box (under gpurun).  $1 = tag.  One ncu capture (speed-of-light, compute / memory workload, launch and occupancy sections) of EVERY kernel of a few main frames
# (1080p, S = 1, one context, launches serialised), from which scripts/summarize_frame.py builds profiles/<tag>_frame.md
# and profiles/kernels_r2.json (warp instructions / DRAM bytes / pipe utilisation per main frame: bench.py's
# roofline_compute and roofline.traffic).
TAG=${1:-r2}
OURS='regex:^(vr_|raster_|tri_setup|resolve_depth|dilate_|row_prefix|shade_|mix_back|remap_|pyr_|absdiff|strided_copy|sobel_|triangulate_|normals_|count_|load_mesh|zero_channel|DeviceScan|.*normals_cov|.*vr_fused)'
mkdir -p gpurun_out
ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
    --metrics smsp__inst_executed.sum,smsp__cycles_active.avg,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k "$OURS" -s 75 -c 72 -f -o gpurun_out/frame_$TAG \
    python bench.py --steps 1 --warmup 1 --pairs 1 --contexts 1 --cpu-pairs 0 --link-probe-s 0 --no-filter-bench > gpurun_out/ncu_frame_$TAG.log 2>&1
tail -3 gpurun_out/ncu_frame_$TAG.log
ls -la gpurun_out | tail -4
