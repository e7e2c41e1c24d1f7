#!/bin/bash
# Runs on the GPU box (under gpurun).  $1 = tag (e.g. r1_v1).  Produces, under gpurun_out/:
#   launches_$1.csv  -- every launch of OUR kernels with its device time (ncu, cold-cache, serialised)
#   full_$1.ncu-rep  -- one --set full capture of the kernel matching $2 (regex), 3 launches
#   bench_$1.json    -- a clean (un-profiled) bench line
TAG=${1:-run}
KREGEX=${2:-normals_kernel}
OURS='regex:^(vr_|raster_|tri_setup|resolve_depth|dilate_|row_prefix|shade_|mix_back|remap_|pyr_|absdiff|strided_copy|sobel_|triangulate_|deh_|normals_|moments_|normals_|count_|load_mesh|zero_channel|DeviceScan)'
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 600 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --pairs 2 --cpu-pairs 0 > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" -s 2 -c 3 -f -o gpurun_out/full_$TAG \
    python bench.py --steps 1 --warmup 1 --pairs 2 --cpu-pairs 0 > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail -8
