#!/bin/bash
# $1 = tag, $2.. = kernel regexes; one --set full capture (2 launches) each -> gpurun_out/full_<tag>_<i>.ncu-rep
TAG=$1; shift
mkdir -p gpurun_out
i=0
for K in "$@"; do
  ncu --set full --clock-control none --import-source on -k "regex:$K" -s 2 -c 2 -f -o gpurun_out/full_${TAG}_$i \
      python bench.py --steps 1 --warmup 1 --pairs 2 --cpu-pairs 0 > gpurun_out/ncu_full_${TAG}_$i.log 2>&1
  i=$((i+1))
done
ls -la gpurun_out | tail -5
