#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/bounce_bench.py --modes 2 --rays 16777216 --steps 1"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_trace_rays_wide -s 1 -c 1 -f -o gpurun_out/prof_wide_r1 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -1 gpurun_out/plain.log
