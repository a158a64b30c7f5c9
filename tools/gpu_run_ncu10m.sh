#!/bin/bash
# ncu evidence for the top-of-tree residency question (NS-1): the ray-buffer kernel on the 10 M-sphere tree (320 MB of
# quantised nodes > 126 MB L2), and the launch list of the default bench command
mkdir -p gpurun_out
timeout 600 python tools/bounce_bench.py --spheres 10000000 --modes 2 --steps 2 > gpurun_out/plain_bounce10m.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_rays_wide -s 1 -c 1 -f \
    -o gpurun_out/prof_wide_10M_r2 python tools/bounce_bench.py --spheres 10000000 --modes 2 --steps 2 > gpurun_out/ncu_wide10m.log 2>&1
echo "ncu wide 10M exit $?"; tail -2 gpurun_out/plain_bounce10m.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"; wc -l gpurun_out/launches_r2.csv
