#!/bin/bash
# ncu launch list of the default bench command (B200_PROFILING.md: per-launch times are cold-cache and serialised)
mkdir -p gpurun_out
# the kernel AUTO settles on in the plain run is pinned for the ncu pass (under ncu its event-timed samples are meaningless)
export SAHBVH_PRIMARY_MODE=${SAHBVH_PRIMARY_MODE:-2}
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/plain_bench.log 2>&1 || { echo "plain bench failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches_r2.csv
