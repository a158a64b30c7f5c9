#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/primary_bench.py > gpurun_out/plain_macro.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_primary_macro -s 3 -c 1 -f \
    -o gpurun_out/prof_macro python tools/primary_bench.py > gpurun_out/ncu_macro.log 2>&1
echo "ncu macro exit $?"; tail -3 gpurun_out/plain_macro.log
