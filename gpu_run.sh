#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
CMD="python tools/bounce_bench.py --modes 2 --rays 16777216 --steps 1"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_trace_rays_wide -s 1 -c 1 -f -o gpurun_out/prof_wide_r1 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -1 gpurun_out/plain.log
