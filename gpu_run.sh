#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
SAHBVH_BUILD_TIMING=1 python tools/build_bench.py 100000 10000000 > gpurun_out/build_timing.log 2>&1; grep -E "^\{" gpurun_out/build_timing.log; grep "gpu builder" gpurun_out/build_timing.log | tail -9
