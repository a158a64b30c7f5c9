#!/bin/bash
# usage: gpurun --timeout 2400 -- bash tools/gpu_run_scratch.sh     (GPU test suite + default bench, outputs under gpurun_out/)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
