#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/pin_bench.py 2>&1 | tee gpurun_out/pin_bench.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "host_pinning or packet or 100k" 2>&1 | tail -3
