#!/bin/bash
# soak run of the differential fuzzer on the GPU (default cases + SAHBVH_FUZZ_EXTRA seeded ones)
mkdir -p gpurun_out
SAHBVH_FUZZ_EXTRA=${SAHBVH_FUZZ_EXTRA:-40} timeout 900 python -m pytest tests/test_fuzz.py -m gpu -q > gpurun_out/fuzz_gpu.log 2>&1
echo "fuzz exit $?" >> gpurun_out/fuzz_gpu.log
tail -25 gpurun_out/fuzz_gpu.log
