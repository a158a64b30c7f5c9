#!/bin/bash
# compare library variants (build/variants/*.so + the in-tree library, selected through SAHBVH_LIB)
# usage: gpu_run_variants.sh "<harness command>"   e.g. "python tools/bounce_bench.py --modes 2"
mkdir -p gpurun_out
cmd=${1:-python tools/primary_bench.py}
: > gpurun_out/variants.log
for lib in cs201_sah-bvh_ray_tracer_b200/build/variants/*.so cs201_sah-bvh_ray_tracer_b200/libsahbvh_b200.so; do
  echo "== $lib" >> gpurun_out/variants.log
  SAHBVH_LIB=$PWD/$lib timeout 300 $cmd 2>&1 | grep -v "^build+upload" >> gpurun_out/variants.log
done
cat gpurun_out/variants.log
