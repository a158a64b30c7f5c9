#!/bin/bash
# compare library variants (build/variants/*.so, selected through SAHBVH_LIB) on the headline primary frame
mkdir -p gpurun_out
: > gpurun_out/variants.log
for lib in cs201_sah-bvh_ray_tracer_b200/build/variants/*.so cs201_sah-bvh_ray_tracer_b200/libsahbvh_b200.so; do
  echo "== $lib" >> gpurun_out/variants.log
  SAHBVH_LIB=$PWD/$lib timeout 300 python tools/primary_bench.py >> gpurun_out/variants.log 2>&1
done
cat gpurun_out/variants.log
timeout 900 python -m pytest tests -m gpu -q -x -k "packet or primary or fuzz or degenerate or kernels_agree" > gpurun_out/pytest_subset.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_subset.log
tail -5 gpurun_out/pytest_subset.log
