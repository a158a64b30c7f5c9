#!/bin/bash
# packet-kernel build variants (SAHBVH_LIB) on the headline frame: tools/primary_bench.py per library
mkdir -p gpurun_out; : > gpurun_out/pktv.log
for lib in cs201_sah-bvh_ray_tracer_b200/libsahbvh_b200.so cs201_sah-bvh_ray_tracer_b200/build/variants/*.so; do
  echo "== $lib" >> gpurun_out/pktv.log
  SAHBVH_LIB=$PWD/$lib timeout 300 python tools/primary_bench.py 2>&1 | grep -v "mode 0" | tail -3 >> gpurun_out/pktv.log
done
cat gpurun_out/pktv.log
