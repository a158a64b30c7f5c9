#!/bin/bash
# build_variant.sh NAME "-DMACRO=.. ..": kernel-tuning build of the library with extra nvcc flags, written to
# cs201_sah-bvh_ray_tracer_b200/build/variants/libsahbvh_NAME.so (select it at run time with SAHBVH_LIB=<path>)
set -e
name=$1; flags=$2
pkg=$(cd "$(dirname "$0")/../cs201_sah-bvh_ray_tracer_b200" && pwd)
root=$(dirname "$pkg")
mkdir -p "$pkg/build/variants/obj_$name"
make -C "$pkg" >/dev/null
for f in sahbvh_kernels gpu_builder render multi_gpu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
    -Xcompiler -fPIC,-ffp-contract=off,-Wall -I"$root/include" -I"$pkg/csrc" $flags -c "$pkg/csrc/$f.cu" -o "$pkg/build/variants/obj_$name/$f.o" &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$pkg/build/variants/libsahbvh_$name.so" \
  "$pkg/build/host_ref_api.o" "$pkg/build/sah_builder.o" "$pkg/build/flatten.o" "$pkg/build/wide.o" \
  "$pkg/build/variants/obj_$name/sahbvh_kernels.o" "$pkg/build/variants/obj_$name/gpu_builder.o" "$pkg/build/variants/obj_$name/render.o" "$pkg/build/variants/obj_$name/multi_gpu.o" -lm -lpthread
echo "built $pkg/build/variants/libsahbvh_$name.so"
