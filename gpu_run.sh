#!/bin/bash
for v in A B C; do echo "variant $v"; SAHBVH_LIB=$PWD/cs201_sah-bvh_ray_tracer_b200/build/var_$v.so python tools/primary_bench.py 2>&1 | grep "mode 1"; done
