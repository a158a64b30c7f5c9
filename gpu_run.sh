#!/bin/bash
for c in 1048576 2097152 4194304 8388608; do SAHBVH_HOST_CHUNK_RAYS=$c python tools/e2e_bench.py 2>&1 | grep chunk; done
