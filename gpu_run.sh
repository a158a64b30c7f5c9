#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "render_frame" > gpurun_out/pytest_render.log 2>&1; echo "pytest exit $?"
tail -30 gpurun_out/pytest_render.log
