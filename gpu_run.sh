#!/bin/bash
python tools/primary_bench.py 2>&1 | grep "mode"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "primary or default_scene or tiny or rebuilt or packet or moved_camera" 2>&1 | tail -2
