#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peer_frame" 2>&1 | tail -15
