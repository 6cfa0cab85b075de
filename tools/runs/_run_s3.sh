#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_shim_gpu.py -m gpu -x -q 2>&1 | tail -6
(cd integration/_build && timeout 200 ./channel_curv_demo 4 32 3) 2>&1 | tail -2 | tee $O/curv_channel_demo.log
