#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_curvilinear.py tests/test_channel_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_curvilinear.py --deselect tests/test_channel_gpu.py 2>&1 | tail -5
for s in central euler; do
  timeout 300 python tools/kbench.py --lattice 8 8 8 --scheme $s --coords channel 2>&1 | grep -v "rk_update\|reduce\|exchange\"" | tee -a $O/kbench_curv_narrow.log
done
timeout 300 python tools/kbench.py --lattice 8 8 8 --scheme central --only 'flux_div' 2>&1 | tail -2
