#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
nvidia-smi -L
timeout 900 python -m pytest tests/test_shim_gpu.py tests/test_multigpu_nccl.py -m gpu -x -q 2>&1 | tail -15
(cd integration/_build && timeout 300 ./tgv_shim_demo 8 32 3 0 2) 2>&1 | tail -2 | tee $O/shim_demo_2gpu.log
(cd integration/_build && timeout 300 ./tgv_shim_demo 8 32 3 0 1) 2>&1 | tail -1 | tee -a $O/shim_demo_2gpu.log
