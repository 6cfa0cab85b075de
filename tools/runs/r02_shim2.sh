#!/bin/bash
# 2-GPU visit: the C++ shim with the deferred-unpack schedule (parity against the reference's 2-GPU CUDA run, bench_shim 1 -> 2)
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_shim_gpu.py tests/test_multigpu_nccl.py -x -q 2>&1 | tail -3
for g in 1 2; do timeout 300 integration/_build/bench_shim $g 20 2>&1 | tail -1 | cut -c1-330; done
timeout 300 integration/_build/bench_shim 2 10 16 16 2>&1 | tail -1 | cut -c1-330
