#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_c6.log 2>&1; echo "pytest rc=$?"; tail -30 $O/pytest_gpu_c6.log
