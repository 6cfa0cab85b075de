#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_curvilinear.py -m gpu -q > $O/pytest_curv.log 2>&1; echo "curv rc=$?"; tail -40 $O/pytest_curv.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_curvilinear.py > $O/pytest_gpu_c2.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu_c2.log
