#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_c5.log 2>&1; echo "pytest rc=$?"; tail -30 $O/pytest_gpu_c5.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e9, d['roofline']['frac'], d['clocks'])"
