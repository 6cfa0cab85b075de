#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_c4.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu_c4.log
timeout 600 python tools/config3.py --lattice 32 16 2 --steps 5 --warmup 2 2>&1 | tail -3 | tee $O/config3_n1.json
timeout 600 python tools/config3.py --lattice 32 16 2 --steps 5 --warmup 2 --coords identity 2>&1 | tail -1 | tee $O/config3_n1_identity.json
