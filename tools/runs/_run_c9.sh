#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_exchange_rk_gpu.py -m gpu -x -q -k "wide_fused or fills_same_rank" 2>&1 | tail -3
timeout 200 python tools/config3.py --lattice 32 16 2 --steps 5 --warmup 2 2>&1 | tail -1 | cut -c280-640 | tee $O/config3_fusedghost_n1.txt
