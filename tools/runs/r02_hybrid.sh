#!/bin/bash
# hybrid (WENO) kernel check: parity tests that touch the wide kernel, then per-kernel timings on identity and stretched grids
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_flux_div_gpu.py tests/test_curvilinear.py tests/test_channel_gpu.py tests/test_exchange_rk_gpu.py tests/test_mms.py tests/test_io.py tests/test_amr.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/kbench.py --lattice 8 8 8 --iters 10 --scheme hybrid --only 'flux_div[;fused_stage[' 2>&1 | grep -v Warning | tee $O/r02_kbench_hybrid_${1:-a}.log
timeout 300 python tools/kbench.py --lattice 8 8 8 --iters 10 --scheme hybrid --coords channel --only 'flux_div[;fused_stage[nin=1,out=1]' 2>&1 | grep -v Warning | tee $O/r02_kbench_hybrid_curv_${1:-a}.log
timeout 300 python tools/kbench.py --lattice 8 8 8 --iters 10 --scheme ck4 --only 'flux_div[;fused_stage[nin=1,out=1]' 2>&1 | grep -v Warning
