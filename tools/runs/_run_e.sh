O=gpurun_out
timeout 900 python -m pytest tests/test_flux_div_gpu.py tests/test_exchange_rk_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -5
for s in hybrid ck4; do timeout 300 python tools/kbench.py --lattice 8 8 8 --scheme $s --only flux_div 2>&1 | tail -2; done
