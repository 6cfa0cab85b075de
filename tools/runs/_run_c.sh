set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_exchange_rk_gpu.py -m gpu -x -q > $O/pytest_gpu_c.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu_c.log
timeout 600 python tools/kbench.py --lattice 8 8 8 --only fused > $O/kbench_c.log 2>&1; cat $O/kbench_c.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_div_narrow --launch-skip 3 -c 1 -f -o $O/ncu_fusedghost_c \
    python tools/kbench.py --lattice 8 8 8 --only 'ghosts[nin=1,out=1' --iters 2 > $O/ncu_fusedghost_c.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_c.json 2> $O/bench_c.err; echo "bench rc=$?"; tail -c 1500 $O/bench_c.json; tail -5 $O/bench_c.err
