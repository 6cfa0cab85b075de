set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_b.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_gpu_b.log
timeout 600 python tools/kbench.py --lattice 8 8 8 > $O/kbench_b.log 2>&1; cat $O/kbench_b.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_b.json 2> $O/bench_b.err; echo "bench rc=$?"; tail -c 1800 $O/bench_b.json; tail -5 $O/bench_b.err
