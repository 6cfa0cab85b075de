#!/bin/bash
# verification visit: GPU parity suite + headline bench
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_v.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu_v.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_v.json 2> $O/bench_v.err; echo "bench rc=$?"
tail -c 3000 $O/bench_v.json
timeout 600 python tools/kbench.py --lattice 8 8 8 > $O/kbench_v.log 2>&1; cat $O/kbench_v.log
timeout 600 python tools/kbench.py --lattice 8 8 8 --scheme hybrid > $O/kbench_v_hybrid.log 2>&1; cat $O/kbench_v_hybrid.log
