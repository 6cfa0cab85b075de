#!/bin/bash
# One GPU-box visit: parity tests, the headline bench, the ncu launch list of the bench command and full captures of
# the dominant kernels. Everything lands in gpurun_out/ (scratch); summaries are copied into profiles/ by hand.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TAG=${1:-a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi_$TAG.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
tail -c 2500 $O/bench_$TAG.json
timeout 600 python tools/kbench.py --lattice 8 8 8 > $O/kbench_$TAG.log 2>&1
cat $O/kbench_$TAG.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spb|flux|rk_|exchange|reduce' -c 400 --csv \
    --log-file $O/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_under_ncu_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_div_narrow --launch-skip 3 -c 1 -f -o $O/ncu_rhs_$TAG \
    python tools/kbench.py --lattice 8 8 8 --only 'flux_div[' --iters 2 > $O/ncu_rhs_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_div_narrow --launch-skip 3 -c 1 -f -o $O/ncu_fused_$TAG \
    python tools/kbench.py --lattice 8 8 8 --only 'nin=1,out=1' --iters 2 > $O/ncu_fused_$TAG.log 2>&1
ls -la $O
