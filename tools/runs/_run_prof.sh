set -u
O=gpurun_out; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r01.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_under_ncu_r01.log 2>&1; echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:flux_div_narrow --launch-skip 12 -c 4 -f -o $O/ncu_r01_stage_kernels_512cube \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_stage_512.log 2>&1; echo "ncu full rc=$?"
ls -la $O | tail -8
