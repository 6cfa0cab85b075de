#!/bin/bash
# round-end verification on one B200: parity suite, smoke, headline bench, ncu launch list of the bench command
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc=$?"; tail -c 2600 $O/bench_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spb|flux|rk_|exchange|reduce|flag' -c 200 --csv \
    --log-file $O/launches_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_under_ncu_final.log 2>&1; echo "ncu rc=$?"
tail -3 $O/launches_final.csv | cut -c1-300
