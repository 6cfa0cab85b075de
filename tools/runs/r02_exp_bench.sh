#!/bin/bash
# sustained (power-capped) regime: the 512^3 bench with the timing-only experiment builds
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for e in "" _exp2 _exp6; do
  echo "== libspade_b200$e.so"
  SPB_B200_LIB=$PWD/spade_b200/libspade_b200$e.so timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(json.dumps({'ms_per_step': d['ms_per_step'], 'stage_ms': r['ms_per_launch'], 'frac': r['frac'], 'rhs_only_ms': r['rhs_only']['ms_per_launch'], 'clocks': d['clocks']}))"
done | tee $O/r02_exp_bench.log
