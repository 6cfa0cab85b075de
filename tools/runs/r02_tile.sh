#!/bin/bash
# 32x16 tile experiment of the narrow kernel (SPB_TILE_32x16=1): parity subset, kbench, sustained bench
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
SPB_TILE_32x16=1 timeout 600 python -m pytest tests/test_flux_div_gpu.py tests/test_exchange_rk_gpu.py tests/test_curvilinear.py -m gpu -x -q 2>&1 | tail -3
for t in "" 1; do
  echo "== SPB_TILE_32x16=$t"
  if [ -n "$t" ]; then export SPB_TILE_32x16=1; else unset SPB_TILE_32x16; fi
  timeout 300 python tools/kbench.py --lattice 8 8 8 --iters 20 --only 'flux_div[;fused_stage[nin=1,out=1];fused_stage+ghosts[nin=1,out=1];fused_stage+ghosts[nin=2,out=1]' 2>&1 | grep -v Warning
  timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-configs --no-parity 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(json.dumps({'ms_per_step': d['ms_per_step'], 'stage_ms': r['ms_per_launch'], 'frac': r['frac'], 'rhs_only': r['rhs_only']['frac'], 'clocks': d['clocks']['sm_mhz']}))"
done | tee $O/r02_tile.log
