#!/bin/bash
# weak-scaling sweeps on one 8-GPU box: configs[1] shape (512^3 per GPU) and configs[3] (256^3 per GPU) at N = 1, 2, 4, 8
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for N in 1 2 4 8; do
  for L in "16 16 16" "8 8 8"; do
    tag=$(echo $L | tr ' ' 'x')
    if [ $N -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --lattice $L > $O/scale_${tag}_n$N.json 2> $O/scale_${tag}_n$N.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N \
        bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --lattice $L > $O/scale_${tag}_n$N.json 2> $O/scale_${tag}_n$N.err
    fi
    python - <<PY
import json
try:
    d=json.loads(open('$O/scale_${tag}_n$N.json').read().strip().splitlines()[-1]); print('$tag', $N, round(d['value']/1e9,2), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e: print('$tag', $N, 'failed', e)
PY
  done
done
