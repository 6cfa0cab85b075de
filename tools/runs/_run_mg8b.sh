#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_multigpu_nccl.py -m gpu -x -q -k "4-1" 2>&1 | tail -4
for L in "8 8 8" "16 16 16"; do
  tag=$(echo $L | tr ' ' 'x')
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 \
     bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --lattice $L > $O/scale_p2p_${tag}_n8.json 2> $O/scale_p2p_${tag}_n8.err
  timeout 100 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --lattice $L > $O/scale_p2p_${tag}_n1.json 2> $O/scale_p2p_${tag}_n1.err
  python - <<PY
import json
for n in (1, 8):
    try:
        d=json.loads(open('$O/scale_p2p_${tag}_n%d.json' % n).read().strip().splitlines()[-1]); print('$tag', n, round(d['value']/1e9,2), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])
    except Exception as e: print('$tag', n, 'failed', e)
PY
done
