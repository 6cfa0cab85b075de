#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_multigpu_nccl.py -m gpu -x -q 2>&1 | tail -3
for ts in 1 0; do
SPB_TWO_STREAMS=$ts timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$ts bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_mg2_ts$ts.json 2> $O/bench_mg2_ts$ts.err
echo "two_streams=$ts rc=$?"; python - <<PY
import json
d=json.loads(open('$O/bench_mg2_ts$ts.json').read().strip().splitlines()[-1]); print(d['value']/1e9, d['ms_per_step'], d['roofline']['ms_per_launch'], d['clocks'])
PY
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1', d['value']/1e9, d['ms_per_step'], d['clocks'])"
