#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python tools/_hostcost.py 2>&1 | tail -4
for hp in 1 0; do
SPB_NCCL_HIGH_PRIO=$hp timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$hp bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --lattice 8 8 8 > $O/bench_mg3_hp$hp.json 2> $O/bench_mg3_hp$hp.err
echo "high_prio=$hp rc=$?"; python - <<PY
import json
d=json.loads(open('$O/bench_mg3_hp$hp.json').read().strip().splitlines()[-1]); print(d['value']/1e9, d['ms_per_step'], d['roofline']['ms_per_launch'])
PY
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --lattice 8 8 8 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1', d['value']/1e9, d['ms_per_step'])"
