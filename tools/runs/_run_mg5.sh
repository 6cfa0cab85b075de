#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 240 python -m pytest tests/test_multigpu_nccl.py -m gpu -x -q 2>&1 | tail -6
run() { name=$1; lat=$2; shift; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --lattice $lat > $O/bench_mg5_$name.json 2> $O/bench_mg5_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('$O/bench_mg5_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value']/1e9,2), round(d['ms_per_step'],4), round(d['roofline']['ms_per_launch'],4))
except Exception as e:
    print('$name failed', e); print(open('$O/bench_mg5_$name.err').read()[-800:])
PY
}
run p2p_256 "8 8 8" SPB_P2P=1
run nccl_256 "8 8 8" SPB_P2P=0
run p2p_512 "16 16 16" SPB_P2P=1
run nccl_512 "16 16 16" SPB_P2P=0
