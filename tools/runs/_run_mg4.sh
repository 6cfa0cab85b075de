#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --lattice 8 8 8 > $O/bench_mg4_$name.json 2> $O/bench_mg4_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('$O/bench_mg4_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value']/1e9,2), round(d['ms_per_step'],4), round(d['roofline']['ms_per_launch'],4))
except Exception as e:
    print('$name failed', e); print(open('$O/bench_mg4_$name.err').read()[-600:])
PY
}
# (a variant with NCCL_P2P_USE_CUDA_MEMCPY=1 hung until the timeout on this NCCL build and was removed)
run base SPB_NCCL_HIGH_PRIO=1
run chan8 SPB_NCCL_HIGH_PRIO=1 NCCL_MIN_P2P_NCHANNELS=8 NCCL_MAX_P2P_NCHANNELS=8
run chan1 SPB_NCCL_HIGH_PRIO=1 NCCL_MIN_P2P_NCHANNELS=1 NCCL_MAX_P2P_NCHANNELS=1
run onestream SPB_NCCL_HIGH_PRIO=1 SPB_TWO_STREAMS=0
