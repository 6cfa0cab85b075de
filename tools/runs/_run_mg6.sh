#!/bin/bash
# the driver's scaling command at N = 2, complete (e2e leg included)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2_full.json 2> $O/bench_n2_full.err; echo "rc=$?"
tail -c 1500 $O/bench_n2_full.json; tail -3 $O/bench_n2_full.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
