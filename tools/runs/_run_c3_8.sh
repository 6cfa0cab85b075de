#!/bin/bash
# config 3 on 8 GPUs: 1024 x 512 x 512 cells, 8192 blocks of 32^3, stretched y, walls, hybrid + visc
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    tools/config3.py --lattice 32 16 2 --steps 5 --warmup 2 > $O/config3_n8.json 2> $O/config3_n8.err; echo "rc=$?"; tail -1 $O/config3_n8.json; tail -3 $O/config3_n8.err
timeout 600 python -m pytest tests/test_multigpu_nccl.py -m gpu -x -q 2>&1 | tail -3
