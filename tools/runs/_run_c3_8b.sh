#!/bin/bash
# config 3 on 8 GPUs with the peer-memory exchange (and the 1-GPU share of the same grid for the efficiency)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
    tools/config3.py --lattice 32 16 2 --steps 5 --warmup 2 > $O/config3_p2p_n8.json 2> $O/config3_p2p_n8.err; echo "rc=$?"; tail -1 $O/config3_p2p_n8.json | cut -c1-900
timeout 100 python tools/config3.py --lattice 32 16 2 --steps 5 --warmup 2 > $O/config3_p2p_n1.json 2> $O/config3_p2p_n1.err; tail -1 $O/config3_p2p_n1.json | cut -c300-700
