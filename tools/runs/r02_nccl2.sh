#!/bin/bash
# 2-GPU visit: the multi-rank parity worker on both exchange paths (peer memory / NCCL send-recv)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for p2p in 1 0; do
  SPB_P2P=$p2p timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 \
     tests/_nccl_worker.py > $O/r02_nccl_w2_p2p$p2p.log 2>&1
  echo "p2p=$p2p rc=$?"; grep -E "ok p2p|AssertionError|umax|rel L2" $O/r02_nccl_w2_p2p$p2p.log | head -8
done
