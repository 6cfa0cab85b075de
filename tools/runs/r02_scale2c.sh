#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for d in 1 0; do
SPB_DEFER_UNPACK=$d SPB_PHASE_EVENTS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 2 --steps 10 --warmup 3 --no-e2e --no-configs --no-parity > $O/r02_s2c_$d.json 2> $O/r02_s2c_$d.err
python - <<PY
import json
d = json.loads([l for l in open("$O/r02_s2c_$d.json") if l.startswith("{")][-1])
print("defer=$d", d["ms_per_step"], [ (round(p["host_loop_ms_per_step"],2), round(p["stage_kernel_ms"],3)) for p in d["phases"]["per_rank"]])
PY
done
