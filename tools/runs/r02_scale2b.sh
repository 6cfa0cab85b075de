#!/bin/bash
# 2-GPU visit: config 2 (512^3 per GPU) with per-phase events incl. every wait / unpack of finish(); deferred unpack A/B
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
one () { # name cfg env...
  name=$1; cfg=$2; shift; shift
  env "$@" SPB_PHASE_EVENTS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config $cfg --steps 10 --warmup 3 --no-e2e --no-configs \
     > $O/r02_s2b_$name.json 2> $O/r02_s2b_$name.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/r02_s2b_$name.json") if l.startswith("{")][-1]); r = d["roofline"]
    print("$name", json.dumps({"value": d["value"], "ms_per_step": d["ms_per_step"], "stage_ms": r["ms_per_launch"], "share": r["step_share"], "parity": d["parity_check"].get("ok"), "err": d["parity_check"].get("error")}))
    for i, p in enumerate(d["phases"]["per_rank"]): print("   rank", i, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in p.items()})
except Exception as e:
    print("$name no line:", e); print(open("$O/r02_s2b_$name.err").read()[-1200:])
PY
}
one c2_nodefer 2 SPB_DEFER_UNPACK=0
one c2_defer 2 SPB_DEFER_UNPACK=1
one c4_nodefer 4 SPB_DEFER_UNPACK=0
one c4_defer 4 SPB_DEFER_UNPACK=1
SPB_P2P=1 timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 tests/_nccl_worker.py 2>&1 | grep -E "ok p2p|Error|rel L2" | head
SPB_P2P=0 timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 tests/_nccl_worker.py 2>&1 | grep -E "ok p2p|Error|rel L2" | head
