#!/bin/bash
# 8-GPU visit (b): where does the N = 8 step lose time? per-rank stage phases (SPB_PHASE_EVENTS), block parts vs block runs
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
one () { # name config env...
  name=$1; cfg=$2; shift; shift
  env "$@" SPB_PHASE_EVENTS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --config $cfg --steps 10 --warmup 3 --no-e2e --no-configs \
     > $O/r02_e8_$name.json 2> $O/r02_e8_$name.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/r02_e8_$name.json") if l.startswith("{")][-1]); r = d["roofline"]
    print("$name", json.dumps({"value": d["value"], "ms_per_step": d["ms_per_step"], "stage_ms": r["ms_per_launch"], "share": r["step_share"]}))
    for i, p in enumerate(d["phases"]["per_rank"]): print("   rank", i, {k: round(v, 3) for k, v in p.items()})
except Exception as e:
    print("$name no line:", e); print(open("$O/r02_e8_$name.err").read()[-800:])
PY
}
one c2_defer 2 SPB_DEFER_UNPACK=1
one c2_parts 2 SPB_DEFER_UNPACK=0
one c2_runs 2 SPB_BLOCK_RUNS=1
one c4_defer 4 SPB_DEFER_UNPACK=1
one c4_parts 4 SPB_DEFER_UNPACK=0
