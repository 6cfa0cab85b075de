#!/bin/bash
# 2-GPU visit: where does the multi-GPU overhead of the stage come from? config 4 (256^3 per GPU) at N = 1 and 2 with
# one / two streams, plus per-phase device timings (SPB_PHASE_EVENTS) of one stage
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
one () { # name env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 4 --steps 40 --warmup 5 --no-e2e \
     > $O/r02_s2_$name.json 2> $O/r02_s2_$name.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/r02_s2_$name.json") if l.startswith("{")][-1]); r = d["roofline"]
    print("$name", json.dumps({"value": d["value"], "ms_per_step": d["ms_per_step"], "stage_ms": r["ms_per_launch"], "share": r["step_share"], "launches": d["gpu_launches"], "phases": d.get("phases")}))
except Exception as e:
    print("$name no line:", e); print(open("$O/r02_s2_$name.err").read()[-800:])
PY
}
timeout 300 python bench.py --config 4 --steps 40 --warmup 5 --no-e2e --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1', d['value'], d['ms_per_step'])"
one default SPB_X=0
one onestream SPB_TWO_STREAMS=0
one nccl SPB_P2P=0
one phases SPB_PHASE_EVENTS=1
