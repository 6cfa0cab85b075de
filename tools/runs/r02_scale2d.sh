#!/bin/bash
# 2-GPU diagnosis (d): what does the coupling cost? (1) two INDEPENDENT 1-GPU runs of config 2 at the same time (each GPU's own pace
# under load), (2) the coupled 2-GPU run with the deferred schedule and the join / step events of every rank
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for g in 0 1; do
  CUDA_VISIBLE_DEVICES=$g timeout 300 python bench.py --gpus 1 --config 2 --steps 10 --warmup 3 --no-e2e --no-configs --no-parity --no-cpu-baseline > $O/r02_s2d_solo$g.json 2> $O/r02_s2d_solo$g.err &
done
wait
SPB_PHASE_EVENTS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 2 --steps 10 --warmup 3 --no-e2e --no-configs --no-parity > $O/r02_s2d_pair.json 2> $O/r02_s2d_pair.err
python - <<PY
import json
for g in (0, 1):
    d = json.loads([l for l in open("$O/r02_s2d_solo%d.json" % g) if l.startswith("{")][-1])
    print("solo gpu", g, "ms_per_step", round(d["ms_per_step"], 3), "stage", round(d["roofline"]["ms_per_launch"], 3), d["clocks"])
d = json.loads([l for l in open("$O/r02_s2d_pair.json") if l.startswith("{")][-1])
print("pair ms_per_step", round(d["ms_per_step"], 3))
for p in d["phases"]["per_rank"]:
    print("  ", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in p.items()})
PY
