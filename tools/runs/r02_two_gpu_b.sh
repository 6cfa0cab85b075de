#!/bin/bash
# 2-GPU visit (b): GPU test-suite on a 2-GPU box (multi-GPU tests run instead of skipping), nccl worker, config 5 at N = 1 / 2, bench_shim
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02_pytest_2gpu.log
timeout 300 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02_bench_c5_n1.json 2> $O/r02_bench_c5_n1.err; echo "config 5 N=1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 5 --steps 10 --warmup 3 --no-e2e \
     > $O/r02_bench_c5_n2.json 2> $O/r02_bench_c5_n2.err; echo "config 5 N=2 rc=$?"
python - <<PY
import json
for n in (1, 2):
    try:
        d = json.loads([l for l in open(f"$O/r02_bench_c5_n{n}.json") if l.startswith("{")][-1])
        r = d["roofline"]
        print(n, json.dumps({"value": d["value"], "ms_per_step": d["ms_per_step"], "frac": r["frac"], "stage_ms": r["ms_per_launch"], "share": r["step_share"], "launches": d["gpu_launches"], "parity": {k: d["parity_check"].get(k) for k in ("ok", "trajectory_rel_l2", "path", "error")}}))
    except Exception as e:
        print(n, "no line:", e); print(open(f"$O/r02_bench_c5_n{n}.err").read()[-1500:])
PY
for g in 1 2; do timeout 300 integration/_build/bench_shim $g 20 2>&1 | tail -1 | cut -c1-400; done
