#!/bin/bash
# 1-GPU visit: the driver's sequence — GPU test-suite, smoke, reference arm, default bench line (with e2e, baselines, configs)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_pytest_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_reference_final.json 2> $O/r02_bench_reference_final.err ) 2>&1 | grep real
( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench_final.json 2> $O/r02_bench_final.err ) 2>&1 | grep real
tail -c 400 $O/r02_bench_final.err
python - <<PY
import json
d = json.loads(open("$O/r02_bench_final.json").read().strip().splitlines()[-1])
r = d["roofline"]
print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}))
print(json.dumps({k: r[k] for k in ("bound", "achieved", "peak", "frac", "traffic", "step_share", "rhs_only")}))
print(json.dumps(d["e2e"])); print(json.dumps(d["cpu_baseline"])); print(json.dumps(d["parity_check"])); print(json.dumps(d["configs"]))
print(json.dumps({k: {kk: v.get(kk) for kk in ("tag", "value", "b200_value", "speedup_over_best_valid")} for k, v in d["ref_gpu_baseline"].items()}))
r = json.loads(open("$O/r02_bench_reference_final.json").read().strip().splitlines()[-1]); print("reference arm", r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["flags"])
PY
