#!/bin/bash
# One GPU-box visit: GPU parity tests, smoke, per-kernel timings (256^3, isolated) and the 512^3 bench (sustained, power-capped)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
TAG=${1:-a}
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02_pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/kbench.py --lattice 8 8 8 --iters 20 2>&1 | grep -v Warning | tee $O/r02_kbench_$TAG.log
timeout 600 python bench.py --steps 30 --warmup 5 ${BENCH_FLAGS:---no-e2e --no-cpu-baseline} > $O/r02_bench_$TAG.json 2> $O/r02_bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("$O/r02_bench_$TAG.json").read().strip().splitlines()[-1])
r = d["roofline"]
print(json.dumps({"ms_per_step": d["ms_per_step"], "value": d["value"], "stage_ms": r["ms_per_launch"], "frac": r["frac"], "rhs_only": r["rhs_only"], "clocks": d["clocks"]}))
PY
