#!/bin/bash
# 1-GPU visit: the bench lines of configs 2 (default, with the config-4 sub-record, e2e and CPU baseline), 5 and 3
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r02_bench_c2_n1.json 2> $O/r02_bench_c2_n1.err; echo "config 2 rc=$?"; tail -c 600 $O/r02_bench_c2_n1.err
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_bench_c5_n1.json 2> $O/r02_bench_c5_n1.err; echo "config 5 rc=$?"; tail -c 600 $O/r02_bench_c5_n1.err
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_bench_c3_n1.json 2> $O/r02_bench_c3_n1.err; echo "config 3 rc=$?"; tail -c 600 $O/r02_bench_c3_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_ref_n1.json 2> $O/r02_bench_ref_n1.err; echo "reference arm rc=$?"
python - <<PY
import json
for c in ("c2", "c5", "c3", "ref"):
    try:
        d = json.loads(open(f"$O/r02_bench_{c}_n1.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(c, "no line", e); continue
    r = d.get("roofline") or {}
    print(c, json.dumps({"value": d["value"], "ms_per_step": d["ms_per_step"], "frac": r.get("frac"), "bound": r.get("bound"), "stage_ms": r.get("ms_per_launch"),
          "share": r.get("step_share"), "rhs_only": (r.get("rhs_only") or {}).get("frac"), "e2e": (d.get("e2e") or {}).get("value"), "parity": d.get("parity_check"),
          "configs": d.get("configs"), "cpu": d.get("cpu_baseline"), "launches": d.get("gpu_launches")}))
PY
