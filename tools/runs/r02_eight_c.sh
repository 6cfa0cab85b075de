#!/bin/bash
# 8-GPU visit (c): the deferred-unpack schedule at 8 ranks — parity worker, configs 2 and 4, C++ shim bench
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
SPB_P2P=1 timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29511 tests/_nccl_worker.py > $O/r02_nccl_w8_p2p1_defer.log 2>&1
echo "nccl worker world=8 p2p=1 rc=$? ok=$(grep -c 'ok p2p' $O/r02_nccl_w8_p2p1_defer.log)"; grep -E "Error|rel L2" $O/r02_nccl_w8_p2p1_defer.log | head -4
run_bench () {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --config $1 --steps 20 --warmup 5 --no-e2e $2 \
     > $O/r02_bench_c$1_n8_defer.json 2> $O/r02_bench_c$1_n8_defer.err; echo "bench config $1 N=8 rc=$?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/r02_bench_c$1_n8_defer.json") if l.startswith("{")][-1]); r = d["roofline"]
    print(json.dumps({"value": d["value"], "ms_per_step": d["ms_per_step"], "frac": r["frac"], "stage_ms": r["ms_per_launch"], "share": r["step_share"], "parity": d["parity_check"].get("ok"), "cfg4": ((d.get("configs") or {}).get("config4") or {}).get("value")}))
except Exception as e:
    print("no line:", e); print(open("$O/r02_bench_c$1_n8_defer.err").read()[-1200:])
PY
}
run_bench 2 ""
run_bench 4 "--no-configs"
timeout 300 integration/_build/bench_shim 8 20 2>&1 | tail -1 | tee $O/r02_bench_shim_n8_defer.json | cut -c1-330
timeout 300 integration/_build/bench_shim 1 20 2>&1 | tail -1 | tee $O/r02_bench_shim_n1_defer.json | cut -c1-330
