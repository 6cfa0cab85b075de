#!/bin/bash
# 8-GPU visit: multi-rank parity worker at world 4 and 8 on both exchange paths, the bench configs at N = 8 (and config 5 at 4),
# the C++ shim bench at 1 and 8 GPUs
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for w in 4 8; do for p2p in 1 0; do
  SPB_P2P=$p2p timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$w --master-addr 127.0.0.1 --master-port 29511 \
     tests/_nccl_worker.py > $O/r02_nccl_w${w}_p2p$p2p.log 2>&1
  echo "nccl worker world=$w p2p=$p2p rc=$? ok=$(grep -c 'ok p2p' $O/r02_nccl_w${w}_p2p$p2p.log)"; grep -E "Error|rel L2" $O/r02_nccl_w${w}_p2p$p2p.log | head -4
done; done
run_bench () {  # config gpus extra-flags
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $2 --config $1 --steps 10 --warmup 3 $3 \
     > $O/r02_bench_c$1_n$2.json 2> $O/r02_bench_c$1_n$2.err; echo "bench config $1 N=$2 rc=$?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/r02_bench_c$1_n$2.json") if l.startswith("{")][-1])
    r = d["roofline"]
    print(json.dumps({"value": d["value"], "ms_per_step": d["ms_per_step"], "frac": r["frac"], "share": r["step_share"], "launches": d["gpu_launches"], "e2e": (d.get("e2e") or {}).get("value"),
          "parity": {k: d["parity_check"].get(k) for k in ("ok", "exchange_bit_exact", "trajectory_rel_l2", "error")}, "cfg4": ((d.get("configs") or {}).get("config4") or {}).get("value")}))
except Exception as e:
    print("no line:", e); print(open("$O/r02_bench_c$1_n$2.err").read()[-1500:])
PY
}
run_bench 4 8 ""
run_bench 2 8 "--no-e2e"
run_bench 5 8 "--no-e2e"
run_bench 5 4 "--no-e2e"
run_bench 3 8 "--no-e2e"
for g in 1 8; do timeout 300 integration/_build/bench_shim $g 20 2>&1 | tail -1 | tee $O/r02_bench_shim_n$g.json | cut -c1-420; done
