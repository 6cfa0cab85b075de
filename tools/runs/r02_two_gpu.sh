#!/bin/bash
# 2-GPU visit: multi-rank parity worker on both exchange paths, the C++ shim on 2 GPUs, bench_shim 1 -> 2 GPUs, bench configs at N = 2
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for p2p in 1 0; do
  SPB_P2P=$p2p timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 \
     tests/_nccl_worker.py > $O/r02_nccl_w2_p2p$p2p.log 2>&1
  echo "nccl worker p2p=$p2p rc=$?"; grep -E "ok p2p|Error|rel L2" $O/r02_nccl_w2_p2p$p2p.log | head -6
done
timeout 600 python -m pytest tests/test_shim_gpu.py -x -q 2>&1 | tail -3
for g in 1 2; do timeout 300 integration/_build/bench_shim $g 20 2>&1 | tail -1 | tee $O/r02_bench_shim_n$g.json; done
timeout 300 integration/_build/bench_shim 2 10 8 8 32 1 2>&1 | tail -1 | tee $O/r02_bench_shim_hybrid_n2.json
for c in 2 4 5 3; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config $c --steps 10 --warmup 3 \
     > $O/r02_bench_c${c}_n2.json 2> $O/r02_bench_c${c}_n2.err; echo "bench config $c N=2 rc=$?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/r02_bench_c${c}_n2.json") if l.startswith("{")][-1])
    r = d["roofline"]
    print(json.dumps({"value": d["value"], "ms_per_step": d["ms_per_step"], "frac": r["frac"], "share": r["step_share"], "e2e": (d.get("e2e") or {}).get("value"), "parity": {k: d["parity_check"].get(k) for k in ("ok", "exchange_bit_exact", "trajectory_rel_l2", "path", "error")}}))
except Exception as e:
    print("no line:", e); print(open("$O/r02_bench_c${c}_n2.err").read()[-1500:])
PY
done
