#!/bin/bash
# One script for every GPU-box visit (replaces the 52 one-off scripts of rounds 1 and 2 under tools/runs/).
#   gpurun [--gpus N] --timeout T -- 'bash tools/gpu_visit.sh <task> [args]'
# Tasks (outputs go to gpurun_out/, TAG defaults to "a"):
#   tests [TAG]              GPU parity suite + smoke()
#   final [TAG]              the driver's sequence: suite, smoke, reference arm, default bench line (e2e, baselines, configs)
#   kbench SCHEME [COORDS]   per-kernel timings at 256^3 (central | hybrid | ck4 | euler; identity | channel)
#   hybrid [TAG]             wide-kernel parity subset + hybrid / stretched / cent_keep<4> timings
#   bench CONFIG N [FLAGS]   one bench line of config 2|3|4|5 on N GPUs (torchrun for N > 1) + a short digest
#   nccl WORLD               multi-rank parity worker on both exchange paths (peer memory, NCCL send/recv)
#   phases N CONFIG          per-rank stage phases, join and step times (SPB_PHASE_EVENTS=1; SPB_DEFER_UNPACK=0|1 from the env)
#   solo2                    two independent 1-GPU runs at the same time, then the coupled 2-GPU run (what does coupling cost?)
#   steps                    fixed overhead of the timed region: 1-GPU bench at 10 and 20 steps
#   shim N [ARGS]            C++ shim: parity against the reference's CUDA run (N <= 2) and integration/bench_shim on 1 and N GPUs
#   profile [TAG]            ncu launch list of the bench command + `ncu --set full` of the stage / rhs / hybrid kernels
#   exp                      timing-only SPB_EXP builds of the narrow kernel (make -C spade_b200/csrc exp EXP=n first)
#   sanitizer                compute-sanitizer memcheck + racecheck over the kernel set (small grids)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
task=${1:-tests}; shift || true
TORCHRUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"

digest () {   # one-line digest of a bench JSON file
  python - "$1" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); r = d["roofline"]
    out = {"value": d["value"], "ms_per_step": round(d["ms_per_step"], 3), "frac": round(r["frac"], 4), "stage_ms": r.get("ms_per_launch"), "share": r.get("step_share"),
           "rhs_only": (r.get("rhs_only") or {}).get("frac"), "launches": d["gpu_launches"], "e2e": (d.get("e2e") or {}).get("value"), "clocks": d.get("clocks"),
           "parity": {k: (d.get("parity_check") or {}).get(k) for k in ("ok", "exchange_bit_exact", "trajectory_rel_l2", "error")},
           "steps_detail": d.get("steps_detail"), "cfg4": ((d.get("configs") or {}).get("config4") or {}).get("value")}
    print(json.dumps(out))
    for p in ((d.get("phases") or {}).get("per_rank") or ([d["phases"]] if d.get("phases") else [])):
        print("   ", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in p.items()})
except Exception as e:
    print("no bench line:", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
}
bench_line () {   # config gpus tag flags...
  local c=$1 n=$2 tag=$3; shift 3
  if [ "$n" -gt 1 ]; then
    timeout 900 $TORCHRUN --nproc-per-node=$n --master-port 29512 bench.py --gpus $n --config $c "$@" > $O/bench_c${c}_n${n}_$tag.json 2> $O/bench_c${c}_n${n}_$tag.err
  else
    timeout 900 python bench.py --gpus 1 --config $c "$@" > $O/bench_c${c}_n${n}_$tag.json 2> $O/bench_c${c}_n${n}_$tag.err
  fi
  echo "bench config $c N=$n rc=$?"; digest $O/bench_c${c}_n${n}_$tag.json
}

case $task in
tests)
  TAG=${1:-a}
  timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu_$TAG.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ;;
final)
  TAG=${1:-a}
  timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_$TAG.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
  ( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference_$TAG.json 2> $O/bench_reference_$TAG.err ) 2>&1 | grep real
  ( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_final_$TAG.json 2> $O/bench_final_$TAG.err ) 2>&1 | grep real
  digest $O/bench_final_$TAG.json
  python - <<PY
import json
d = json.loads(open("$O/bench_final_$TAG.json").read().strip().splitlines()[-1])
print(json.dumps(d["e2e"])); print(json.dumps(d["cpu_baseline"])); print(json.dumps(d["configs"]))
print(json.dumps({k: {kk: v.get(kk) for kk in ("tag", "value", "b200_value", "speedup_over_best_valid")} for k, v in (d.get("ref_gpu_baseline") or {}).items() if isinstance(v, dict)}))
r = json.loads(open("$O/bench_reference_$TAG.json").read().strip().splitlines()[-1]); print("reference arm", r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"].get("flags"))
PY
  ;;
kbench)
  timeout 300 python tools/kbench.py --lattice 8 8 8 --iters 20 --scheme ${1:-central} --coords ${2:-identity} 2>&1 | grep -v Warning | tee $O/kbench_${1:-central}_${2:-identity}.log ;;
hybrid)
  TAG=${1:-a}
  timeout 900 python -m pytest tests/test_flux_div_gpu.py tests/test_curvilinear.py tests/test_channel_gpu.py tests/test_exchange_rk_gpu.py tests/test_mms.py tests/test_io.py tests/test_amr.py -m gpu -x -q 2>&1 | tail -4
  timeout 300 python tools/kbench.py --lattice 8 8 8 --iters 10 --scheme hybrid --only 'flux_div[;fused_stage[' 2>&1 | grep -v Warning | tee $O/kbench_hybrid_$TAG.log
  timeout 300 python tools/kbench.py --lattice 8 8 8 --iters 10 --scheme hybrid --coords channel --only 'flux_div[;fused_stage[nin=1,out=1]' 2>&1 | grep -v Warning | tee $O/kbench_hybrid_curv_$TAG.log
  timeout 300 python tools/kbench.py --lattice 8 8 8 --iters 10 --scheme ck4 --only 'flux_div[;fused_stage[nin=1,out=1]' 2>&1 | grep -v Warning ;;
bench)
  c=${1:-2}; n=${2:-1}; shift 2 || true
  bench_line $c $n a --steps 20 --warmup 5 "$@" ;;
nccl)
  w=${1:-2}
  for p2p in 1 0; do
    SPB_P2P=$p2p timeout 280 $TORCHRUN --nproc-per-node=$w --master-port 29511 tests/_nccl_worker.py > $O/nccl_worker_world${w}_p2p$p2p.log 2>&1
    echo "nccl worker world=$w p2p=$p2p rc=$? ok=$(grep -c 'ok p2p' $O/nccl_worker_world${w}_p2p$p2p.log)"; grep -E "Error|rel L2" $O/nccl_worker_world${w}_p2p$p2p.log | head -4
  done ;;
phases)
  n=${1:-2}; c=${2:-2}
  SPB_PHASE_EVENTS=1 bench_line $c $n phases --steps 10 --warmup 3 --no-e2e --no-configs --no-parity ;;
solo2)
  for g in 0 1; do
    CUDA_VISIBLE_DEVICES=$g timeout 300 python bench.py --gpus 1 --config 2 --steps 10 --warmup 3 --no-e2e --no-configs --no-parity --no-cpu-baseline > $O/solo2_gpu$g.json 2> $O/solo2_gpu$g.err &
  done
  wait
  for g in 0 1; do echo "solo gpu $g"; digest $O/solo2_gpu$g.json; done
  SPB_PHASE_EVENTS=1 bench_line 2 2 pair --steps 10 --warmup 3 --no-e2e --no-configs --no-parity ;;
steps)
  for st in 10 20; do bench_line 2 1 steps$st --steps $st --warmup 3 --no-e2e --no-configs --no-parity --no-cpu-baseline; done ;;
shim)
  n=${1:-1}; shift || true
  [ "$n" -le 2 ] && timeout 600 python -m pytest tests/test_shim_gpu.py -x -q 2>&1 | tail -3
  for g in 1 $n; do timeout 300 integration/_build/bench_shim $g 20 "$@" 2>&1 | tail -1 | tee $O/bench_shim_n$g.json | cut -c1-420; done ;;
profile)
  TAG=${1:-a}
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spb|flux|rk_|exchange|reduce|flag' -c 400 --csv \
      --log-file $O/launches_bench_512cube_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > $O/bench_under_ncu_$TAG.log 2>&1; echo "launch list rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:flux_div_narrow --launch-skip 4 -c 4 -f -o $O/ncu_full_stage_kernels_512cube_$TAG \
      python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-parity > $O/ncu_stage_$TAG.log 2>&1; echo "ncu stage rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_div_narrow --launch-skip 3 -c 1 -f -o $O/ncu_full_rhs_256cube_$TAG \
      python tools/kbench.py --lattice 8 8 8 --only 'flux_div[' --iters 2 > $O/ncu_rhs_$TAG.log 2>&1; echo "ncu rhs rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_div_kernel --launch-skip 3 -c 1 -f -o $O/ncu_full_hybrid_stage_256cube_$TAG \
      python tools/kbench.py --lattice 8 8 8 --scheme hybrid --only 'fused_stage[nin=1,out=1]' --iters 2 > $O/ncu_hybrid_$TAG.log 2>&1; echo "ncu hybrid rc=$?"
  ls -la $O/*_$TAG.ncu-rep ;;
exp)
  for e in "" _exp1 _exp2 _exp3 _exp6 _exp7; do
    [ -f spade_b200/libspade_b200$e.so ] || continue
    echo "== libspade_b200$e.so"
    SPB_B200_LIB=$PWD/spade_b200/libspade_b200$e.so timeout 300 python tools/kbench.py --lattice 8 8 8 --iters 20 \
        --only 'flux_div[;fused_stage[nin=1,out=1];fused_stage+ghosts[nin=1,out=1]' 2>&1 | grep -v Warning
  done | tee $O/exp.log ;;
sanitizer)
  for tool in ${SAN_TOOLS:-memcheck racecheck}; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitizer_case.py > $O/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|hazard|Invalid" $O/sanitizer_$tool.log | head -30
  done ;;
*)
  echo "unknown task $task"; sed -n 2,20p "$0"; exit 2 ;;
esac
