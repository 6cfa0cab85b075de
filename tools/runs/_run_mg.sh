set -u
O=gpurun_out; mkdir -p $O
nvidia-smi -L
timeout 900 python -m pytest tests/test_multigpu_nccl.py -m gpu -x -q > $O/pytest_mg.log 2>&1; echo "pytest rc=$?"; tail -30 $O/pytest_mg.log
N=${1:-2}
for n in 1 $N; do
  if [ $n = 1 ]; then timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_mg_$n.json 2> $O/bench_mg_$n.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_mg_$n.json 2> $O/bench_mg_$n.err; fi
  echo "bench n=$n rc=$?"; tail -c 700 $O/bench_mg_$n.json; tail -3 $O/bench_mg_$n.err
done
