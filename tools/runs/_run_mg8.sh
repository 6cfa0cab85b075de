set -u
O=gpurun_out; mkdir -p $O
N=${1:-8}
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_multigpu_nccl.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_mg_$N.json 2> $O/bench_mg_$N.err
echo "bench n=$N rc=$?"; tail -c 900 $O/bench_mg_$N.json; tail -3 $O/bench_mg_$N.err
