timeout 1500 python -m pytest tests/test_exchange_rk_gpu.py tests/test_shim_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 300 python tools/kbench.py --lattice 8 8 8 --scheme hybrid --only fused 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_h.json').read().strip().splitlines()[-1])
print(d['value']/1e9, d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['roofline']['alg_bytes_per_cell'])
PY
tail -3 gpurun_out/bench_h.err
