#!/bin/bash
# 1-GPU visit: GPU test-suite, smoke, kbench (central + hybrid), the launch list of the bench command, and `ncu --set full`
# captures of the dominant kernels (the four fused stage kernels of one 512^3 step, plain flux_div and the hybrid stage at 256^3)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_pytest_1gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_pytest_1gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/kbench.py --lattice 8 8 8 --iters 20 2>&1 | grep -v Warning | tee $O/r02_kbench_central.log
timeout 300 python tools/kbench.py --lattice 8 8 8 --iters 10 --scheme hybrid 2>&1 | grep -v Warning | tee $O/r02_kbench_hybrid.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spb|flux|rk_|exchange|reduce|flag' -c 400 --csv \
    --log-file $O/r02_launches_bench_512cube.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > $O/r02_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flux_div_narrow --launch-skip 4 -c 4 -f -o $O/r02_ncu_full_stage_kernels_512cube \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-parity > $O/r02_ncu_stage.log 2>&1; echo "ncu stage rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_div_narrow --launch-skip 3 -c 1 -f -o $O/r02_ncu_full_rhs_256cube \
    python tools/kbench.py --lattice 8 8 8 --only 'flux_div[' --iters 2 > $O/r02_ncu_rhs.log 2>&1; echo "ncu rhs rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_div_kernel --launch-skip 3 -c 1 -f -o $O/r02_ncu_full_hybrid_stage_256cube \
    python tools/kbench.py --lattice 8 8 8 --scheme hybrid --only 'fused_stage[nin=1,out=1]' --iters 2 > $O/r02_ncu_hybrid.log 2>&1; echo "ncu hybrid rc=$?"
ls -la $O/*.ncu-rep
