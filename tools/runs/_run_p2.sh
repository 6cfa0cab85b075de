#!/bin/bash
# ncu --set full of the config-3 kernels (hybrid WENO/Ducros + viscous, fused stage): identity and general coordinates
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_div_kernel --launch-skip 3 -c 1 -f -o $O/ncu_r01_hybrid_stage_256cube \
    python tools/kbench.py --lattice 8 8 8 --scheme hybrid --only 'fused_stage[nin=1,out=1' --iters 2 > $O/ncu_hyb.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_div_kernel --launch-skip 3 -c 1 -f -o $O/ncu_r01_hybrid_stage_curv_256cube \
    python tools/kbench.py --lattice 8 8 8 --scheme hybrid --coords channel --only 'fused_stage[nin=1,out=1' --iters 2 > $O/ncu_hyb_curv.log 2>&1; echo "rc=$?"
ls -la $O | tail -5
