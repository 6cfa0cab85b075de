#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_div_narrow --launch-skip 3 -c 1 -f -o $O/ncu_r01_narrow_stage_ghosts_256cube \
    python tools/kbench.py --lattice 8 8 8 --only 'fused_stage+ghosts[nin=1,out=1' --iters 2 > $O/ncu_nrw.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_div_narrow --launch-skip 3 -c 1 -f -o $O/ncu_r01_narrow_rhs_256cube \
    python tools/kbench.py --lattice 8 8 8 --only 'flux_div[' --iters 2 > $O/ncu_nrw2.log 2>&1; echo "rc=$?"
