#!/bin/bash
# timing-only experiment builds of the narrow kernel (SPB_EXP bits: 1 no barrier (2), 2 no flux hand-off, 4 no published differences)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for e in "" _exp1 _exp2 _exp3 _exp6 _exp7; do
  echo "== libspade_b200$e.so"
  SPB_B200_LIB=$PWD/spade_b200/libspade_b200$e.so timeout 300 python tools/kbench.py --lattice 8 8 8 --iters 20 \
      --only 'flux_div[;fused_stage[nin=1,out=1];fused_stage+ghosts[nin=1,out=1]' 2>&1 | grep -v Warning
done | tee $O/r02_exp.log
