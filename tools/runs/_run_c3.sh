#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_curvilinear.py tests/test_shim_gpu.py -m gpu -q 2>&1 | tail -5
(cd integration/_build && ./channel_curv_demo 4 32 3) 2>&1 | tail -2 | tee $O/curv_demo.log
for c in identity channel; do for s in central hybrid ck4; do
  echo "== coords=$c scheme=$s"; timeout 300 python tools/kbench.py --lattice 8 8 8 --scheme $s --coords $c --only 'flux_div[' 2>&1 | tail -1
  timeout 300 python tools/kbench.py --lattice 8 8 8 --scheme $s --coords $c --only 'fused_stage[nin=1,out=1' 2>&1 | tail -1
done; done | tee $O/kbench_curv.log
