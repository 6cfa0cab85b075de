timeout 300 python tools/kbench.py --lattice 16 16 16 --block 16 --only fused 2>&1 | tail -8
timeout 300 python tools/kbench.py --lattice 16 16 16 --block 16 --only 'flux_div[' 2>&1 | tail -2
timeout 300 python tools/kbench.py --lattice 16 16 16 --block 16 --only exchange 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_div_kernel --launch-skip 3 -c 1 -f -o gpurun_out/ncu_hybrid_f \
    python tools/kbench.py --lattice 8 8 8 --scheme hybrid --only 'flux_div[' --iters 2 > gpurun_out/ncu_hybrid_f.log 2>&1
