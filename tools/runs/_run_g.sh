timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python tools/kbench.py --lattice 16 16 16 --block 16 --only fused 2>&1 | tail -8
timeout 300 python tools/kbench.py --lattice 16 16 16 --block 16 --only 'flux_div[' 2>&1 | tail -1
timeout 300 python tools/kbench.py --lattice 8 8 8 --only 'ghosts[nin=1,out=1' 2>&1 | tail -1
