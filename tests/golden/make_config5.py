"""Generates tests/golden/config5_amr.npz from the UNMODIFIED reference (oracle/_ref, dev container only): BASELINE config 5,
an AMR-refined block grid (reference src/amr, development/amr-exchange/main.cc:30-51 pattern: refine the blocks that
intersect a sphere, constraint factor2), as the block boxes and SPADE's own exchange_config_t tables (injection +
interpolation lists) for 1, 2, 4 and 8 ranks of SPADE's contiguous partition (grid/partition.h:27-84).

Two grids:
  b_*  the bench grid: 8x8x8 root blocks of 32^3 cells on [0, 2 pi)^3, periodic; pass 1 refines the roots that intersect the
       sphere |x - c| < 0.30 * 2 pi, pass 2 the children that intersect |x - c| < 0.12 * 2 pi (bench.py --config 5)
  p_*  the parity grid of bench.py's guard and of tests/_nccl_worker.py: 2x2x2 roots of 32x8x8 cells, root 0 refined

    python tests/golden/make_config5.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref  # noqa: E402
from util import oracle_cfg  # noqa: E402

NG = 2
L = 2 * np.pi
RANKS = (1, 2, 4, 8)


def intersects(boxes, centre, radius):
    """blocks whose box comes within `radius` of `centre` (closest-point test)"""
    lo, hi = boxes[:, 0::2], boxes[:, 1::2]
    near = np.clip(centre, lo, hi)
    return np.flatnonzero(((near - centre) ** 2).sum(axis=1) < radius ** 2)


def tables(out, prefix, roots, n):
    cfg1 = oracle_cfg(roots, n, NG, periodic=(1, 1, 1), nranks=1)
    nblk, boxes = ref.block_boxes(cfg1, cap=1 << 15)
    out[f"{prefix}_boxes"] = boxes
    out[f"{prefix}_roots"] = np.array(roots)
    out[f"{prefix}_cells"] = np.array(n)
    for nranks in RANKS:
        cfg = oracle_cfg(roots, n, NG, periodic=(1, 1, 1), nranks=nranks)
        for rank in range(nranks):
            s_, r_, _ = ref.exchange_tables(cfg, rank, cap=1 << 19)
            si, ri, _ = ref.interp_tables(cfg, rank, cap=1 << 19)
            for nm, arr in (("send", s_), ("recv", r_), ("isend", si), ("irecv", ri)):
                out[f"{prefix}_{nm}_{nranks}_{rank}"] = arr.astype(np.int32)       # every field fits (block ids, boxes, tags)
        print(prefix, "ranks", nranks, "done", flush=True)
    return nblk, boxes


def main():
    out = {}
    centre = np.array([0.5 * L] * 3)
    # ---- bench grid
    roots, n = (8, 8, 8), (32, 32, 32)
    ref.set_amr([])
    _, root_boxes = ref.block_boxes(oracle_cfg(roots, n, NG), cap=1 << 15)
    pass1 = intersects(root_boxes, centre, 0.30 * L)
    ref.set_amr([list(pass1)])
    _, boxes1 = ref.block_boxes(oracle_cfg(roots, n, NG), cap=1 << 15)
    size1 = boxes1[:, 1] - boxes1[:, 0]
    fine = np.flatnonzero(size1 < 0.75 * (L / roots[0]))
    pass2 = fine[np.isin(fine, intersects(boxes1, centre, 0.12 * L))]
    passes = [list(map(int, pass1)), list(map(int, pass2))]
    ref.set_amr(passes)
    nblk, boxes = tables(out, "b", roots, n)
    sizes = np.round((L / roots[0]) / (boxes[:, 1] - boxes[:, 0])).astype(int)
    print("bench grid:", nblk, "blocks; per level", {int(s): int((sizes == s).sum()) for s in np.unique(sizes)})
    out["b_passes"] = np.array([len(p) for p in passes] + [b for p in passes for b in p])
    # ---- parity grid
    ref.set_amr([[0]])
    nblk, _ = tables(out, "p", (2, 2, 2), (32, 8, 8))
    print("parity grid:", nblk, "blocks")
    ref.set_amr([])
    path = os.path.join(HERE, "config5_amr.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
