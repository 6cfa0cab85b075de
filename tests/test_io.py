"""On-disk formats either side of the path: io::output_vtk (reference src/io/io_vtk.h:276-288) and
io::binary_write / binary_read (src/io/io_native.h:18-56). tests/golden/vtk_small/ holds the files the UNMODIFIED reference
wrote for a 2x1x2 lattice of 8x4x4 blocks (tests/golden/make_vtk_golden.cc); the drop-in's writers must produce the same
bytes."""
import filecmp
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "vtk_small")
NB, N, NG = (2, 1, 2), (8, 4, 4), 2
BOUNDS = [0.0, 2.0, 0.0, 1.0, 0.0, 3.0]


def golden_state():
    nlb = NB[0] * NB[1] * NB[2]
    size = nlb * (N[2] + 2 * NG) * (N[1] + 2 * NG) * (N[0] + 2 * NG) * 5
    o = np.arange(size, dtype=np.int64)
    return ((o % 1013) * 0.125 + (o % 5)).reshape(nlb, N[2] + 2 * NG, N[1] + 2 * NG, N[0] + 2 * NG, 5)


def test_vtk_files_match_the_reference_byte_for_byte(tmp_path):
    import spade_b200.api as sp
    blocks = sp.cartesian_blocks_t(NB, BOUNDS)
    boxes = np.array([blocks.get_block_box(lb) for lb in range(blocks.total_num_blocks())])
    sp.io.write_vtk_files(str(tmp_path), "sol", golden_state(), N, (NG,) * 3, boxes, range(len(boxes)), len(boxes))
    assert filecmp.cmp(os.path.join(tmp_path, "sol.pvts"), os.path.join(GOLD, "sol.pvts"), shallow=False)
    for lb in range(len(boxes)):
        name = os.path.join("data_sol", f"b{lb:09d}.vts")
        assert filecmp.cmp(os.path.join(tmp_path, name), os.path.join(GOLD, name), shallow=False), name


def test_vtk_pieces_of_two_ranks_are_the_one_rank_files(tmp_path):
    """every rank writes only its own blocks (global block id in the file name); together they are the same set of files"""
    import spade_b200.api as sp
    blocks = sp.cartesian_blocks_t(NB, BOUNDS)
    boxes = np.array([blocks.get_block_box(lb) for lb in range(blocks.total_num_blocks())])
    q = golden_state()
    for lo, hi, root in ((0, 2, True), (2, 4, False)):
        sp.io.write_vtk_files(str(tmp_path), "sol", q[lo:hi], N, (NG,) * 3, boxes[lo:hi], range(lo, hi), len(boxes), write_base=root)
    for lb in range(len(boxes)):
        name = os.path.join("data_sol", f"b{lb:09d}.vts")
        assert filecmp.cmp(os.path.join(tmp_path, name), os.path.join(GOLD, name), shallow=False), name


@pytest.mark.gpu
def test_gpu_output_vtk_and_checkpoint_roundtrip(tmp_path):
    import spade_b200.api as sp
    blocks = sp.cartesian_blocks_t(NB, BOUNDS)
    grid = sp.cartesian_grid_t(N, blocks, sp.identity(), sp.pool_t())
    qa = sp.grid_array.from_host(grid, golden_state())
    sp.io.output_vtk(str(tmp_path), "sol", qa)
    for lb in range(4):
        name = os.path.join("data_sol", f"b{lb:09d}.vts")
        assert filecmp.cmp(os.path.join(tmp_path, name), os.path.join(GOLD, name), shallow=False), name
    f = os.path.join(tmp_path, "q.bin")
    sp.io.binary_write(f, qa)
    assert np.array_equal(np.fromfile(f), golden_state().ravel())          # the file IS the array's memory, block after block
    qb = sp.grid_array(grid, 0.0)
    sp.io.binary_read(f, qb)
    assert np.array_equal(qb.to_host(), golden_state())
