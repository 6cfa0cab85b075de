"""Worker of tests/test_multirank_gloo.py: world_size ranks over gloo on CPU. Exercises the host side of the
N > 1 path: partition, per-rank exchange plans, the message layout and the send/recv plumbing of
arr_exchange_t (with numpy standing in for the pack/unpack kernels, which need a GPU), and pool_t.reduce."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import make_state, oracle_cfg, zero_ghosts  # noqa: E402


def box_view(q, lb, mn, size, ng):
    """q[lb] cells [mn, mn+size) in (i,j,k) -> view shaped [sz, sy, sx, 5] (reference order: ix fastest)."""
    i0, j0, k0 = (int(m) + ng for m in mn)
    return q[lb, k0:k0 + size[2], j0:j0 + size[1], i0:i0 + size[0], :]


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import spade_b200.api as sp
    from oracle import port
    nb, n, ng, periodic = (2, 2, 3), (8, 4, 4), 2, (1, 1, 0)
    pool = sp.pool_t.from_torch()
    assert pool.rank() == rank and pool.size() == world

    # --- partition (partition.h:40-71): contiguous runs, the library's own builder on every rank
    plan = sp.arr_exchange_t.__new__(sp.arr_exchange_t)
    import ctypes as C
    from spade_b200._lib import lib, check, int3
    h = C.c_void_p()
    check(lib().spb_exchange_create(C.byref(h), int3(nb), int3(n), int3((ng,) * 3), int3(periodic), rank, world))
    nloc, first = int(lib().spb_exchange_local_blocks(h)), int(lib().spb_exchange_first_block(h))
    g2r, g2l = port.partition(nb[0] * nb[1] * nb[2], world)
    assert nloc == int((g2r == rank).sum()) and first == int(np.argmax(g2r == rank))
    ns, nr = lib().spb_exchange_num_send(h), lib().spb_exchange_num_recv(h)
    send = np.zeros((ns, 16), dtype=np.int64)
    recv = np.zeros((nr, 16), dtype=np.int64)
    offs = np.zeros((world, 6), dtype=np.int64)
    i64 = C.POINTER(C.c_int64)
    check(lib().spb_exchange_tables(h, send.ctypes.data_as(i64), recv.ctypes.data_as(i64), offs.ctypes.data_as(i64)))

    # --- plans agree across ranks: what r sends to p is, transaction by transaction, what p expects from r
    allsend, allrecv = [None] * world, [None] * world
    dist.all_gather_object(allsend, send)
    dist.all_gather_object(allrecv, recv)
    cols = [0, 1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12, 13, 14]          # everything but the two local-block ids
    for p in range(world):
        mine = send[send[:, 2] == p][:, cols]
        theirs = allrecv[p][allrecv[p][:, 1] == rank][:, cols]
        assert np.array_equal(mine, theirs), f"rank {rank} -> {p}: send/recv lists differ"
        assert offs[p, 0] == int((send[send[:, 2] == p][:, 9:12].prod(axis=1)).sum())

    # --- the data path with numpy pack/unpack in the reference message layout (make_exchange.h:52-79)
    qglob = zero_ghosts(make_state(nb, n, ng, seed=4), ng)
    want = port.exchange(oracle_cfg(nb, n, ng, periodic=periodic), qglob.ravel()).reshape(qglob.shape)
    q = qglob[first:first + nloc].copy()
    sendbufs, recvbufs = {}, {}
    for p in range(world):
        if p == rank:
            continue
        tr = send[send[:, 2] == p]
        if len(tr):
            sendbufs[p] = torch.from_numpy(np.concatenate([box_view(q, t[8], t[5:8], t[9:12], ng).reshape(-1) for t in tr]))
        ncell = int(offs[p, 1])
        if ncell:
            recvbufs[p] = torch.empty(5 * ncell, dtype=torch.float64)
    plan.pool = pool
    for req in plan.sendrecv(sendbufs, recvbufs):
        req.wait()
    for t in send[send[:, 2] == rank]:                                # same-rank transactions: direct copies
        box_view(q, t[15], t[12:15], t[9:12], ng)[...] = box_view(q, t[8], t[5:8], t[9:12], ng)
    for p, buf in recvbufs.items():
        pos = 0
        for t in recv[recv[:, 1] == p]:
            cnt = int(5 * t[9] * t[10] * t[11])
            box_view(q, t[15], t[12:15], t[9:12], ng)[...] = buf.numpy()[pos:pos + cnt].reshape(t[11], t[10], t[9], 5)
            pos += cnt
        assert pos == buf.numel()
    assert np.array_equal(q, want[first:first + nloc]), f"rank {rank}: exchanged ghosts differ from the oracle"
    lib().spb_exchange_destroy(h)

    # --- overlap schedule: the runs of rank-boundary blocks are exactly the source blocks of the off-rank sends
    blocks = sp.cartesian_blocks_t(nb, [0.0, 1.0] * 3)
    grid = sp.cartesian_grid_t(n, blocks, sp.identity(), pool)
    handle = sp.arr_exchange_t(grid, (ng,) * 3, periodic)
    first_runs, second_runs = handle.boundary_block_runs()
    covered = sorted(b for b0, b1 in first_runs for b in range(b0, b1))
    assert covered == sorted(set(int(t[8]) for t in send if t[2] != rank))
    rest = sorted(b for b0, b1 in second_runs for b in range(b0, b1))
    assert sorted(covered + rest) == list(range(nloc)) and not set(covered) & set(rest)

    # --- pool_t.reduce (compute_pool.h:247-284) and sync
    assert pool.reduce(float(rank + 1), sp.RED_MAX) == float(world)
    assert pool.reduce(float(rank + 1), sp.RED_SUM) == world * (world + 1) / 2
    pool.sync()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
