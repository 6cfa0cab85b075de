#!/usr/bin/env python
"""Pack / unpack kernel timing for the message sizes of the multi-GPU bench, on ONE GPU: the plan of rank 0 of `ranks` for a
16x16x(16*ranks) (or 8x8x(8*ranks)) lattice, messages packed into and unpacked from local buffers.
    python tools/xbench.py [--lattice 16] [--ranks 8]"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lattice", type=int, default=16)
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    import torch
    import spade_b200.api as sp
    lib = sp.lib()
    L, n = a.lattice, a.ranks
    pool = sp.pool_t(0, n)
    blocks = sp.cartesian_blocks_t((L, L, L * n), [0.0, 1.0, 0.0, 1.0, 0.0, float(n)])
    grid = sp.cartesian_grid_t((32,) * 3, blocks, sp.identity(), pool)
    q = sp.grid_array(grid, 1.0)
    ex = sp.make_exchange(q, (1, 1, 1))
    peers = [p for p in range(n) if p != 0 and ex.send_cells[p]]
    bufs = {p: torch.zeros(5 * max(ex.send_cells[p], ex.recv_cells[p]), dtype=torch.float64, device="cuda") for p in peers}
    dq = C.c_void_p(q.data.data_ptr())

    def timeit(name, fn, nbytes):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        print(json.dumps({"kernel": name, "ms": round(ms, 4), "MB": round(nbytes / 1e6, 1), "GBps": round(nbytes / ms / 1e6, 1)}), flush=True)

    for p in peers:
        nb = 80 * ex.send_cells[p]
        timeit(f"pack -> peer {p}", lambda p=p: sp.check(lib.spb_exchange_pack(ex._h, dq, p, C.c_void_p(bufs[p].data_ptr()), None)), nb)
        timeit(f"unpack <- peer {p}", lambda p=p: sp.check(lib.spb_exchange_unpack(ex._h, dq, p, C.c_void_p(bufs[p].data_ptr()), None)), nb)
    timeit("same-rank exchange", lambda: sp.check(lib.spb_exchange_local(ex._h, dq, None)), 80 * ex.send_cells[0])


if __name__ == "__main__":
    main()
