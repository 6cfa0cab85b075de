#!/usr/bin/env python
"""Per-kernel timing of the hot path through the C ABI (CUDA events, warm, inputs larger than L2).
Development tool: `python tools/kbench.py [--lattice 8 8 8] [--scheme central|hybrid] [--iters 10]`.
Prints one JSON object per kernel with algorithmic GB/s (SURVEY 8d byte counts)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lattice", type=int, nargs=3, default=[8, 8, 8])
    ap.add_argument("--block", type=int, default=32)
    ap.add_argument("--scheme", default="central")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--coords", default="identity", help="identity | channel (diagonal_coords: scaled x, integrated tanh y)")
    a = ap.parse_args()
    import torch
    import bench
    import spade_b200.api as sp
    bench.BLOCK = a.block
    torch.cuda.set_device(0)
    lat = tuple(a.lattice)
    L = 2 * 3.141592653589793
    blocks = sp.cartesian_blocks_t(lat, [0.0, L] * 3)
    coords = sp.identity() if a.coords == "identity" else sp.diagonal_coords(sp.scaled_coord_1D(2.0), sp.integrated_tanh_1D(0.0, L, 0.1, 4.0), None)
    grid = sp.cartesian_grid_t((a.block,) * 3, blocks, coords, sp.pool_t())
    gas = sp.ideal_gas_t(bench.GAMMA, bench.RGAS)
    mu = (bench.P0 / (bench.RGAS * bench.T0)) * bench.U0 / bench.REYNOLDS
    visc = sp.visc_lr(sp.constant_viscosity_t(mu, bench.PRANDTL), gas)
    schemes = {"central": sp.compose(sp.totani_lr(gas), visc),
               "hybrid": sp.compose(sp.hybrid_scheme_t(sp.totani_lr(gas), sp.fweno_t(gas), sp.ducros_t(1e-2), sp.full_flux), visc),
               "ck4": sp.compose(sp.cent_keep(4, gas), visc),
               "euler": sp.totani_lr(gas)}
    flux = sp.flux_desc(schemes[a.scheme])
    q = bench.device_state(sp, grid, torch)
    ks = [sp.grid_array(grid, 0.0) for _ in range(4)]
    ex = sp.make_exchange(q, (1, 1, 1))
    ex.exchange(q)
    cells = grid.local_cells()
    n = a.block
    ghost_frac = ((n + 4) ** 3 - n ** 3) / n ** 3
    dt = 1e-7

    def timeit(name, fn, bytes_per_cell):
        if a.only and not any(o in name for o in a.only.split(';')):
            return
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
        print(json.dumps({"kernel": name, "ms": round(ms, 4), "Gcells_per_s": round(cells / ms / 1e6, 2),
                          "alg_GBps": round(bytes_per_cell * cells / ms / 1e6, 1)}), flush=True)

    import ctypes as C
    lib = sp.lib()

    def rk(nk):
        coeff = (C.c_double * nk)(*([dt] * nk))
        kp = (C.c_void_p * nk)(*[k.data.data_ptr() for k in ks[:nk]])
        sp.check(lib.spb_rk_update(q.h, C.c_void_p(q.data.data_ptr()), kp, nk, coeff, gas.gamma, gas.R, None))

    import ctypes as C
    lib = sp.lib()
    timeit(f"flux_div[{a.scheme}]", lambda: sp.flux_div(q, ks[0], flux, sp.overwrite), 80.0)
    timeit(f"flux_div_incr[{a.scheme}]", lambda: sp.flux_div(q, ks[0], flux, sp.increment), 120.0)
    q2 = q.clone()
    from spade_b200._lib import StageDesc

    def fused(nin, out, ghost=False):
        sd = StageDesc()
        sd.nin = nin
        for i in range(nin):
            sd.inp[i] = ks[i].data.data_ptr()
            sd.cq[i] = dt
            sd.co[i] = 0.5
        sd.cq_self, sd.co_self = dt, 1.0
        sd.out = ks[2].data.data_ptr() if out else None
        sp.check(lib.spb_flux_div_rk_stage_exchange(q.h, C.c_void_p(q.data.data_ptr()), C.c_void_p(q2.data.data_ptr()), C.byref(flux),
                                                    C.byref(sd), ex._h if ghost else None, 0, grid.num_local_blocks, None))

    if a.scheme in ("central", "euler", "hybrid", "ck4"):
        for nin, out in ((0, 1), (1, 1), (2, 1), (1, 0)):
            timeit(f"fused_stage[nin={nin},out={out}]", lambda nin=nin, out=out: fused(nin, out), 80.0 + 40.0 * nin + 40.0 * out)
    if a.scheme in ("central", "euler") and a.coords == "identity":
        for nin, out in ((0, 1), (1, 1), (2, 1), (1, 0)):
            timeit(f"fused_stage+ghosts[nin={nin},out={out}]", lambda nin=nin, out=out: fused(nin, out, True),
                   80.0 + 40.0 * nin + 40.0 * out + 40.0 * ghost_frac)
    timeit("exchange", lambda: ex.exchange(q), 80.0 * ghost_frac)
    for nk in (1, 2, 4):
        timeit(f"rk_update[nk={nk}]", lambda nk=nk: rk(nk), 80.0 + 40.0 * nk)
    timeit("reduce_umax", lambda: sp.transform_reduce(q, sp.FN_WAVESPEED, sp.RED_MAX, gas), 40.0)


if __name__ == "__main__":
    main()
