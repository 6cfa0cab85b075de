#!/usr/bin/env python
"""BASELINE.json configs[2]: compressible channel on a stretched grid — hybrid(totani_lr, fweno_t, ducros_t) + visc_lr,
rk4_t, x/z periodic, isothermal no-slip walls in y, y = integrated_tanh_1D, 32^3 blocks, SPADE's contiguous block partition.

    python tools/config3.py [--lattice 32 16 2] [--steps 5] [--warmup 2]            (lattice = blocks PER GPU)
    python -m torch.distributed.run --nproc-per-node 8 ... tools/config3.py         (8 GPUs: 32 x 16 x 16 blocks = 1024x512x512)

Not the headline bench (bench.py measures configs[1]); a record of the config-3 path for profiles/. Prints one JSON line:
cell-stage-updates/s (whole job, device time, max over ranks) and the stage kernel's share. The kernel is FP64-bound
(~2 kflop per cell), so the roofline figure is the fraction of the FP64 pipe, not of HBM."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lattice", type=int, nargs=3, default=[32, 16, 2])
    ap.add_argument("--block", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--rate", type=float, default=1.3, help="integrated_tanh_1D stretching rate")
    ap.add_argument("--coords", default="tanh", choices=["tanh", "identity"])
    a = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    import spade_b200.api as sp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)
    pool = sp.pool_t.from_torch()
    n, B, NG = world, a.block, 2
    lat = (a.lattice[0], a.lattice[1], a.lattice[2] * n)
    pi = float(np.pi)
    bounds = [0.0, 4 * pi, -1.0, 1.0, 0.0, 2 * pi]
    blocks = sp.cartesian_blocks_t(lat, bounds)
    coords = sp.identity() if a.coords == "identity" else sp.diagonal_coords(None, sp.integrated_tanh_1D(-1.0, 1.0, 0.1, a.rate), None)
    grid = sp.cartesian_grid_t((B,) * 3, blocks, coords, pool)
    gamma, rgas, p0, t0, u0 = 1.4, 287.15, 101325.0, 300.0, 69.4
    gas = sp.ideal_gas_t(gamma, rgas)
    mu = (p0 / (rgas * t0)) * u0 / 3000.0
    conv = sp.hybrid_scheme_t(sp.totani_lr(gas), sp.fweno_t(gas), sp.ducros_t(1e-2), sp.full_flux)
    flux = sp.flux_desc(sp.compose(conv, sp.visc_lr(sp.constant_viscosity_t(mu, 0.72), gas)))

    # laminar parabolic profile + seeded perturbation + a planted pressure jump that wakes the shock sensor
    q = sp.grid_array(grid, 0.0, (NG,) * 3)
    idx = torch.arange(-NG, B + NG, dtype=torch.float64, device="cuda") + 0.5
    gen = torch.Generator(device="cuda").manual_seed(12345 + rank)
    for b0 in range(0, grid.num_local_blocks, 128):
        b1 = min(grid.num_local_blocks, b0 + 128)
        org = torch.tensor([blocks.get_block_box(grid.first_block + l)[0::2] for l in range(b0, b1)], dtype=torch.float64, device="cuda")
        dx = [grid.get_dx(d) for d in range(3)]
        X = (org[:, 0, None] + idx[None, :] * dx[0])[:, None, None, :]
        Y = (org[:, 1, None] + idx[None, :] * dx[1])[:, None, :, None]
        Z = (org[:, 2, None] + idx[None, :] * dx[2])[:, :, None, None]
        v = q.data[b0:b1]
        v[..., 0] = p0 * torch.where(torch.sin(0.5 * X) > 0.3, 1.2, 1.0) + 0 * Y + 0 * Z
        v[..., 1] = t0 + 0 * X + 0 * Y + 0 * Z
        v[..., 2] = u0 * (1 - Y * Y) + 0 * X + 0 * Z
        v[..., 3] = 0.02 * u0 * torch.sin(X) * torch.cos(pi * Y) * torch.cos(Z)
        v[..., 4] = 0.02 * u0 * torch.sin(Z) * torch.cos(X + pi * Y)
        v *= 1 + 1e-3 * (2 * torch.rand(v.shape, dtype=torch.float64, device="cuda", generator=gen) - 1)
    rhs = sp.grid_array(grid, 0.0, (NG,) * 3)
    handle = sp.make_exchange(q, (True, False, True))
    wall = sp.noslip_isothermal_wall(t0)
    bc = sp.exchange_bc_t(handle, sp.boundary.ymin | sp.boundary.ymax, wall)
    bc(q, 0.0)
    umax = sp.transform_reduce(q, sp.FN_WAVESPEED, sp.RED_MAX, gas)
    dmin = min(grid.get_dx(0), grid.get_dx(2))
    if a.coords == "tanh":
        _, jac, _ = grid.metric_tables((NG,) * 3)
        dmin = min(dmin, float(np.abs(jac[1][:, NG:-NG]).min()) * grid.get_dx(1))
    else:
        dmin = min(dmin, grid.get_dx(1))
    dt = 0.2 * dmin / umax
    alg = sp.rk4_t
    ti = sp.integrator_t(sp.time_axis_t(0.0, dt), alg, sp.integrator_data_t(q, rhs, alg), sp.flux_div_rhs_t(flux, sp.overwrite), bc,
                         sp.state_transform_t(gas))
    assert ti._plan is not None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        ti.advance()
    barrier()
    ev = []
    ti.stage_events = ev
    n0 = sp.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        ti.advance()
    e1.record()
    barrier()
    ti.stage_events = None
    ms = e0.elapsed_time(e1)
    stage_ms = sum(x.elapsed_time(y) for x, y, _ in ev) / max(1, len(ev))
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    um = sp.transform_reduce(q, sp.FN_WAVESPEED, sp.RED_MAX, gas)
    if not (um == um) or um > 10 * umax:
        raise SystemExit(f"config3: solution diverged (umax {um})")
    cells = grid.local_cells() * n
    if rank == 0:
        print(json.dumps({
            "workload": f"channel {lat[0]*B}x{lat[1]*B}x{lat[2]*B} cells ({lat[0]}x{lat[1]}x{lat[2]} blocks of {B}^3), y = "
                        + (f"integrated_tanh_1D(-1, 1, 0.1, {a.rate})" if a.coords == "tanh" else "uniform")
                        + ", hybrid(totani_lr, fweno_t, ducros_t(1e-2), full_flux) + visc_lr, rk4_t, walls in y (no-slip isothermal), x/z periodic",
            "n_gpus": n, "steps": a.steps, "warmup": a.warmup, "value": cells * 4 * a.steps / (ms * 1e-3), "unit": "cell-stage-updates/s",
            "ms_per_step": ms / a.steps, "stage_kernel_ms": stage_ms, "stage_kernel_share": stage_ms * 4 * a.steps / ms,
            "stage_kernel_cell_evals_per_s": grid.local_cells() / (stage_ms * 1e-3), "gpu_launches": sp.launch_count() - n0,
            "umax": um, "dt": dt, "partition": f"contiguous block runs, rank r = z-slab r ({n} ranks)"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
