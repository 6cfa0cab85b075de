"""Worker of tests/test_multigpu_nccl.py: one process per GPU over NCCL. The RK trajectory of the block-partitioned
grid (SPADE's contiguous partition, ghost messages over NVLink) must match the oracle's single-domain result:
exchange bit-exact, trajectories to 1e-12 (reference src/grid/make_exchange.h:111-410, src/grid/partition.h:27-84)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import GAMMA, RGAS, make_state, oracle_cfg, product_flux, rel_l2, zero_ghosts  # noqa: E402


def amr_case(sp, port, pool, rank, world):
    """AMR grid partitioned over the ranks (BASELINE config 5; tables of the unmodified reference for this rank count,
    tests/golden/config5_amr.npz): exchange bit-exact, and an RK4 trajectory through the overlapped schedule — the donor
    blocks of off-rank interpolation sends must be advanced BEFORE their messages are packed (spb_exchange_boundary_blocks)."""
    fix = np.load(os.path.join(ROOT, "tests", "golden", "config5_amr.npz"))
    if f"p_send_{world}_0" not in fix:
        return
    from util import amr_state
    boxes, n, ng = fix["p_boxes"], tuple(int(x) for x in fix["p_cells"]), 2
    per, extra = divmod(len(boxes), world)
    lo, nloc = rank * per + min(rank, extra), per + (1 if rank < extra else 0)
    tabs = tuple(fix[f"p_{k}_{world}_{rank}"].astype(np.int64) for k in ("send", "recv", "isend", "irecv"))
    cfg = oracle_cfg(tuple(int(x) for x in fix["p_roots"]), n, ng, scheme=0, integrator=0)
    q0 = amr_state(boxes, n, ng, seed=61)
    port.set_amr(boxes, np.concatenate([fix[f"p_send_{world}_{r}"] for r in range(world)]).astype(np.int64),
                 np.concatenate([fix[f"p_isend_{world}_{r}"] for r in range(world)]).astype(np.int64))
    try:
        qz = zero_ghosts(q0, ng)
        want = port.exchange(cfg, qz.ravel()).reshape(q0.shape)
        grid = sp.cartesian_grid_t.from_boxes(n, boxes[lo:lo + nloc], pool, first_block=lo)
        qa = sp.grid_array.from_host(grid, qz[lo:lo + nloc])
        ex = sp.make_exchange(qa, (1, 1, 1), tables=tabs)
        ex.exchange(qa)
        assert np.array_equal(qa.to_host(), want[lo:lo + nloc]), f"rank {rank}: AMR exchange differs from the oracle"
        qe = port.exchange(cfg, q0.ravel()).reshape(q0.shape)
        dt = 0.2 * float((boxes[:, 1] - boxes[:, 0]).min()) / n[0] / port.reduce_umax(cfg, qe.ravel())
        want = port.advance(cfg, qe.ravel(), dt, 2).reshape(q0.shape)
        gas = sp.ideal_gas_t(GAMMA, RGAS)
        qa = sp.grid_array.from_host(grid, qe[lo:lo + nloc])
        ex = sp.make_exchange(qa, (1, 1, 1), tables=tabs)
        ti = sp.integrator_t(sp.time_axis_t(0.0, dt), sp.rk4_t, sp.integrator_data_t(qa, sp.grid_array(grid, 0.0), sp.rk4_t),
                             sp.flux_div_rhs_t(sp.flux_desc(product_flux(0)), sp.overwrite), sp.exchange_bc_t(ex), sp.state_transform_t(gas))
        for _ in range(2):
            ti.advance()
        err = rel_l2(ti.solution().to_host(), want[lo:lo + nloc])
        assert err < 1e-12 and ti._plan is not None, f"rank {rank} AMR fused + overlap: rel L2 {err}"
    finally:
        port.set_amr()


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    opts = dist.ProcessGroupNCCL.Options()
    opts.is_high_priority_stream = True
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
    rank, world = dist.get_rank(), dist.get_world_size()
    import spade_b200.api as sp
    from oracle import port
    pool = sp.pool_t.from_torch()

    for nb, n, periodic in (((2, 2, 4), (32, 16, 8), (1, 1, 1)), ((3, 1, 2), (16, 16, 16), (1, 0, 1)), ((2, 3, 4), (16, 8, 8), (1, 0, 1))):
        ng = 2
        if nb[0] * nb[1] * nb[2] < world:
            continue                      # every rank must own a block (SPADE's partition gives the first nblocks % size ranks one extra)
        blocks = sp.cartesian_blocks_t(nb, [0.0, 2 * np.pi] * 3)
        grid = sp.cartesian_grid_t(n, blocks, sp.identity(), pool)
        lo, nloc = grid.first_block, grid.num_local_blocks
        q0 = make_state(nb, n, ng, seed=31)
        cfg = oracle_cfg(nb, n, ng, scheme=0, integrator=0, periodic=periodic)

        # exchange alone: bit-exact
        qz = zero_ghosts(q0, ng)
        want = port.exchange(cfg, qz.ravel()).reshape(q0.shape)
        qa = sp.grid_array.from_host(grid, qz[lo:lo + nloc])
        ex = sp.make_exchange(qa, periodic)
        ex.exchange(qa)
        assert np.array_equal(qa.to_host(), want[lo:lo + nloc]), f"rank {rank}: exchange differs from the oracle"

        q0 = port.exchange(cfg, q0.ravel()).reshape(q0.shape)
        dt = 0.2 * (2 * np.pi / (nb[0] * n[0])) / port.reduce_umax(cfg, q0.ravel())
        want = port.advance(cfg, q0.ravel(), dt, 2).reshape(q0.shape)
        gas = sp.ideal_gas_t(GAMMA, RGAS)
        flux = sp.flux_desc(product_flux(0))
        results = []
        for mode in ("unfused", "fused", "fused+overlap"):
            qa = sp.grid_array.from_host(grid, q0[lo:lo + nloc])
            ra = sp.grid_array(grid, 0.0)
            ex = sp.make_exchange(qa, periodic)
            data = sp.integrator_data_t(qa, ra, sp.rk4_t)
            rhs = (lambda r, qq, t: sp.flux_div(qq, r, flux, sp.overwrite)) if mode == "unfused" else sp.flux_div_rhs_t(flux, sp.overwrite)
            bc = sp.exchange_bc_t(ex) if mode == "fused+overlap" else (lambda qq, t: ex.exchange(qq))
            ti = sp.integrator_t(sp.time_axis_t(0.0, dt), sp.rk4_t, data, rhs, bc, sp.state_transform_t(gas))
            for _ in range(2):
                ti.advance()
            got = ti.solution().to_host()
            err = rel_l2(got, want[lo:lo + nloc])
            assert err < 1e-12, f"rank {rank} {mode}: rel L2 {err}"
            results.append(got)
        assert np.array_equal(results[1], results[2]), f"rank {rank}: overlapped schedule changes the result"
        # a wide-stencil functor set (hybrid WENO + viscous): the stage kernel's owning threads store the same-rank ghosts, the
        # off-rank ghosts travel as messages, boundary and interior blocks run on two streams
        cfg1 = oracle_cfg(nb, n, ng, scheme=1, integrator=0, periodic=periodic)
        want1 = port.advance(cfg1, q0.ravel(), dt, 2).reshape(q0.shape)
        qh = sp.grid_array.from_host(grid, q0[lo:lo + nloc])
        exh = sp.make_exchange(qh, periodic)
        th = sp.integrator_t(sp.time_axis_t(0.0, dt), sp.rk4_t, sp.integrator_data_t(qh, sp.grid_array(grid, 0.0), sp.rk4_t),
                             sp.flux_div_rhs_t(sp.flux_desc(product_flux(1)), sp.overwrite), sp.exchange_bc_t(exh), sp.state_transform_t(gas))
        for _ in range(2):
            th.advance()
        err = rel_l2(th.solution().to_host(), want1[lo:lo + nloc])
        assert err < 1e-12 and th._fuse_exchange, f"rank {rank} hybrid fused: rel L2 {err}"
        umax = sp.transform_reduce(qa, sp.FN_WAVESPEED, sp.RED_MAX, gas)
        umax_want = port.reduce_umax(cfg, want.ravel())
        local_now = rel_l2(qa.to_host(), want[lo:lo + nloc])
        assert umax == umax_want or abs(umax / umax_want - 1) < 1e-12, \
            f"rank {rank} lattice {nb}: umax {umax!r} vs oracle {umax_want!r}; rel L2 of the reduced array now {local_now:.3e}"
    amr_case(sp, port, pool, rank, world)
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok p2p={int(bool(ex._p2p))}")


if __name__ == "__main__":
    main()
