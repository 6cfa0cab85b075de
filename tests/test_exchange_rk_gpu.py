"""Parity of the exchange kernels (bit-exact), the fused RK stage update, the reductions and short RK
trajectories against the oracle (reference src/grid/make_exchange.h:111-410,
src/time-integration/advance.h:57-102,236-402, src/algs/transform_reduce.h:53-191)."""
import ctypes as C

import numpy as np
import pytest

from util import GAMMA, RGAS, make_state, oracle_cfg, product_flux, product_setup, rel_l2, zero_ghosts, interior

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("periodic", [(1, 1, 1), (1, 0, 1), (0, 0, 0)])
@pytest.mark.parametrize("nb,n", [((2, 2, 2), (16, 16, 16)), ((1, 3, 2), (8, 4, 12)), ((1, 1, 1), (32, 32, 32))])
def test_single_rank_exchange_bit_exact(periodic, nb, n):
    from oracle import port
    ng = 2
    q = zero_ghosts(make_state(nb, n, ng, seed=2), ng)
    cfg = oracle_cfg(nb, n, ng, periodic=periodic)
    want = port.exchange(cfg, q.ravel()).reshape(q.shape)
    sp, blocks, grid = product_setup(nb, n, ng)
    qa = sp.grid_array.from_host(grid, q)
    sp.make_exchange(qa, periodic).exchange(qa)
    assert np.array_equal(qa.to_host(), want)


@pytest.mark.parametrize("size", [2, 3])
def test_pack_unpack_between_simulated_ranks_bit_exact(size):
    """All ranks of a `size`-rank partition are played on one GPU: pack on the sender, hand the message
    buffer over, unpack on the receiver; the union must equal the oracle's global exchange."""
    from oracle import port
    import torch
    import spade_b200.api as sp
    nb, n, ng = (2, 2, 3), (8, 8, 8), 2
    periodic = (1, 1, 0)
    q = zero_ghosts(make_state(nb, n, ng, seed=4), ng)
    cfg = oracle_cfg(nb, n, ng, periodic=periodic)
    want = port.exchange(cfg, q.ravel()).reshape(q.shape)
    ranks = []
    for r in range(size):
        _, blocks, grid = product_setup(nb, n, ng, rank=r, size=size)
        lo = grid.first_block
        qa = sp.grid_array.from_host(grid, q[lo:lo + grid.num_local_blocks])
        ranks.append((grid, qa, sp.make_exchange(qa, periodic)))
    lib = sp.lib()
    bufs = {}
    for r, (grid, qa, ex) in enumerate(ranks):
        for p in range(size):
            if p != r and ex.send_cells[p]:
                b = torch.empty(5 * ex.send_cells[p], dtype=torch.float64, device="cuda")
                sp.check(lib.spb_exchange_pack(ex._h, C.c_void_p(qa.data.data_ptr()), p, C.c_void_p(b.data_ptr()), None))
                bufs[(r, p)] = b
        sp.check(lib.spb_exchange_local(ex._h, C.c_void_p(qa.data.data_ptr()), None))
    for r, (grid, qa, ex) in enumerate(ranks):
        for p in range(size):
            if p != r and ex.recv_cells[p]:
                assert ex.recv_cells[p] == ranks[p][2].send_cells[r]
                sp.check(lib.spb_exchange_unpack(ex._h, C.c_void_p(qa.data.data_ptr()), p, C.c_void_p(bufs[(p, r)].data_ptr()), None))
    torch.cuda.synchronize()
    got = np.concatenate([qa.to_host() for _, qa, _ in ranks], axis=0)
    assert np.array_equal(got, want)


def test_reductions():
    from oracle import port
    nb, n, ng = (2, 1, 2), (16, 8, 12), 2
    q = make_state(nb, n, ng, seed=9)
    cfg = oracle_cfg(nb, n, ng)
    sp, blocks, grid = product_setup(nb, n, ng)
    qa = sp.grid_array.from_host(grid, q)
    gas = sp.ideal_gas_t(GAMMA, RGAS)
    assert sp.transform_reduce(qa, sp.FN_WAVESPEED, sp.RED_MAX, gas) == pytest.approx(port.reduce_umax(cfg, q.ravel()), rel=1e-15)
    qi = interior(q, ng)
    assert sp.transform_reduce(qa, sp.FN_VAR, sp.RED_MAX, gas, ivar=3) == qi[..., 3].max()
    assert sp.transform_reduce(qa, sp.FN_VAR, sp.RED_SUM, gas, ivar=0) == pytest.approx(qi[..., 0].sum(), rel=1e-13)
    assert sp.transform_reduce(qa, sp.FN_ABSVAR, sp.RED_MAX, gas, ivar=4) == np.abs(qi[..., 4]).max()


def _advance_product(nb, n, ng, q, scheme, integ, dt, nsteps, periodic=(1, 1, 1), fused=False, bc_object=False):
    sp, blocks, grid = product_setup(nb, n, ng)
    gas = sp.ideal_gas_t(GAMMA, RGAS)
    qa = sp.grid_array.from_host(grid, q)
    ra = sp.grid_array(grid, 0.0)
    ex = sp.make_exchange(qa, periodic)
    flux = sp.flux_desc(product_flux(scheme))
    alg = {0: sp.rk4_t, 1: sp.ssprk3_opt, 2: sp.ssprk3_t, 3: sp.rk2_t, 4: sp.ssprk34_t, 5: sp.rk38r_t}[integ]
    data = sp.integrator_data_t(qa, ra, alg)
    rhs_calc = sp.flux_div_rhs_t(flux, sp.overwrite) if fused else (lambda r, qq, t: sp.flux_div(qq, r, flux, sp.overwrite))
    ti = sp.integrator_t(sp.time_axis_t(0.0, dt), alg, data, rhs_calc,
                         sp.exchange_bc_t(ex) if bc_object else (lambda qq, t: ex.exchange(qq)), sp.state_transform_t(gas))
    assert (ti._plan is not None) == fused
    n0 = sp.launch_count()
    for _ in range(nsteps):
        ti.advance()
    if bc_object and fused:
        # one kernel per stage: no separate same-rank exchange launch, and the fused path was not abandoned
        assert ti._fuse_exchange and sp.launch_count() - n0 == nsteps * alg.rows()
    return ti.solution().to_host()


@pytest.mark.parametrize("integ", [0, 1, 2, 3, 4, 5])
def test_rk_trajectory_matches_oracle(integ):
    from oracle import port
    nb, n, ng = (2, 2, 1), (16, 8, 8), 2
    q0 = make_state(nb, n, ng, seed=13)
    cfg = oracle_cfg(nb, n, ng, scheme=0, integrator=integ)
    q0 = port.exchange(cfg, q0.ravel()).reshape(q0.shape)
    dx = 2 * np.pi / 32
    dt = 0.2 * dx / port.reduce_umax(cfg, q0.ravel())
    want = port.advance(cfg, q0.ravel(), dt, 5).reshape(q0.shape)
    got = _advance_product(nb, n, ng, q0, 0, integ, dt, 5)
    assert rel_l2(got, want) < 1e-12
    assert rel_l2(got - q0, want - q0) < 1e-9      # the increment itself, not just the state


@pytest.mark.parametrize("integ,scheme", [(i, s) for i in (0, 2, 3) for s in (0, 3, 4, 1, 2, 6, 11, 12)] + [(4, 0), (5, 0), (4, 1), (5, 1)])
def test_fused_stage_kernel_trajectory_matches_oracle(integ, scheme):
    """flux_div + RK stage update in one kernel (spb_flux_div_rk_stage), rk4 (with the pre-combined final update),
    ssprk3 (odd number of stages: result ends in the second buffer) and rk2; blocks that are not multiples of the tile."""
    from oracle import port
    nb, n, ng = (2, 1, 2), (40, 12, 8), 2
    q0 = make_state(nb, n, ng, seed=23, jump=scheme in (1, 6, 12))
    cfg = oracle_cfg(nb, n, ng, scheme=scheme, integrator=integ)
    q0 = port.exchange(cfg, q0.ravel()).reshape(q0.shape)
    dt = 0.2 * (2 * np.pi / 80) / port.reduce_umax(cfg, q0.ravel())
    want = port.advance(cfg, q0.ravel(), dt, 3).reshape(q0.shape)
    got = _advance_product(nb, n, ng, q0, scheme, integ, dt, 3, fused=True)
    assert rel_l2(got, want) < 1e-12
    assert rel_l2(got - q0, want - q0) < 1e-9
    unfused = _advance_product(nb, n, ng, q0, scheme, integ, dt, 3, fused=False)
    assert rel_l2(got, unfused) < 1e-13


@pytest.mark.parametrize("nb,n,periodic", [((2, 1, 2), (40, 12, 8), (1, 1, 1)), ((2, 2, 2), (16, 16, 16), (1, 1, 1)),
                                           ((1, 1, 1), (32, 32, 32), (1, 1, 1)), ((3, 2, 1), (32, 8, 16), (1, 0, 1)),
                                           ((2, 2, 3), (64, 16, 6), (0, 1, 0))])
@pytest.mark.parametrize("integ", [0, 2])
def test_fused_stage_kernel_fills_same_rank_ghosts_bit_exact(nb, n, periodic, integ):
    """spb_flux_div_rk_stage_exchange: the stage kernel also stores its q_out planes into the neighbour blocks' ghost
    cells. The ghosts must be bit-identical to an exchange of the result (make_exchange.h:166-203), the interior
    bit-identical to the fused stage kernel followed by a separate exchange, and the trajectory within 1e-12 of the
    oracle. Ragged tiles, 16^3 blocks, a block that is its own neighbour, walls (ghosts there stay untouched)."""
    from oracle import port
    ng = 2
    q0 = make_state(nb, n, ng, seed=29)
    cfg = oracle_cfg(nb, n, ng, scheme=0, integrator=integ, periodic=periodic)
    dt = 0.2 * (2 * np.pi / (nb[0] * n[0])) / port.reduce_umax(cfg, q0.ravel())
    want = port.advance(cfg, q0.ravel(), dt, 2).reshape(q0.shape)
    got = _advance_product(nb, n, ng, q0, 0, integ, dt, 2, periodic=periodic, fused=True, bc_object=True)
    sep = _advance_product(nb, n, ng, q0, 0, integ, dt, 2, periodic=periodic, fused=True, bc_object=False)
    assert np.array_equal(got, sep)                                      # interior and ghosts, bit for bit
    assert np.array_equal(got, port.exchange(cfg, got.ravel()).reshape(got.shape))   # ghosts == exchange(interior)
    assert rel_l2(interior(got, ng), interior(want, ng)) < 1e-12


def test_rk4_hybrid_weno_trajectory_and_conservation():
    from oracle import port
    nb, n, ng = (2, 2, 2), (8, 8, 8), 2
    q0 = make_state(nb, n, ng, seed=17, jump=True)
    cfg = oracle_cfg(nb, n, ng, scheme=1, integrator=0)
    q0 = port.exchange(cfg, q0.ravel()).reshape(q0.shape)
    dx = 2 * np.pi / 16
    dt = 0.2 * dx / port.reduce_umax(cfg, q0.ravel())
    want = port.advance(cfg, q0.ravel(), dt, 3).reshape(q0.shape)
    got = _advance_product(nb, n, ng, q0, 1, 0, dt, 3)
    assert rel_l2(got, want) < 1e-12

    def total_mass(q):
        qi = interior(q, ng)
        return (qi[..., 0] / (RGAS * qi[..., 1])).sum()
    assert total_mass(got) == pytest.approx(total_mass(q0), rel=1e-12)   # periodic box conserves mass


@pytest.mark.parametrize("integ,hs", [(0, False), (2, False), (3, False), (2, True), (3, True)])
def test_generic_integrate_advance_matches_oracle(integ, hs):
    """integrator_t with identity_transform: the generic path of advance.h:109-230 (spb_axpy_roundtrip reproduces the
    reference's scale / add / unscale passes bit for bit; the flux kernel is the only source of round-off)."""
    from oracle import port
    nb, n, ng = (2, 2, 1), (16, 8, 8), 2
    cfg = oracle_cfg(nb, n, ng, scheme=0, integrator=integ)
    q0 = port.exchange(cfg, make_state(nb, n, ng, seed=17).ravel())
    dt = 1e-5
    want = port.advance_generic(cfg, q0, dt, 3, hs)
    sp, blocks, grid = product_setup(nb, n, ng)
    alg = {(0, False): sp.rk4_t, (2, False): sp.ssprk3_t, (3, False): sp.rk2_t, (2, True): sp.ssprk3hs_t, (3, True): sp.rk2hs_t}[(integ, hs)]
    qa = sp.grid_array.from_host(grid, q0.reshape((-1, n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng, 5)))
    ra = sp.grid_array(grid, 0.0)
    ex = sp.make_exchange(qa, (1, 1, 1))
    flux = sp.flux_desc(product_flux(0))
    ti = sp.integrator_t(sp.time_axis_t(0.0, dt), alg, sp.integrator_data_t(qa, ra, alg),
                         lambda r, qq, t: sp.flux_div(qq, r, flux, sp.overwrite), lambda qq, t: ex.exchange(qq))
    for _ in range(3):
        ti.advance()
    got = ti.solution().to_host().ravel()
    assert rel_l2(got, want) < 1e-12
    assert rel_l2(got - q0, want - q0) < 1e-9


def test_checkpoint_file_is_the_reference_byte_order(tmp_path):
    """io::binary_write / binary_read (reference src/io/io_native.h:18-56): block lb_glob at byte offset lb_glob * block_bytes,
    padded block in the array's memory order, no header — i.e. the file is the global array as the reference stores it."""
    nb, n, ng = (2, 2, 1), (8, 4, 4), 2
    sp, blocks, grid = product_setup(nb, n, ng)
    q = make_state(nb, n, ng, seed=9)
    qa = sp.grid_array.from_host(grid, q)
    f = str(tmp_path / "q.bin")
    sp.io.binary_write(f, qa)
    assert open(f, "rb").read() == q.tobytes()
    back = sp.grid_array(grid, 0.0)
    sp.io.binary_read(f, back)
    assert np.array_equal(back.to_host(), q)
    # a rank owning blocks 2..3 of the same file reads its own slab
    sp2, _, grid2 = product_setup(nb, n, ng, rank=1, size=2)
    part = sp2.grid_array(grid2, 0.0)
    sp2.io.binary_read(f, part)
    assert np.array_equal(part.to_host(), q[grid2.first_block:grid2.first_block + grid2.num_local_blocks])


@pytest.mark.parametrize("scheme,ng", [(1, 2), (2, 2), (12, 2), (13, 3), (14, 4)])
@pytest.mark.parametrize("nb,n,periodic", [((2, 1, 2), (40, 12, 8), (1, 1, 1)), ((3, 2, 1), (32, 8, 16), (1, 0, 1)),
                                           ((1, 1, 1), (16, 16, 16), (1, 1, 1))])
def test_wide_fused_stage_kernel_fills_same_rank_ghosts_bit_exact(scheme, ng, nb, n, periodic):
    """The wide-stencil stage kernels (hybrid WENO, cent_keep<4|6|8>, WALE closure) with an exchange plan: the threads that own
    the cells of a block's outer shell also store them into the neighbours' ghost cells, for 2, 3 and 4 exchange cells. Ghosts
    bit-identical to an exchange of the result, everything bit-identical to the kernel followed by a separate exchange."""
    from oracle import port
    q0 = make_state(nb, n, ng, seed=37, jump=scheme in (1, 12))
    cfg = oracle_cfg(nb, n, ng, scheme=scheme, integrator=0, periodic=periodic)
    dt = 0.2 * (2 * np.pi / (nb[0] * n[0])) / port.reduce_umax(cfg, q0.ravel())
    sp, blocks, grid = product_setup(nb, n, ng)
    gas = sp.ideal_gas_t(GAMMA, RGAS)
    results = []
    for bc_object in (True, False):
        qa = sp.grid_array.from_host(grid, q0, (ng,) * 3)
        ra = sp.grid_array(grid, 0.0, (ng,) * 3)
        ex = sp.make_exchange(qa, periodic)
        ti = sp.integrator_t(sp.time_axis_t(0.0, dt), sp.rk4_t, sp.integrator_data_t(qa, ra, sp.rk4_t),
                             sp.flux_div_rhs_t(product_flux(scheme), sp.overwrite),
                             sp.exchange_bc_t(ex) if bc_object else (lambda qq, t: ex.exchange(qq)), sp.state_transform_t(gas))
        n0 = sp.launch_count()
        ti.advance()
        ti.advance()
        if bc_object:
            assert ti._fuse_exchange and sp.launch_count() - n0 == 2 * 4
        results.append(ti.solution().to_host())
    got, sep = results
    assert np.array_equal(got, sep)
    assert np.array_equal(got, port.exchange(cfg, got.ravel()).reshape(got.shape))
    want = port.advance(cfg, q0.ravel(), dt, 2).reshape(q0.shape)
    assert rel_l2(interior(got, ng), interior(want, ng)) < 1e-12
