"""The other half of a channel run on the GPU: algs::boundary_fill (reference src/grid/boundary_fill.h:32-133) and
pde_algs::source_term (src/pde-algs/source_term.h:25-51) against the oracle — bit-exact for the ghost fill and the
source term, 1e-12 for wall-bounded RK trajectories (exchange + boundary_fill + flux_div + source_term + RK)."""
import numpy as np
import pytest

from util import GAMMA, RGAS, make_state, oracle_cfg, product_flux, product_setup, rel_l2

pytestmark = pytest.mark.gpu
WALL_T = 310.0


def _kernels(sp):
    return {
        "noslip_isothermal_y": (sp.boundary.ymin | sp.boundary.ymax, sp.noslip_isothermal_wall(WALL_T),
                                dict(mask=(0, 0, 1, 1, 0, 0), a=(1, -1, -1, -1, -1), b=(0, 2 * WALL_T, 0, 0, 0))),
        "adiabatic_all": (sp.identifier_t(1, 1, 1, 1, 1, 1), sp.noslip_adiabatic_wall(), dict(mask=(1,) * 6, a=(1, 1, -1, -1, -1))),
        "symmetry_x_z": (sp.boundary.xmin | sp.boundary.xmax | sp.boundary.zmin | sp.boundary.zmax, sp.symmetry_plane(),
                         dict(mask=(1, 1, 0, 0, 1, 1), a=(1, 1, 1, 1, 1), a_normal=-1.0)),
        "extrap1_ymax": (sp.boundary.ymax, sp.boundary.extrapolate(1), dict(mask=(0, 0, 0, 1, 0, 0), kind=1, order=1)),
        "extrap2_all": (sp.identifier_t(1, 1, 1, 1, 1, 1), sp.boundary.extrapolate(2), dict(mask=(1,) * 6, kind=1, order=2)),
    }


@pytest.mark.parametrize("name", ["noslip_isothermal_y", "adiabatic_all", "symmetry_x_z", "extrap1_ymax", "extrap2_all"])
def test_boundary_fill_bit_exact(name):
    from oracle import port, ref
    nb, n, ng = (2, 3, 2), (8, 4, 8), 2
    sp, blocks, grid = product_setup(nb, n, ng)
    which, kern, o = _kernels(sp)[name]
    q = make_state(nb, n, ng, seed=41)
    want = port.boundary_fill(oracle_cfg(nb, n, ng, periodic=(0, 0, 0)), ref.make_bc(**o), q.ravel()).reshape(q.shape)
    qa = sp.grid_array.from_host(grid, q)
    sp.boundary_fill(qa, which, kern)
    assert np.array_equal(qa.to_host(), want)


def test_source_term_bit_exact():
    from oracle import port, ref
    nb, n, ng = (2, 1, 2), (8, 8, 4), 2
    sp, blocks, grid = product_setup(nb, n, ng)
    q = make_state(nb, n, ng, seed=42)
    rhs0 = np.random.default_rng(3).standard_normal(q.shape)
    want = port.source_term(oracle_cfg(nb, n, ng), ref.make_bc(mask=(0,) * 6, force=(3.5, -0.25, 0.125)), q.ravel(), rhs0.ravel())
    qa, ra = sp.grid_array.from_host(grid, q), sp.grid_array.from_host(grid, rhs0)
    sp.source_term(qa, ra, sp.body_force_t(3.5, -0.25, 0.125))
    assert np.array_equal(ra.to_host().ravel(), want)


@pytest.mark.parametrize("integ", [0, 1, 2])
def test_channel_trajectory_matches_oracle(integ):
    """x/z periodic, isothermal no-slip walls in y, body force along x; the callbacks are the ones a SPADE channel
    solver passes to integrator_t (bc = exchange + boundary_fill, rhs = flux_div + source_term)."""
    from oracle import port, ref
    nb, n, ng = (2, 2, 2), (16, 8, 8), 2
    periodic = (1, 0, 1)
    sp, blocks, grid = product_setup(nb, n, ng)
    cfg = oracle_cfg(nb, n, ng, scheme=0, integrator=integ, periodic=periodic)
    bc = ref.make_bc(mask=(0, 0, 1, 1, 0, 0), a=(1, -1, -1, -1, -1), b=(0, 2 * WALL_T, 0, 0, 0), force=(40.0, 0.0, 0.0))
    q0 = make_state(nb, n, ng, seed=43)
    q0 = port.boundary_fill(cfg, bc, port.exchange(cfg, q0.ravel())).reshape(q0.shape)
    dt = 0.2 * (2 * np.pi / 32) / port.reduce_umax(cfg, q0.ravel())
    want = port.advance_channel(cfg, bc, q0.ravel(), dt, 4).reshape(q0.shape)

    gas = sp.ideal_gas_t(GAMMA, RGAS)
    flux = sp.flux_desc(product_flux(0))
    qa, ra = sp.grid_array.from_host(grid, q0), sp.grid_array(grid, 0.0)
    ex = sp.make_exchange(qa, periodic)
    wall, force = sp.noslip_isothermal_wall(WALL_T), sp.body_force_t(40.0)

    def calc_rhs(r, qq, t):
        sp.flux_div(qq, r, flux, sp.overwrite)
        sp.source_term(qq, r, force)

    def boundary_cond(qq, t):
        ex.exchange(qq)
        sp.boundary_fill(qq, sp.boundary.ymin | sp.boundary.ymax, wall)

    alg = {0: sp.rk4_t, 1: sp.ssprk3_opt, 2: sp.ssprk3_t}[integ]
    ti = sp.integrator_t(sp.time_axis_t(0.0, dt), alg, sp.integrator_data_t(qa, ra, alg), calc_rhs, boundary_cond, sp.state_transform_t(gas))
    for _ in range(4):
        ti.advance()
    got = ti.solution().to_host()
    assert rel_l2(got, want) < 1e-12
    assert rel_l2(got - q0, want - q0) < 1e-9


@pytest.mark.parametrize("scheme", [0, 1])
def test_wall_bounded_fused_stage_path_matches_oracle(scheme):
    """exchange_bc_t(handle, boundaries, wall): the boundary callback of a channel solver as an object, so that integrator_t
    keeps its one-kernel-per-stage path (narrow kernel + fused same-rank ghosts for scheme 0, wide kernel for the hybrid
    scheme) and fills the walls after the exchange like `handle.exchange(q); boundary_fill(q, ymin || ymax, wall)`."""
    from oracle import port, ref
    nb, n, ng = (2, 2, 2), (16, 8, 8), 2
    periodic = (1, 0, 1)
    sp, blocks, grid = product_setup(nb, n, ng)
    cfg = oracle_cfg(nb, n, ng, scheme=scheme, integrator=0, periodic=periodic)
    bc = ref.make_bc(mask=(0, 0, 1, 1, 0, 0), a=(1, -1, -1, -1, -1), b=(0, 2 * WALL_T, 0, 0, 0), force=(0.0, 0.0, 0.0))
    q0 = make_state(nb, n, ng, seed=47, jump=scheme == 1)
    q0 = port.boundary_fill(cfg, bc, port.exchange(cfg, q0.ravel())).reshape(q0.shape)
    dt = 0.2 * (2 * np.pi / 32) / port.reduce_umax(cfg, q0.ravel())
    want = port.advance_channel(cfg, bc, q0.ravel(), dt, 3).reshape(q0.shape)
    gas = sp.ideal_gas_t(GAMMA, RGAS)
    qa, ra = sp.grid_array.from_host(grid, q0), sp.grid_array(grid, 0.0)
    ex = sp.make_exchange(qa, periodic)
    bcs = sp.exchange_bc_t(ex, sp.boundary.ymin | sp.boundary.ymax, sp.noslip_isothermal_wall(WALL_T))
    ti = sp.integrator_t(sp.time_axis_t(0.0, dt), sp.rk4_t, sp.integrator_data_t(qa, ra, sp.rk4_t),
                         sp.flux_div_rhs_t(product_flux(scheme), sp.overwrite), bcs, sp.state_transform_t(gas))
    assert ti._plan is not None
    for _ in range(3):
        ti.advance()
    got = ti.solution().to_host()
    assert rel_l2(got, want) < 1e-12
    assert rel_l2(got - q0, want - q0) < 1e-9


@pytest.mark.parametrize("stretched", [False, True])
def test_config1_channel_at_full_size(stretched):
    """BASELINE.json configs[0] at its full size: 4x4x4 blocks of 16^3 cells, 2 exchange cells, x/z periodic, isothermal
    no-slip walls in y, body force, totani_lr + visc_lr, rk4, 10 steps (6 on the tanh-stretched grid) against the oracle."""
    from oracle import port, ref
    nb, n, ng = (4, 4, 4), (16, 16, 16), 2
    periodic = (1, 0, 1)
    bounds = [0.0, 2 * np.pi, -1.0, 1.0, 0.0, np.pi]
    nsteps = 6 if stretched else 10
    cfg = oracle_cfg(nb, n, ng, scheme=0, integrator=0, periodic=periodic, bounds=bounds)
    bc = ref.make_bc(mask=(0, 0, 1, 1, 0, 0), a=(1, -1, -1, -1, -1), b=(0, 2 * WALL_T, 0, 0, 0), force=(40.0, 0.0, 0.0))
    q0 = make_state(nb, n, ng, seed=101, bounds=bounds)
    q0 = port.boundary_fill(cfg, bc, port.exchange(cfg, q0.ravel())).reshape(q0.shape)
    maps = (None, ("tanh", -1.0, 1.0, 0.1, 4.0), None)
    dt = (0.05 if stretched else 0.2) * (2.0 / 64) / port.reduce_umax(cfg, q0.ravel())
    if stretched:
        port.set_coords(ref.make_coords(maps))
    try:
        want = port.advance_channel(cfg, bc, q0.ravel(), dt, nsteps).reshape(q0.shape)
    finally:
        port.set_coords(None)
    import spade_b200.api as sp
    coords = sp.diagonal_coords(None, sp.integrated_tanh_1D(-1.0, 1.0, 0.1, 4.0), None) if stretched else sp.identity()
    grid = sp.cartesian_grid_t(n, sp.cartesian_blocks_t(nb, bounds), coords, sp.pool_t(0, 1))
    gas = sp.ideal_gas_t(GAMMA, RGAS)
    flux = sp.flux_desc(product_flux(0))
    qa, ra = sp.grid_array.from_host(grid, q0), sp.grid_array(grid, 0.0)
    ex = sp.make_exchange(qa, periodic)
    wall, force = sp.noslip_isothermal_wall(WALL_T), sp.body_force_t(40.0)

    def calc_rhs(r, qq, t):
        sp.flux_div(qq, r, flux, sp.overwrite)
        sp.source_term(qq, r, force)

    ti = sp.integrator_t(sp.time_axis_t(0.0, dt), sp.rk4_t, sp.integrator_data_t(qa, ra, sp.rk4_t), calc_rhs,
                         sp.exchange_bc_t(ex, sp.boundary.ymin | sp.boundary.ymax, wall), sp.state_transform_t(gas))
    for _ in range(nsteps):
        ti.advance()
    got = ti.solution().to_host()
    assert rel_l2(got, want) < 1e-12
    assert rel_l2(got - q0, want - q0) < 1e-9
