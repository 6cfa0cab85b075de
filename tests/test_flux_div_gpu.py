"""Parity of the CUDA RHS kernel (through the C ABI) with the oracle restatement of
pde_algs::flux_div / flux_div_basic (reference src/pde-algs/flux-div/flux_div_basic.h:17-77).
Gate (BASELINE.json north_star): 1e-12 relative L2 in fp64 over the whole 5-vector field."""
import numpy as np
import pytest

from util import make_state, oracle_cfg, product_flux, product_setup, rel_l2, interior

pytestmark = pytest.mark.gpu
TOL = 1e-12


def run_product(nb, n, ng, q, scheme, increment=False, rhs0=None, bounds=None, **kw):
    sp, blocks, grid = product_setup(nb, n, ng, bounds)
    qa = sp.grid_array.from_host(grid, q, (ng,) * 3)
    ra = sp.grid_array(grid, 0.0, (ng,) * 3) if rhs0 is None else sp.grid_array.from_host(grid, rhs0, (ng,) * 3)
    sp.flux_div(qa, ra, product_flux(scheme, **kw), sp.increment if increment else sp.overwrite)
    return ra.to_host()


@pytest.mark.parametrize("scheme", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 17, 18, 19])
def test_schemes_match_oracle(scheme):
    from oracle import port
    nb, n, ng = (2, 1, 2), (32, 16, 8), 2
    q = make_state(nb, n, ng, seed=scheme, jump=scheme in (1, 6, 8, 10, 12, 19))
    cfg = oracle_cfg(nb, n, ng, scheme=scheme)
    want = port.flux_div(cfg, q.ravel()).reshape(q.shape)
    got = run_product(nb, n, ng, q, scheme)
    assert rel_l2(got, want) < TOL
    # ghost cells of rhs are not touched (they started at zero like the reference's overwrite)
    assert np.array_equal(got[:, 0], np.zeros_like(got[:, 0]))


@pytest.mark.parametrize("n", [(16, 16, 16), (40, 12, 4), (8, 20, 6), (64, 8, 8), (4, 4, 4)])
def test_ragged_block_shapes(n):
    from oracle import port
    nb, ng = (1, 2, 1), 2
    for scheme in (0, 1):
        q = make_state(nb, n, ng, seed=7)
        cfg = oracle_cfg(nb, n, ng, scheme=scheme)
        want = port.flux_div(cfg, q.ravel()).reshape(q.shape)
        got = run_product(nb, n, ng, q, scheme)
        assert rel_l2(got, want) < TOL


def test_anisotropic_spacing_and_one_ghost():
    from oracle import port
    nb, n, ng = (2, 2, 1), (16, 8, 12), 1
    bounds = [0.0, 6.0, -1.0, 1.0, 0.0, 3.0]
    q = make_state(nb, n, ng, seed=3, bounds=bounds)
    cfg = oracle_cfg(nb, n, ng, scheme=0, bounds=bounds)
    want = port.flux_div(cfg, q.ravel()).reshape(q.shape)
    got = run_product(nb, n, ng, q, 0, bounds=bounds)
    assert rel_l2(got, want) < TOL


def test_lattice_far_from_the_origin_keeps_the_uniform_kernel_and_the_gate():
    """A rank of a weak-scaled box sits far from the origin (z up to 2 pi N): box.size/num_cell formed from the rounded block
    bounds (cartesian_grid.h:134-135) then wobbles by 10-80 ulp from block to block. The narrow kernel treats such a lattice
    as uniform (one spacing from the constant bank: the fast variant; spb_grid::spacing_round_tol) and must stay inside
    the 1e-12 gate against the oracle, which uses the reference's per-block spacings."""
    from oracle import port
    nb, n, ng = (1, 1, 16), (16, 8, 8), 2
    L = 2 * np.pi
    bounds = [0.0, L, 0.0, L, 2 * L, 3 * L]                  # the z-slab of rank 2 of a weak-scaled TGV box: 20 ulp of wobble
    bsize = (bounds[5] - bounds[4]) / nb[2]
    dz = np.array([((bounds[4] + b * bsize + bsize) - (bounds[4] + b * bsize)) / n[2] for b in range(nb[2])])
    assert np.ptp(dz) / dz[0] > 8 * 2.220446049250313e-16    # the case the old 8-ulp test sent to the slow per-block variant
    for scheme in (0, 3, 4):
        q = make_state(nb, n, ng, seed=5, bounds=bounds)
        cfg = oracle_cfg(nb, n, ng, scheme=scheme, bounds=bounds)
        want = port.flux_div(cfg, q.ravel()).reshape(q.shape)
        got = run_product(nb, n, ng, q, scheme, bounds=bounds)
        assert rel_l2(got, want) < TOL


def test_increment_trait():
    from oracle import port
    nb, n, ng = (1, 1, 2), (32, 8, 8), 2
    q = make_state(nb, n, ng, seed=11)
    rng = np.random.default_rng(5)
    rhs0 = rng.normal(size=q.shape) * 1e3
    cfg = oracle_cfg(nb, n, ng, scheme=0)
    want = port.flux_div(cfg, q.ravel(), rhs=rhs0.ravel(), increment=True).reshape(q.shape)
    got = run_product(nb, n, ng, q, 0, increment=True, rhs0=rhs0)
    assert rel_l2(got, want) < TOL


def test_uniform_state_has_zero_rhs_and_linearity_of_block_range():
    sp, blocks, grid = product_setup((2, 2, 2), (32, 32, 32), 2)
    q = np.zeros((8, 36, 36, 36, 5))
    q[..., 0], q[..., 1], q[..., 2], q[..., 3], q[..., 4] = 101325.0, 300.0, 10.0, -3.0, 2.0
    qa = sp.grid_array.from_host(grid, q)
    ra = sp.grid_array(grid, 1.0)
    sp.flux_div(qa, ra, product_flux(1), sp.overwrite)
    r = interior(ra.to_host(), 2)
    assert np.abs(r).max() < 1e-6 * 101325.0 * 10.0   # pure cancellation
    # block-range launches tile the full launch exactly
    qs = make_state((2, 2, 2), (32, 32, 32), 2, seed=1)
    qa = sp.grid_array.from_host(grid, qs)
    full = sp.grid_array(grid, 0.0)
    parts = sp.grid_array(grid, 0.0)
    sp.flux_div(qa, full, product_flux(0), sp.overwrite)
    sp.flux_div(qa, parts, product_flux(0), sp.overwrite, blocks=(0, 3))
    sp.flux_div(qa, parts, product_flux(0), sp.overwrite, blocks=(3, 8))
    assert np.array_equal(full.to_host(), parts.to_host())


def test_matches_reference_library_if_present(ref_lib):
    nb, n, ng = (2, 2, 2), (16, 16, 16), 2
    q = make_state(nb, n, ng, seed=21, jump=True)
    for scheme in (0, 1):
        cfg = oracle_cfg(nb, n, ng, scheme=scheme)
        want = ref_lib.flux_div(cfg, q.ravel()).reshape(q.shape)
        got = run_product(nb, n, ng, q, scheme)
        assert rel_l2(got, want) < TOL


@pytest.mark.parametrize("scheme,ng", [(13, 3), (14, 4), (15, 3), (16, 4), (13, 4), (0, 3), (1, 4)])
@pytest.mark.parametrize("n", [(32, 16, 8), (40, 12, 6)])
def test_cent_keep_6_8_and_deeper_exchange_layers(scheme, ng, n):
    """cent_keep<6> / cent_keep<8> (convective.h:97-192; 3 / 4 exchange cells, 6- / 8-cell stencils, 7 / 9 staged planes), and the
    two-cell schemes on arrays that carry more exchange cells than they need."""
    from oracle import port
    nb = (2, 1, 2)
    q = make_state(nb, n, ng, seed=scheme, jump=scheme == 1)
    cfg = oracle_cfg(nb, n, ng, scheme=scheme)
    want = port.flux_div(cfg, q.ravel()).reshape(q.shape)
    got = run_product(nb, n, ng, q, scheme)
    assert rel_l2(got, want) < TOL


@pytest.mark.parametrize("scheme,ng", [(13, 3), (14, 4)])
def test_cent_keep_6_8_rk4_trajectory(scheme, ng):
    from oracle import port
    from util import GAMMA, RGAS
    nb, n = (2, 2, 1), (16, 8, 8)
    cfg = oracle_cfg(nb, n, ng, scheme=scheme, integrator=0)
    q0 = port.exchange(cfg, make_state(nb, n, ng, seed=5).ravel())
    dt = 0.2 * (2 * np.pi / 32) / port.reduce_umax(cfg, q0)
    want = port.advance(cfg, q0, dt, 2)
    sp, blocks, grid = product_setup(nb, n, ng)
    shape = (-1, n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng, 5)
    for fused in (True, False):
        qa = sp.grid_array.from_host(grid, q0.reshape(shape), (ng,) * 3)
        ra = sp.grid_array(grid, 0.0, (ng,) * 3)
        ex = sp.make_exchange(qa, (1, 1, 1))
        ti = sp.integrator_t(sp.time_axis_t(0.0, dt), sp.rk4_t, sp.integrator_data_t(qa, ra, sp.rk4_t),
                             sp.flux_div_rhs_t(product_flux(scheme), sp.overwrite), sp.exchange_bc_t(ex),
                             sp.state_transform_t(sp.ideal_gas_t(GAMMA, RGAS)), fused=fused)
        assert (ti._plan is not None) == fused
        ti.advance()
        ti.advance()
        assert rel_l2(ti.solution().to_host().ravel(), want) < TOL
