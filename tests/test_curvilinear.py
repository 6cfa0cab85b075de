"""General coordinates (coords::diagonal_coords, BASELINE config 3: stretched channel grid).

What pins what:
  * convective functors (totani_lr, fweno_t, cent_keep<4>): the oracle restatement is bit-exact against the reference's
    own flux_div(basic) on stretched grids (oracle/_ref/libspade_ref_curv.so: the reference compiled with the
    two-declaration repair of core/coord_system.h:255,274, see oracle/ref_driver_curv.cc) and against the committed
    vectors tests/golden/curvilinear_small.npz generated from it; the CUDA path is compared with both to 1e-12.
  * viscous / sensor terms: the reference has no gradient on general coordinates (info_gradient.h:83) — PARITY UNPINNED.
    The completion d/dx_d = (1/m_d) d/dxi_d is checked for consistency instead: a uniformly scaled coordinate must
    reproduce the identity-coordinate result on the scaled box (which IS pinned), and the viscous RHS on a tanh-stretched
    grid converges at second order to the analytic one.
Reference paths: flux_div_basic.h:49-71, coord_system.h:65-177,250-267,295-302, info_metric.h:24-32, source_term.h:38-46."""
import os

import numpy as np
import pytest

from util import GAMMA, RGAS, make_state, oracle_cfg, product_flux, rel_l2, interior

TOL = 1e-12
HERE = os.path.dirname(os.path.abspath(__file__))
NB, N, NG = (2, 2, 1), (8, 4, 4), 2
BOUNDS = [0.0, 2 * np.pi, -1.0, 1.0, 0.5, 2.0]
MAPS = {"channel": (("scaled", 2.0), ("tanh", -1.0, 1.0, 0.1, 1.3), None),
        "tanh_quad": (None, ("tanh", -1.0, 1.0, 0.1, 1.3), ("quad",))}


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "curvilinear_small.npz"))


@pytest.fixture()
def coords_off():
    from oracle import port
    yield
    port.set_coords(None)


def oracle_flux_div(cfg, cd, q, **kw):
    from oracle import port
    port.set_coords(cd)
    try:
        return port.flux_div(cfg, np.ascontiguousarray(q).ravel(), **kw).reshape(q.shape)
    finally:
        port.set_coords(None)


# ---------------------------------------------------------------- CPU: the oracle against the reference ----
@pytest.mark.parametrize("name", list(MAPS))
@pytest.mark.parametrize("scheme", [3, 5, 7])
def test_oracle_matches_golden_reference_vectors(golden, name, scheme):
    from oracle import ref
    q = make_state(NB, N, NG, seed=200 + scheme, bounds=BOUNDS)
    cfg = oracle_cfg(NB, N, NG, scheme=scheme, bounds=BOUNDS)
    got = oracle_flux_div(cfg, ref.make_coords(MAPS[name]), q)
    assert np.array_equal(got, golden[f"{name}_rhs_{scheme}"])


@pytest.mark.parametrize("name", list(MAPS))
def test_oracle_geometry_matches_reference_tables(golden, name):
    """calc_jacobian at the computational centre, info::metric at the MAPPED centre (info_metric.h:31) — block 3."""
    from oracle import port, ref
    cd = ref.make_coords(MAPS[name])
    nbx, nby = NB[0], NB[1]
    b = (3 % nbx, (3 // nbx) % nby, 3 // (nbx * nby))
    xc = []
    for d in range(3):
        bsize = (BOUNDS[2 * d + 1] - BOUNDS[2 * d]) / NB[d]
        lo = BOUNDS[2 * d] + b[d] * bsize
        dx = (lo + bsize - lo) / N[d]
        xc.append([lo + (i + 0.5) * dx for i in range(N[d])])
    jac, nrm, xyz = golden[f"{name}_jac"], golden[f"{name}_nrm"], golden[f"{name}_xyz"]
    for k in range(N[2]):
        for j in range(N[1]):
            for i in range(N[0]):
                c = (xc[0][i], xc[1][j], xc[2][k])
                m = [port.coord_deriv(cd, d, c[d]) for d in range(3)]
                assert jac[k, j, i] == 1.0 / (m[0] * m[1] * m[2])
                x = [port.coord_map(cd, d, c[d]) for d in range(3)]
                assert list(xyz[k, j, i]) == x
                mp = [port.coord_deriv(cd, d, x[d]) for d in range(3)]
                jp = 1.0 / (mp[0] * mp[1] * mp[2])
                assert list(nrm[k, j, i]) == [1.0 / (mp[d] * jp) for d in range(3)]


@pytest.mark.parametrize("scheme", [3, 5, 7])
def test_oracle_matches_reference_library_if_present(scheme):
    from oracle import ref
    if not ref.curv_available():
        pytest.skip("oracle/_ref/libspade_ref_curv.so not built (needs /root/reference)")
    nb, n = (1, 2, 2), (6, 8, 4)
    for name, maps in MAPS.items():
        cd = ref.make_coords(maps)
        q = make_state(nb, n, NG, seed=scheme, bounds=BOUNDS)
        cfg = oracle_cfg(nb, n, NG, scheme=scheme, bounds=BOUNDS)
        rhs0 = np.random.default_rng(9).normal(size=q.size) * 1e3
        for kw in (dict(), dict(rhs=rhs0, increment=True)):
            want = ref.curv_flux_div(cfg, cd, q.ravel(), **kw).reshape(q.shape)
            assert np.array_equal(oracle_flux_div(cfg, cd, q, **kw), want)
        for d in range(3):
            for x in (0.31, -0.7, 1.9):
                assert port_map(cd, d, x) == ref.curv_map(cd, d, x)


def port_map(cd, d, x):
    from oracle import port
    assert port.coord_deriv(cd, d, x) == __import__("oracle.ref", fromlist=["ref"]).curv_deriv(cd, d, x)
    return port.coord_map(cd, d, x)


def test_oracle_rk4_on_stretched_grid_matches_golden_trajectory(golden, coords_off):
    from oracle import port, ref
    cfg = oracle_cfg(NB, N, NG, scheme=3, integrator=0, bounds=BOUNDS)
    q0 = port.exchange(cfg, make_state(NB, N, NG, seed=77, bounds=BOUNDS).ravel())
    port.set_coords(ref.make_coords(MAPS["channel"]))
    q2 = port.advance(cfg, q0, float(golden["adv_dt"][0]), 2)
    assert rel_l2(q2, golden["adv_q2"].ravel()) < 1e-14


@pytest.mark.parametrize("scheme", [0, 2, 4])
def test_scaled_coordinate_equals_identity_on_the_scaled_box(scheme):
    """x = k xi with constant k: the curvilinear path (Jacobian, metric vectors, gradient transform) must reproduce the
    identity-coordinate result on the box stretched by k — which is pinned against the reference. This is the check of
    the gradient transform (the sensor of the hybrid schemes reads the same transformed gradient). The fweno_t schemes
    are excluded on purpose: the reference scales their flux part by the metric but not the Rusanov dissipation
    (convective.h:363-378), so on general coordinates they are NOT equivalent to the scaled box — that behaviour is
    pinned by the reference vectors above instead."""
    from oracle import ref
    nb, n = (2, 1, 1), (8, 6, 4)
    k = (2.0, 0.5, 3.0)
    b_comp = [0.0, 1.0, -1.0, 1.0, 0.25, 1.25]
    b_phys = [b_comp[2 * d + s] * k[d] for d in range(3) for s in range(2)]
    q = make_state(nb, n, NG, seed=31 + scheme, bounds=b_phys, jump=False)
    want = oracle_flux_div(oracle_cfg(nb, n, NG, scheme=scheme, bounds=b_phys), None, q)
    for phys in (True, False):
        cd = ref.make_coords([("scaled", kk) for kk in k], metric_at_physical=phys)
        got = oracle_flux_div(oracle_cfg(nb, n, NG, scheme=scheme, bounds=b_comp), cd, q)
        assert rel_l2(got, want) < 1e-13


def shear_state(nb, n, ng, bounds, cd):
    """u = U sin(y), p and T uniform, y the PHYSICAL wall-normal coordinate of a tanh-stretched grid."""
    from oracle import port
    nlb = nb[0] * nb[1] * nb[2]
    q = np.zeros((nlb, n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng, 5))
    ys = []
    for lb in range(nlb):
        bj = (lb // nb[0]) % nb[1]
        bsize = (bounds[3] - bounds[2]) / nb[1]
        lo = bounds[2] + bj * bsize
        dx = bsize / n[1]
        y = np.array([port.coord_map(cd, 1, lo + (j + 0.5) * dx) for j in range(-ng, n[1] + ng)])
        q[lb, ..., 0], q[lb, ..., 1] = 101325.0, 300.0
        q[lb, ..., 2] = 30.0 * np.sin(y)[None, :, None]
        ys.append(y[ng:-ng])
    return q, ys


def test_viscous_rhs_on_tanh_grid_converges_at_second_order():
    """visc_lr alone on y = tanh-stretched: rhs_xmom -> mu U d2/dy2 sin y, rhs_energy -> mu U^2 d/dy(sin y cos y)."""
    from oracle import ref
    mu, U = 1.0e-2, 30.0
    cd = ref.make_coords((None, ("tanh", -1.0, 1.0, 0.1, 4.0), None), metric_at_physical=False)
    errs = []
    for ny in (16, 32, 64):
        nb, n = (1, 2, 1), (4, ny // 2, 4)
        bounds = [0.0, 1.0, -1.0, 1.0, 0.0, 1.0]
        q, ys = shear_state(nb, n, NG, bounds, cd)
        r = interior(oracle_flux_div(oracle_cfg(nb, n, NG, scheme=4, mu=mu, bounds=bounds), cd, q), NG)
        e = 0.0
        for lb in range(2):
            y = ys[lb]
            e = max(e, np.abs(r[lb, 0, :, 0, 2] - (-mu * U * np.sin(y))).max() / (mu * U))
            e = max(e, np.abs(r[lb, 0, :, 0, 1] - mu * U * U * np.cos(2 * y)).max() / (mu * U * U))
        errs.append(e)
    assert errs[0] / errs[1] > 3.3 and errs[1] / errs[2] > 3.3 and errs[2] < 2e-3


# ---------------------------------------------------------------- GPU: the CUDA path through the C ABI ----
def product_coords(sp, maps, metric_at="physical"):
    def one(m):
        if m is None:
            return None
        return {"scaled": sp.scaled_coord_1D, "tanh": sp.integrated_tanh_1D, "quad": sp.quad_1D}[m[0]](*m[1:])
    return sp.diagonal_coords(*[one(m) for m in maps], metric_at=metric_at)


def product_grid(nb, n, bounds, maps, metric_at="physical"):
    import spade_b200.api as sp
    blocks = sp.cartesian_blocks_t(nb, bounds)
    return sp, sp.cartesian_grid_t(n, blocks, product_coords(sp, maps, metric_at), sp.pool_t(0, 1))


def run_product(nb, n, ng, q, scheme, bounds, maps, metric_at="physical", increment=False, rhs0=None):
    sp, grid = product_grid(nb, n, bounds, maps, metric_at)
    qa = sp.grid_array.from_host(grid, q, (ng,) * 3)
    ra = sp.grid_array(grid, 0.0, (ng,) * 3) if rhs0 is None else sp.grid_array.from_host(grid, rhs0, (ng,) * 3)
    sp.flux_div(qa, ra, product_flux(scheme), sp.increment if increment else sp.overwrite)
    return ra.to_host()


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(MAPS))
@pytest.mark.parametrize("scheme", [3, 5, 7])
def test_cuda_convective_flux_div_matches_reference_vectors(golden, name, scheme):
    q = make_state(NB, N, NG, seed=200 + scheme, bounds=BOUNDS)
    got = run_product(NB, N, NG, q, scheme, BOUNDS, MAPS[name])
    assert rel_l2(got, golden[f"{name}_rhs_{scheme}"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("metric_at", ["physical", "computational"])
@pytest.mark.parametrize("scheme", list(range(9)) + [11, 12])
def test_cuda_all_schemes_match_oracle_on_stretched_grid(scheme, metric_at):
    from oracle import ref
    nb, n = (2, 2, 1), (32, 16, 8)
    q = make_state(nb, n, NG, seed=scheme, bounds=BOUNDS, jump=scheme in (1, 6, 8, 12))
    cd = ref.make_coords(MAPS["tanh_quad"], metric_at_physical=metric_at == "physical")
    want = oracle_flux_div(oracle_cfg(nb, n, NG, scheme=scheme, bounds=BOUNDS), cd, q)
    got = run_product(nb, n, NG, q, scheme, BOUNDS, MAPS["tanh_quad"], metric_at)
    assert rel_l2(got, want) < TOL
    assert np.array_equal(got[:, 0], np.zeros_like(got[:, 0]))


@pytest.mark.gpu
@pytest.mark.parametrize("n", [(16, 16, 16), (40, 12, 4), (8, 20, 6)])
def test_cuda_ragged_blocks_and_increment_on_stretched_grid(n):
    from oracle import ref
    nb = (1, 2, 1)
    cd = ref.make_coords(MAPS["channel"])
    for scheme in (0, 1):
        q = make_state(nb, n, NG, seed=7, bounds=BOUNDS)
        rhs0 = np.random.default_rng(5).normal(size=q.shape) * 1e3
        cfg = oracle_cfg(nb, n, NG, scheme=scheme, bounds=BOUNDS)
        want = oracle_flux_div(cfg, cd, q, rhs=rhs0.ravel(), increment=True)
        got = run_product(nb, n, NG, q, scheme, BOUNDS, MAPS["channel"], increment=True, rhs0=rhs0)
        assert rel_l2(got, want) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("scheme,fused", [(0, True), (0, False), (1, True), (3, True)])
def test_cuda_rk4_trajectory_on_stretched_grid(scheme, fused, coords_off):
    from oracle import port, ref
    nb, n = (2, 2, 2), (16, 16, 16)
    cfg = oracle_cfg(nb, n, NG, scheme=scheme, integrator=0, bounds=BOUNDS)
    q0 = port.exchange(cfg, make_state(nb, n, NG, seed=3, bounds=BOUNDS).ravel())
    dt = 2e-5
    port.set_coords(ref.make_coords(MAPS["channel"]))
    want = port.advance(cfg, q0, dt, 2)
    port.set_coords(None)
    sp, grid = product_grid(nb, n, BOUNDS, MAPS["channel"])
    gas = sp.ideal_gas_t(GAMMA, RGAS)
    qa = sp.grid_array.from_host(grid, q0.reshape((-1, n[2] + 2 * NG, n[1] + 2 * NG, n[0] + 2 * NG, 5)))
    ra = sp.grid_array(grid, 0.0)
    ex = sp.make_exchange(qa, (1, 1, 1))
    data = sp.integrator_data_t(qa, ra, sp.rk4_t)
    rhs = sp.flux_div_rhs_t(product_flux(scheme), sp.overwrite)
    ti = sp.integrator_t(sp.time_axis_t(0.0, dt), sp.rk4_t, data, rhs, sp.exchange_bc_t(ex), sp.state_transform_t(gas), fused=fused)
    n0 = sp.launch_count()
    ti.advance()
    ti.advance()
    assert rel_l2(ti.solution().to_host().ravel(), want) < TOL
    if fused:
        # ONE kernel per stage on stretched grids too: RHS + update + same-rank ghosts (the ghost warp of the narrow kernel for the
        # one-ghost-cell functors, the owning threads of the wide kernel for the others)
        assert sp.launch_count() - n0 == 2 * 4


@pytest.mark.gpu
def test_cuda_source_term_divides_by_the_jacobian(coords_off):
    from oracle import port, ref
    nb, n = (2, 2, 1), (8, 4, 4)
    q = make_state(nb, n, NG, seed=51, bounds=BOUNDS)
    rhs0 = np.random.default_rng(5).standard_normal(q.shape)
    force = (40.0, -0.5, 0.25)
    cfg = oracle_cfg(nb, n, NG, periodic=(0, 0, 0), bounds=BOUNDS)
    port.set_coords(ref.make_coords(MAPS["channel"]))
    want = port.source_term(cfg, ref.make_bc(mask=(0,) * 6, force=force), q.ravel(), rhs0.ravel()).reshape(q.shape)
    port.set_coords(None)
    sp, grid = product_grid(nb, n, BOUNDS, MAPS["channel"])
    qa, ra = sp.grid_array.from_host(grid, q), sp.grid_array.from_host(grid, rhs0)
    sp.source_term(qa, ra, sp.body_force_t(*force))
    assert rel_l2(ra.to_host(), want) < 1e-14


def test_python_coordinate_mirror_matches_the_oracle_mappings():
    """spade_b200.api.diagonal_coords / integrated_tanh_1D / scaled_coord_1D / quad_1D (numpy) against the C restatement of
    core/coord_system.h:54-177, and the metric tables handed to spb_grid_set_metric against the oracle's geometry (CPU only)."""
    import spade_b200.api as sp
    from oracle import port, ref
    cd = ref.make_coords((("scaled", 2.0), ("tanh", -1.0, 1.0, 0.1, 1.3), ("quad",)))
    maps = [sp.scaled_coord_1D(2.0), sp.integrated_tanh_1D(-1.0, 1.0, 0.1, 1.3), sp.quad_1D()]
    xs = np.linspace(-1.1, 1.9, 41)
    for d, m in enumerate(maps):
        want_x = np.array([port.coord_map(cd, d, x) for x in xs])
        want_m = np.array([port.coord_deriv(cd, d, x) for x in xs])
        assert np.allclose(m.map(xs), want_x, rtol=1e-14, atol=1e-15)
        assert np.allclose(m.coord_deriv(xs), want_m, rtol=1e-14, atol=1e-15)
    nb, n, ng = (2, 2, 1), (8, 4, 4), 2
    for metric_at in ("physical", "computational"):
        grid = sp.cartesian_grid_t(n, sp.cartesian_blocks_t(nb, BOUNDS), sp.diagonal_coords(*maps, metric_at=metric_at), sp.pool_t(0, 1))
        area, jac, face = grid.metric_tables((ng,) * 3)
        for d in range(3):
            assert area[d].shape == (4, n[d] + 2 * ng) and face[d].shape == (4, n[d] + 2 * ng + 1)
            lb = 3
            b = (lb % nb[0], (lb // nb[0]) % nb[1], lb // (nb[0] * nb[1]))
            bsize = (BOUNDS[2 * d + 1] - BOUNDS[2 * d]) / nb[d]
            lo = BOUNDS[2 * d] + b[d] * bsize
            dx = (lo + bsize - lo) / n[d]
            for i in range(-ng, n[d] + ng):
                xc = lo + (i + 0.5) * dx
                assert np.isclose(jac[d][lb, i + ng], port.coord_deriv(cd, d, xc), rtol=1e-14)
                xa = port.coord_map(cd, d, xc) if metric_at == "physical" else xc
                assert np.isclose(area[d][lb, i + ng], port.coord_deriv(cd, d, xa), rtol=1e-13)
                assert np.isclose(face[d][lb, i + ng], port.coord_deriv(cd, d, lo + i * dx), rtol=1e-14)


@pytest.mark.gpu
def test_cuda_mapping_that_degenerates_beyond_the_wall_is_accepted():
    """quad_1D has dz/dzeta = 2 zeta = 0 at zeta = 0: with 4 exchange cells of width 0.125 below zeta = 0.5 the outermost ghost
    FACE sits exactly there (and integrated_tanh_1D folds beyond |eta| = 1.2). Only the entries of interior cells and their
    faces are read by the kernels, so the grid must be accepted and cent_keep<8> + visc_lr must still match the oracle."""
    from oracle import ref
    nb, n, ng = (1, 1, 2), (12, 20, 6), 4
    bounds = [0.0, 2 * np.pi, -1.0, 1.0, 0.5, 2.0]
    maps = (("scaled", 2.0), ("tanh", -1.0, 1.0, 0.1, 4.0), ("quad",))
    q = make_state(nb, n, ng, seed=14, bounds=bounds)
    want = oracle_flux_div(oracle_cfg(nb, n, ng, scheme=14, bounds=bounds), ref.make_coords(maps), q)
    got = run_product(nb, n, ng, q, 14, bounds, maps)
    assert rel_l2(got, want) < TOL
