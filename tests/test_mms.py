"""Manufactured-solution order check (the reference's closest thing to a known-answer test is development/mms/main.cc —
trig fields for p, T, u, v, w, analytic -div(F_conv) and div(tau), observed order over a sequence of grids; it checks orders,
not values). Independent sanity test of the physics: the oracle restatement (bit-exact against the reference, and the 1e-12
yardstick of the CUDA path) must converge to the analytic right-hand side at the design order of each functor, on uniform
and on stretched (diagonal_coords) grids. CPU only; the analytic RHS comes from sympy.

Fields as development/mms/main.cc:61-65 with alpha = 1 on the periodic box [0, 2 pi)^3 (2 pi-periodic, positive p and T)."""
import numpy as np
import pytest

from util import interior, oracle_cfg

GAMMA, RGAS, MU, PR = 1.4, 287.15, 0.5, 0.72


def build_analytic():
    import sympy as sy
    x, y, z = sy.symbols("x y z")
    p = 5.0 + 2.0 * sy.cos(3 * x) * sy.sin(2 * y) + sy.sin(4 * z)
    T = 10.0 + 2.0 * sy.cos(2 * x) * sy.sin(3 * y) + sy.sin(4 * z)
    u = sy.sin(3 * x) * sy.cos(2 * y) * sy.cos(2 * z)
    v = sy.cos(3 * x) * sy.cos(2 * y) * sy.cos(3 * z)
    w = sy.sin(3 * x) * sy.sin(2 * y) * sy.cos(4 * z)
    vel, X = [u, v, w], [x, y, z]
    rho = p / (RGAS * T)
    Et = RGAS * T / (GAMMA - 1.0) + 0.5 * (u * u + v * v + w * w)
    # convective: rhs = -div F, F_d = (rho u_d, (rho Et + p) u_d, rho u u_d + p delta_xd, ...)
    conv = [-sum(sy.diff(rho * vel[d], X[d]) for d in range(3)),
            -sum(sy.diff((rho * Et + p) * vel[d], X[d]) for d in range(3))]
    for c in range(3):
        conv.append(-sum(sy.diff(rho * vel[c] * vel[d], X[d]) for d in range(3)) - sy.diff(p, X[c]))
    # viscous (viscous.h:39-80): rhs = +div tau (momentum), +div(u.tau + cond grad T) (energy); beta = -2 mu / 3
    div = sum(sy.diff(vel[d], X[d]) for d in range(3))
    tau = [[MU * (sy.diff(vel[i], X[j]) + sy.diff(vel[j], X[i])) + (-2.0 * MU / 3.0 * div if i == j else 0) for j in range(3)] for i in range(3)]
    cond = (GAMMA * RGAS / (GAMMA - 1.0)) * (MU / PR)
    visc = [sy.Integer(0),
            sum(sy.diff(sum(vel[i] * tau[i][d] for i in range(3)) + cond * sy.diff(T, X[d]), X[d]) for d in range(3))]
    for c in range(3):
        visc.append(sum(sy.diff(tau[c][d], X[d]) for d in range(3)))
    f = lambda e: sy.lambdify((x, y, z), e, "numpy")
    return dict(state=[f(e) for e in (p, T, u, v, w)], conv=[f(e) for e in conv], visc=[f(e) for e in visc])


@pytest.fixture(scope="module")
def analytic():
    return build_analytic()


def _grid(nb, n, ng, bounds, maps):
    """physical cell-centre coordinates of every padded cell, [nlb] lists of (X, Y, Z) meshes in the array's index order"""
    from oracle import port, ref
    cd = ref.make_coords(maps, metric_at_physical=False) if maps else None
    out = []
    for lb in range(nb[0] * nb[1] * nb[2]):
        b = (lb % nb[0], (lb // nb[0]) % nb[1], lb // (nb[0] * nb[1]))
        ax = []
        for d in range(3):
            bs = (bounds[2 * d + 1] - bounds[2 * d]) / nb[d]
            lo = bounds[2 * d] + b[d] * bs
            xc = lo + (np.arange(-ng, n[d] + ng) + 0.5) * (bs / n[d])
            ax.append(np.array([port.coord_map(cd, d, t) for t in xc]) if cd is not None else xc)
        Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
        out.append((X, Y, Z))
    return cd, out


def _error(analytic, scheme, which, ncell, ng, maps=None):
    from oracle import port
    nb, n = (2, 1, 1), (ncell // 2, ncell, ncell)
    two_pi = 2 * np.pi
    # stretched case: x = 2 xi on [0, pi), y and z identity: still periodic in physical space
    bounds = [0.0, np.pi if maps else two_pi, 0.0, two_pi, 0.0, two_pi]
    cd, meshes = _grid(nb, n, ng, bounds, maps)
    q = np.zeros((2, n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng, 5))
    want = np.zeros_like(q)
    for lb, (X, Y, Z) in enumerate(meshes):
        for v in range(5):
            q[lb, ..., v] = analytic["state"][v](X, Y, Z)
            want[lb, ..., v] = np.broadcast_to(analytic[which][v](X, Y, Z), X.shape)
    cfg = oracle_cfg(nb, n, ng, scheme=scheme, mu=MU, prandtl=PR, bounds=bounds)
    port.set_coords(cd)
    try:
        got = port.flux_div(cfg, q.ravel()).reshape(q.shape)
    finally:
        port.set_coords(None)
    d = interior(got, ng) - interior(want, ng)
    return float(np.sqrt((d ** 2).mean()) / np.sqrt((interior(want, ng) ** 2).mean()))


@pytest.mark.parametrize("scheme,which,ng,order,sizes", [
    (3, "conv", 2, 2, (16, 32, 64)),      # totani_lr
    (4, "visc", 2, 2, (16, 32, 64)),      # visc_lr
    (7, "conv", 2, 4, (16, 32, 64)),      # cent_keep<4>
    (15, "conv", 3, 6, (48, 64, 96)),     # cent_keep<6>: 5.4, 5.7, 5.8 -> 6 over 32..96 cells (wavenumber-4 fields)
    (16, "conv", 4, 8, (48, 64, 96)),     # cent_keep<8>: 7.0, 7.5, 7.7 -> 8
])
def test_observed_order_on_a_uniform_grid(analytic, scheme, which, ng, order, sizes):
    errs = [_error(analytic, scheme, which, n, ng) for n in sizes]
    rates = [np.log(errs[i] / errs[i + 1]) / np.log(sizes[i + 1] / sizes[i]) for i in range(len(sizes) - 1)]
    assert rates[-1] > order - (0.35 if order <= 4 else 0.5), (errs, rates)
    assert errs[-1] < errs[0]


@pytest.mark.parametrize("scheme,which,order", [(3, "conv", 2), (4, "visc", 2), (7, "conv", 4)])
def test_observed_order_on_a_scaled_grid(analytic, scheme, which, order):
    """x = 2 xi (scaled_coord_1D): Jacobian, metric vectors and the gradient transform at work on the analytic solution"""
    sizes = (16, 32, 64)
    maps = (("scaled", 2.0), None, None)
    errs = [_error(analytic, scheme, which, n, 2, maps) for n in sizes]
    rates = [np.log(errs[i] / errs[i + 1]) / np.log(2.0) for i in range(2)]
    assert rates[-1] > order - 0.35, (errs, rates)


def test_wale_eddy_viscosity_vanishes_in_pure_shear():
    """The property WALE is built for (Nicoud & Ducros 1999; reference development/subgrid/main.cc checks the model against its
    analytic value): in a pure shear u(y) the traceless symmetric part of the squared velocity-gradient tensor is zero, so
    mu_t = 0 and visc_lr over sgs_visc_t must give exactly the laminar result; in a general field it must not."""
    from oracle import port
    from util import make_state
    nb, n, ng = (1, 2, 1), (8, 8, 8), 2
    q = np.zeros((2, 12, 12, 12, 5))
    y = np.concatenate([(np.arange(-ng, n[1] + ng) + 0.5) * (np.pi / 8) + lb * np.pi for lb in range(2)]).reshape(2, 12)
    q[..., 0], q[..., 1] = 101325.0, 300.0
    q[..., 2] = 30.0 * np.sin(y)[:, None, :, None]
    lam = port.flux_div(oracle_cfg(nb, n, ng, scheme=0), q.ravel())
    les = port.flux_div(oracle_cfg(nb, n, ng, scheme=11), q.ravel())
    assert np.array_equal(lam, les)
    q = make_state(nb, n, ng, seed=5)
    lam = port.flux_div(oracle_cfg(nb, n, ng, scheme=0), q.ravel())
    les = port.flux_div(oracle_cfg(nb, n, ng, scheme=11), q.ravel())
    assert np.linalg.norm(les - lam) > 1e-4 * np.linalg.norm(lam)


def test_ducros_sensor_switches_the_dissipative_flux_off_in_solid_rotation():
    """state_sensor::ducros_t (state_sensor.h:32-42; the reference checks it against its analytic value in
    development/shock-sense/main.cc:60-101): theta^2/(theta^2 + |omega|^2 + eps) is exactly 0 for a solid-body rotation (the
    discrete divergence of a linear field vanishes), so hybrid_scheme_t(totani, fweno, ducros, diss_flux) = F0 + 0 F1 must equal
    totani_lr + visc_lr bit for bit; a compressing field switches the WENO flux on."""
    from oracle import port
    nb, n, ng = (1, 1, 1), (8, 8, 8), 2
    ax = (np.arange(-ng, 8 + ng) + 0.5) * 0.25 - 1.0
    Z, Y, X = np.meshgrid(ax, ax, ax, indexing="ij")
    q = np.zeros((1, 12, 12, 12, 5))
    q[0, ..., 0], q[0, ..., 1] = 101325.0, 300.0
    q[0, ..., 2], q[0, ..., 3] = -8.0 * Y, 8.0 * X
    bounds = [-1.0, 1.0] * 3
    central = port.flux_div(oracle_cfg(nb, n, ng, scheme=0, bounds=bounds), q.ravel())
    hybrid = port.flux_div(oracle_cfg(nb, n, ng, scheme=8, bounds=bounds), q.ravel())
    assert np.array_equal(central, hybrid)
    q[0, ..., 2], q[0, ..., 3], q[0, ..., 4] = -8.0 * X, -8.0 * Y, -8.0 * Z          # pure compression: sensor ~ 1
    central = port.flux_div(oracle_cfg(nb, n, ng, scheme=0, bounds=bounds), q.ravel())
    hybrid = port.flux_div(oracle_cfg(nb, n, ng, scheme=8, bounds=bounds), q.ravel())
    assert np.linalg.norm(hybrid - central) > 1e-3 * np.linalg.norm(central)


# ---------------------------------------------------------------- config 3's functor set on config 3's kind of grid
TANH = (None, ("tanh", -1.0, 1.0, 0.1, 1.3), None)


def _error_tanh(analytic, scheme, ncell, ng, use_gpu):
    """relative L2 error of the TOTAL right-hand side (convective + viscous) against the analytic one on x, z periodic and
    y = integrated_tanh_1D(-1, 1, 0.1, 1.3) (consistent metric: coord_deriv at the computational cell centre); all cells
    including the exchange cells carry the analytic state, so no boundary treatment enters"""
    from oracle import port
    nb, n = (2, 1, 1), (ncell // 2, ncell, ncell)
    two_pi = 2 * np.pi
    bounds = [0.0, two_pi, -1.0, 1.0, 0.0, two_pi]
    cd, meshes = _grid(nb, n, ng, bounds, TANH)
    q = np.zeros((2, n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng, 5))
    want = np.zeros_like(q)
    for lb, (X, Y, Z) in enumerate(meshes):
        for v in range(5):
            q[lb, ..., v] = analytic["state"][v](X, Y, Z)
            want[lb, ..., v] = np.broadcast_to(analytic["conv"][v](X, Y, Z), X.shape) + np.broadcast_to(analytic["visc"][v](X, Y, Z), X.shape)
    if use_gpu:
        import spade_b200.api as sp
        gas = sp.ideal_gas_t(GAMMA, RGAS)
        vl = sp.constant_viscosity_t(MU, PR)
        conv = sp.hybrid_scheme_t(sp.totani_lr(gas), sp.fweno_t(gas), sp.ducros_t(1e-2), sp.full_flux) if scheme == 1 else sp.totani_lr(gas)
        coords = sp.diagonal_coords(None, sp.integrated_tanh_1D(-1.0, 1.0, 0.1, 1.3), None, metric_at="computational")
        grid = sp.cartesian_grid_t(n, sp.cartesian_blocks_t(nb, bounds), coords, sp.pool_t(0, 1))
        qa = sp.grid_array.from_host(grid, q, (ng,) * 3)
        ra = sp.grid_array(grid, 0.0, (ng,) * 3)
        sp.flux_div(qa, ra, sp.compose(conv, sp.visc_lr(vl, gas)), sp.overwrite)
        got = ra.to_host()
    else:
        cfg = oracle_cfg(nb, n, ng, scheme=scheme, mu=MU, prandtl=PR, bounds=bounds)
        port.set_coords(cd)
        try:
            got = port.flux_div(cfg, q.ravel()).reshape(q.shape)
        finally:
            port.set_coords(None)
    d = interior(got, ng) - interior(want, ng)
    return float(np.sqrt((d ** 2).mean()) / np.sqrt((interior(want, ng) ** 2).mean()))


@pytest.mark.parametrize("scheme", [0, 1])
def test_observed_order_on_a_tanh_grid_oracle(analytic, scheme):
    """totani_lr + visc_lr, and config 3's hybrid(totani_lr, fweno_t, ducros_t) + visc_lr, on a tanh-stretched grid: the
    viscous / sensor gradient transform (this library's completion, parity unpinned) is at work on every face, and the total
    right-hand side must converge at second order to the analytic one"""
    sizes = (16, 32, 64)
    errs = [_error_tanh(analytic, scheme, n, 2, False) for n in sizes]
    rates = [np.log(errs[i] / errs[i + 1]) / np.log(2.0) for i in range(2)]
    assert rates[-1] > 1.7 and errs[-1] < 0.25 * errs[0], (errs, rates)


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", [0, 1])
def test_observed_order_on_a_tanh_grid_gpu(analytic, scheme):
    """the same order check through the CUDA path (C ABI): the kernels themselves against the analytic right-hand side, not
    against the oracle — the independent check of the unpinned curvilinear viscous and Ducros terms"""
    sizes = (16, 32, 64)
    errs = [_error_tanh(analytic, scheme, n, 2, True) for n in sizes]
    rates = [np.log(errs[i] / errs[i + 1]) / np.log(2.0) for i in range(2)]
    assert rates[-1] > 1.7 and errs[-1] < 0.25 * errs[0], (errs, rates)
    # and the two paths agree on the finest grid to the parity tolerance
    assert abs(errs[-1] - _error_tanh(analytic, scheme, sizes[-1], 2, False)) < 1e-9 * errs[-1] + 1e-12
