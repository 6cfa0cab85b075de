"""Shared helpers of the test-suite: seeded synthetic initial conditions in the reference's memory
order and error norms. Test infrastructure only."""
import numpy as np

P0, T0, U0, RGAS, GAMMA = 101325.0, 300.0, 34.7, 287.15, 1.4


def block_index(lb, nb):
    return lb % nb[0], (lb // nb[0]) % nb[1], lb // (nb[0] * nb[1])


def make_state(nb, n, ng, seed=0, perturb=1e-2, bounds=None, jump=False):
    """Taylor-Green-like field (development/cuda-tgv/main.cc:107-116) + seeded relative perturbation,
    all cells incl. ghosts filled analytically; shape [nlb, nk', nj', ni', 5] (global blocks)."""
    if bounds is None:
        bounds = [0.0, 2 * np.pi] * 3
    nlb = nb[0] * nb[1] * nb[2]
    q = np.zeros((nlb, n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng, 5))
    rho0 = P0 / (RGAS * T0)
    for lb in range(nlb):
        b = block_index(lb, nb)
        ax = []
        for d in range(3):
            bsize = (bounds[2 * d + 1] - bounds[2 * d]) / nb[d]
            lo = bounds[2 * d] + b[d] * bsize
            dx = bsize / n[d]
            ax.append(lo + (np.arange(-ng, n[d] + ng) + 0.5) * dx)
        Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
        q[lb, ..., 0] = P0 + rho0 * U0 * U0 / 16 * (np.cos(2 * X) + np.cos(2 * Y)) * (np.cos(2 * Z) + 2)
        q[lb, ..., 1] = T0 * (1 + 0.02 * np.sin(X + 2 * Y - Z))
        q[lb, ..., 2] = U0 * np.sin(X) * np.cos(Y) * np.cos(Z)
        q[lb, ..., 3] = -U0 * np.cos(X) * np.sin(Y) * np.cos(Z)
        q[lb, ..., 4] = 0.3 * U0 * np.sin(Z) * np.cos(X + Y)
        if jump:   # planted pressure jump to wake the shock sensor
            q[lb, ..., 0] *= np.where(np.sin(X) > 0.3, 1.4, 1.0)
    if perturb:
        rng = np.random.default_rng(12345 + seed)
        q *= 1 + perturb * rng.uniform(-1, 1, q.shape)
    return q


def amr_state(boxes, n, ng, seed):
    """smooth positive state + seeded perturbation on blocks given by their boxes, all cells incl. ghosts analytic"""
    q = np.zeros((boxes.shape[0], n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng, 5))
    for lb in range(boxes.shape[0]):
        ax = [boxes[lb, 2 * d] + (np.arange(-ng, n[d] + ng) + 0.5) * (boxes[lb, 2 * d + 1] - boxes[lb, 2 * d]) / n[d] for d in range(3)]
        Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
        q[lb, ..., 0] = 101325.0 * (1 + 0.05 * np.cos(X) * np.cos(Y))
        q[lb, ..., 1] = 300.0 * (1 + 0.02 * np.sin(X + 2 * Y - Z))
        q[lb, ..., 2] = U0 * np.sin(X) * np.cos(Y) * np.cos(Z)
        q[lb, ..., 3] = -U0 * np.cos(X) * np.sin(Y) * np.cos(Z)
        q[lb, ..., 4] = 10.0 * np.sin(Z) * np.cos(X + Y)
    return q * (1 + 1e-2 * np.random.default_rng(seed).uniform(-1, 1, q.shape))


def zero_ghosts(q, ng):
    out = q.copy()
    mask = np.zeros(q.shape[1:4], dtype=bool)
    mask[ng:-ng, ng:-ng, ng:-ng] = True
    out[:, ~mask, :] = 0.0
    return out


def interior(q, ng):
    return q[:, ng:-ng, ng:-ng, ng:-ng, :]


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


# scheme id of oracle (spo_cfg.scheme) -> constructor of the product-side functor
SGS = (0.55, 0.4, 0.9)      # wale_t(gas, cw, delta, prt) of schemes 11, 12


def product_flux(scheme, gamma=GAMMA, R=RGAS, mu=1e-2, prandtl=0.72, eps=1e-2):
    import spade_b200.api as sp
    gas = sp.ideal_gas_t(gamma, R)
    vl = sp.constant_viscosity_t(mu, prandtl)
    t, w, v, du = sp.totani_lr(gas), sp.fweno_t(gas), sp.visc_lr(vl, gas), sp.ducros_t(eps)
    ck4 = sp.cent_keep(4, gas)
    return {0: lambda: sp.compose(t, v),
            1: lambda: sp.compose(sp.hybrid_scheme_t(t, w, du, sp.full_flux), v),
            2: lambda: sp.compose(ck4, v),
            3: lambda: t,
            4: lambda: v,
            5: lambda: w,
            6: lambda: sp.compose(sp.hybrid_scheme_t(ck4, w, du, sp.full_flux), v),
            7: lambda: ck4,
            8: lambda: sp.compose(sp.hybrid_scheme_t(t, w, du, sp.diss_flux), v),
            11: lambda: sp.compose(t, sp.visc_lr(sp.sgs_visc_t(vl, sp.wale_t(gas, *SGS)), gas)),
            12: lambda: sp.compose(sp.hybrid_scheme_t(t, w, du, sp.full_flux), sp.visc_lr(sp.sgs_visc_t(vl, sp.wale_t(gas, *SGS)), gas)),
            13: lambda: sp.compose(sp.cent_keep(6, gas), v),
            14: lambda: sp.compose(sp.cent_keep(8, gas), v),
            15: lambda: sp.cent_keep(6, gas),
            16: lambda: sp.cent_keep(8, gas),
            17: lambda: sp.fweno_t(gas, sp.disable_smooth),
            18: lambda: sp.weno_t(sp.rusanov_t(gas), sp.disable_smooth),
            19: lambda: sp.compose(sp.hybrid_scheme_t(t, sp.fweno_t(gas, sp.disable_smooth), du, sp.full_flux), v),
            9: lambda: sp.weno_t(sp.rusanov_t(gas)),
            10: lambda: sp.compose(sp.hybrid_scheme_t(t, sp.weno_t(sp.rusanov_t(gas)), du, sp.full_flux), v)}[scheme]()


def oracle_cfg(nb, n, ng=2, scheme=0, mu=1e-2, prandtl=0.72, eps=1e-2, periodic=(1, 1, 1), nranks=1, integrator=0,
               bounds=None):
    from oracle import ref
    return ref.make_cfg(nb, n, ng, bounds=bounds, periodic=periodic, scheme=scheme, gamma=GAMMA, R=RGAS, mu=mu,
                        prandtl=prandtl, sensor_eps=eps, nranks=nranks, integrator=integrator, sgs=SGS)


def product_setup(nb, n, ng=2, bounds=None, rank=0, size=1):
    """grid + blocks on the product side for the same lattice as oracle_cfg."""
    import spade_b200.api as sp
    if bounds is None:
        bounds = [0.0, 2 * np.pi] * 3
    blocks = sp.cartesian_blocks_t(nb, bounds)
    grid = sp.cartesian_grid_t(n, blocks, sp.identity(), sp.pool_t(rank, size))
    return sp, blocks, grid
