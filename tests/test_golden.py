"""The oracle (oracle/spade_oracle.c) and the CUDA path against the committed golden vectors, which were produced
by the unmodified reference (tests/golden/make_golden.py). The CPU half runs anywhere; the GPU half is the same
comparison through the C ABI."""
import os

import numpy as np
import pytest

from util import GAMMA, RGAS, oracle_cfg, product_flux, product_setup, rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))
NB, N, NG = (2, 1, 2), (8, 4, 4), 2


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "hotpath_small.npz"))


@pytest.fixture(scope="module")
def tabs():
    return np.load(os.path.join(HERE, "golden", "exchange_tables.npz"))


# ---------------------------------------------------------------- oracle vs golden (CPU)
@pytest.mark.parametrize("scheme", range(9))
def test_oracle_flux_div(gold, scheme):
    from oracle import port
    q = gold[f"fdiv_q_{scheme}"]
    got = port.flux_div(oracle_cfg(NB, N, NG, scheme=scheme), q.ravel()).reshape(q.shape)
    assert rel_l2(got, gold[f"fdiv_rhs_{scheme}"]) < 1e-14


@pytest.mark.parametrize("tag,periodic", [("ppp", (1, 1, 1)), ("pwp", (1, 0, 1)), ("www", (0, 0, 0))])
def test_oracle_exchange(gold, tag, periodic):
    from oracle import port
    q = gold[f"exch_in_{tag}"]
    got = port.exchange(oracle_cfg(NB, N, NG, periodic=periodic), q.ravel()).reshape(q.shape)
    assert np.array_equal(got, gold[f"exch_out_{tag}"])


@pytest.mark.parametrize("integ,name", [(0, "rk4"), (1, "ssprk3opt")])
def test_oracle_trajectory(gold, integ, name):
    from oracle import port
    q0 = gold[f"adv_q0_{name}"]
    dt, umax = gold[f"adv_dt_{name}"]
    cfg = oracle_cfg(NB, N, NG, scheme=0, integrator=integ)
    assert port.reduce_umax(cfg, q0.ravel()) == umax
    got = port.advance(cfg, q0.ravel(), float(dt), 3).reshape(q0.shape)
    assert rel_l2(got, gold[f"adv_q3_{name}"]) < 1e-14


@pytest.mark.parametrize("nranks", [1, 2, 4, 8])
@pytest.mark.parametrize("tag,periodic", [("ppp", (1, 1, 1)), ("pwp", (1, 0, 1))])
def test_oracle_exchange_tables(tabs, nranks, tag, periodic):
    from oracle import port
    cfg = oracle_cfg((2, 2, 2), (16, 16, 16), 2, periodic=periodic, nranks=nranks)
    for rank in range(nranks):
        s, r, o = port.exchange_tables(cfg, rank)
        assert np.array_equal(s, tabs[f"send_{nranks}_{tag}_{rank}"])
        assert np.array_equal(r, tabs[f"recv_{nranks}_{tag}_{rank}"])
        assert np.array_equal(o, tabs[f"offs_{nranks}_{tag}_{rank}"])


def test_mem_map_offsets_exhaustive():
    """Appendix B: off(v,i,j,k,lb) = v + 5*((i+g) + (n0+2g)*((j+g) + (n1+2g)*((k+g) + (n2+2g)*lb))) (mem_map.h:484-496)."""
    from oracle import port
    n, g, nlb = (5, 3, 4), 2, 3
    cfg = oracle_cfg((nlb, 1, 1), n, g)
    seen = set()
    for lb in range(nlb):
        for k in range(-g, n[2] + g):
            for j in range(-g, n[1] + g):
                for i in range(-g, n[0] + g):
                    for v in range(5):
                        off = port.offset(cfg, v, i, j, k, lb)
                        want = v + 5 * ((i + g) + (n[0] + 2 * g) * ((j + g) + (n[1] + 2 * g) * ((k + g) + (n[2] + 2 * g) * lb)))
                        assert off == want
                        seen.add(off)
    assert len(seen) == port.array_size(cfg) and max(seen) == len(seen) - 1      # a bijection onto [0, size)


# ---------------------------------------------------------------- the product's host-side plan builder (CPU, no GPU needed)
@pytest.mark.parametrize("nranks", [1, 2, 4, 8])
@pytest.mark.parametrize("tag,periodic", [("ppp", (1, 1, 1)), ("pwp", (1, 0, 1))])
def test_product_exchange_plan_bit_exact(tabs, nranks, tag, periodic):
    """spb_exchange_create is host-only: its transaction lists must be bit-identical to exchange_config_t's."""
    import ctypes as C
    from spade_b200._lib import lib, check, int3
    for rank in range(nranks):
        h = C.c_void_p()
        check(lib().spb_exchange_create(C.byref(h), int3((2, 2, 2)), int3((16, 16, 16)), int3((2, 2, 2)), int3(periodic), rank, nranks))
        ns, nr = lib().spb_exchange_num_send(h), lib().spb_exchange_num_recv(h)
        s = np.zeros((ns, 16), dtype=np.int64)
        r = np.zeros((nr, 16), dtype=np.int64)
        o = np.zeros((nranks, 6), dtype=np.int64)
        i64 = C.POINTER(C.c_int64)
        check(lib().spb_exchange_tables(h, s.ctypes.data_as(i64), r.ctypes.data_as(i64), o.ctypes.data_as(i64)))
        lib().spb_exchange_destroy(h)
        assert np.array_equal(s, tabs[f"send_{nranks}_{tag}_{rank}"])
        assert np.array_equal(r, tabs[f"recv_{nranks}_{tag}_{rank}"])
        assert np.array_equal(o, tabs[f"offs_{nranks}_{tag}_{rank}"])


# ---------------------------------------------------------------- CUDA path vs golden (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("scheme", range(9))
def test_cuda_flux_div(gold, scheme):
    sp, blocks, grid = product_setup(NB, N, NG)
    q = gold[f"fdiv_q_{scheme}"]
    qa = sp.grid_array.from_host(grid, q)
    ra = sp.grid_array(grid, 0.0)
    sp.flux_div(qa, ra, product_flux(scheme), sp.overwrite)
    assert rel_l2(ra.to_host(), gold[f"fdiv_rhs_{scheme}"]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("tag,periodic", [("ppp", (1, 1, 1)), ("pwp", (1, 0, 1)), ("www", (0, 0, 0))])
def test_cuda_exchange(gold, tag, periodic):
    sp, blocks, grid = product_setup(NB, N, NG)
    qa = sp.grid_array.from_host(grid, gold[f"exch_in_{tag}"])
    sp.make_exchange(qa, periodic).exchange(qa)
    assert np.array_equal(qa.to_host(), gold[f"exch_out_{tag}"])


@pytest.mark.gpu
@pytest.mark.parametrize("integ,name", [(0, "rk4"), (1, "ssprk3opt")])
def test_cuda_trajectory(gold, integ, name):
    sp, blocks, grid = product_setup(NB, N, NG)
    gas = sp.ideal_gas_t(GAMMA, RGAS)
    qa = sp.grid_array.from_host(grid, gold[f"adv_q0_{name}"])
    ra = sp.grid_array(grid, 0.0)
    ex = sp.make_exchange(qa, (1, 1, 1))
    flux = sp.flux_desc(product_flux(0))
    alg = sp.rk4_t if integ == 0 else sp.ssprk3_opt
    dt, umax = gold[f"adv_dt_{name}"]
    assert sp.transform_reduce(qa, sp.FN_WAVESPEED, sp.RED_MAX, gas) == pytest.approx(umax, rel=1e-15)
    ti = sp.integrator_t(sp.time_axis_t(0.0, float(dt)), alg, sp.integrator_data_t(qa, ra, alg),
                         lambda r, qq, t: sp.flux_div(qq, r, flux, sp.overwrite), lambda qq, t: ex.exchange(qq),
                         sp.state_transform_t(gas))
    for _ in range(3):
        ti.advance()
    assert rel_l2(ti.solution().to_host(), gold[f"adv_q3_{name}"]) < 1e-12


# ---------------------------------------------------------------- channel rows: boundary_fill, source_term
CH_NB, CH_N = (2, 2, 1), (8, 4, 4)


@pytest.fixture(scope="module")
def chan():
    return np.load(os.path.join(HERE, "golden", "channel_small.npz"))


CHANNEL_WALL_T = 310.0
CHANNEL_FORCE = (40.0, -0.5, 0.25)
CHANNEL_BCS = {
    "noslip_isothermal_y": dict(mask=(0, 0, 1, 1, 0, 0), a=(1, -1, -1, -1, -1), b=(0, 2 * CHANNEL_WALL_T, 0, 0, 0)),
    "adiabatic_all": dict(mask=(1, 1, 1, 1, 1, 1), a=(1, 1, -1, -1, -1)),
    "symmetry_x_z": dict(mask=(1, 1, 0, 0, 1, 1), a=(1, 1, 1, 1, 1), a_normal=-1.0),
    "extrap2_all": dict(mask=(1, 1, 1, 1, 1, 1), kind=1, order=2),
}


@pytest.mark.parametrize("name", sorted(CHANNEL_BCS))
def test_oracle_boundary_fill(chan, name):
    from oracle import port, ref
    q = chan["q"]
    got = port.boundary_fill(oracle_cfg(CH_NB, CH_N, NG, periodic=(0, 0, 0)), ref.make_bc(**CHANNEL_BCS[name]), q.ravel())
    assert np.array_equal(got.reshape(q.shape), chan[f"fill_{name}"])


def test_oracle_source_term_and_channel_trajectory(chan):
    from oracle import port, ref
    q = chan["q"]
    cfgw = oracle_cfg(CH_NB, CH_N, NG, periodic=(0, 0, 0))
    got = port.source_term(cfgw, ref.make_bc(mask=(0,) * 6, force=CHANNEL_FORCE), q.ravel(), chan["src_rhs0"].ravel())
    assert np.array_equal(got.reshape(q.shape), chan["src_rhs1"])
    cfgc = oracle_cfg(CH_NB, CH_N, NG, scheme=0, integrator=0, periodic=(1, 0, 1))
    bc = ref.make_bc(force=CHANNEL_FORCE, **CHANNEL_BCS["noslip_isothermal_y"])
    dt, umax = chan["adv_dt"]
    assert port.reduce_umax(cfgc, chan["adv_q0"].ravel()) == umax
    got = port.advance_channel(cfgc, bc, chan["adv_q0"].ravel(), float(dt), 3).reshape(q.shape)
    assert rel_l2(got, chan["adv_q3"]) < 1e-14


@pytest.mark.gpu
def test_gpu_channel_golden(chan):
    """boundary_fill / source_term / wall-bounded forced rk4 trajectory of the CUDA path against the reference's own output"""
    sp, blocks, grid = product_setup(CH_NB, CH_N, NG)
    q = chan["q"]
    kern = {"noslip_isothermal_y": (sp.boundary.ymin | sp.boundary.ymax, sp.noslip_isothermal_wall(CHANNEL_WALL_T)),
            "adiabatic_all": (sp.identifier_t(1, 1, 1, 1, 1, 1), sp.noslip_adiabatic_wall()),
            "symmetry_x_z": (sp.boundary.xmin | sp.boundary.xmax | sp.boundary.zmin | sp.boundary.zmax, sp.symmetry_plane()),
            "extrap2_all": (sp.identifier_t(1, 1, 1, 1, 1, 1), sp.boundary.extrapolate(2))}
    for name, (which, k) in kern.items():
        qa = sp.grid_array.from_host(grid, q)
        sp.boundary_fill(qa, which, k)
        assert np.array_equal(qa.to_host(), chan[f"fill_{name}"]), name
    qa, ra = sp.grid_array.from_host(grid, q), sp.grid_array.from_host(grid, chan["src_rhs0"])
    sp.source_term(qa, ra, sp.body_force_t(*CHANNEL_FORCE))
    assert np.array_equal(ra.to_host(), chan["src_rhs1"])

    gas = sp.ideal_gas_t(GAMMA, RGAS)
    flux = sp.flux_desc(product_flux(0))
    qa, ra = sp.grid_array.from_host(grid, chan["adv_q0"]), sp.grid_array(grid, 0.0)
    ex = sp.make_exchange(qa, (1, 0, 1))
    wall, force = sp.noslip_isothermal_wall(CHANNEL_WALL_T), sp.body_force_t(*CHANNEL_FORCE)

    def calc_rhs(r, qq, t):
        sp.flux_div(qq, r, flux, sp.overwrite)
        sp.source_term(qq, r, force)

    def boundary_cond(qq, t):
        ex.exchange(qq)
        sp.boundary_fill(qq, sp.boundary.ymin | sp.boundary.ymax, wall)

    ti = sp.integrator_t(sp.time_axis_t(0.0, float(chan["adv_dt"][0])), sp.rk4_t, sp.integrator_data_t(qa, ra, sp.rk4_t),
                         calc_rhs, boundary_cond, sp.state_transform_t(gas))
    for _ in range(3):
        ti.advance()
    assert rel_l2(ti.solution().to_host(), chan["adv_q3"]) < 1e-12
