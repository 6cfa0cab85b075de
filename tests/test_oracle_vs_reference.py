"""Pins the oracle: the plain-C restatement (oracle/spade_oracle.c) against the UNMODIFIED reference
(oracle/_ref/libspade_ref.so, compiled from /root/reference/src by oracle/Makefile).

CPU only. Where the prebuilt reference library is absent the tests skip; the committed golden vectors
(tests/golden, generated from the reference library by tests/golden/make_golden.py) pin the oracle
instead (tests/test_golden.py).

Reference paths: flux_div_basic.h:17-77, make_exchange.h:111-410, exchange_config.h:286-419,
advance.h:57-102,236-402, transform_reduce.h:53-191, partition.h:27-84, fluid_state.h:103-135."""
import numpy as np
import pytest

from util import GAMMA, RGAS, make_state, oracle_cfg, rel_l2, zero_ghosts


@pytest.mark.parametrize("scheme", range(20))          # 17-19: fweno_t / weno_t / hybrid with disable_smooth
def test_flux_div_all_schemes(ref_lib, scheme):
    from oracle import port
    nb, n, ng = (2, 1, 2), (8, 6, 4), {13: 3, 14: 4, 15: 3, 16: 4}.get(scheme, 2)
    q = make_state(nb, n, ng, seed=scheme, jump=scheme in (1, 6, 8, 10, 12, 19))
    cfg = oracle_cfg(nb, n, ng, scheme=scheme)
    want = ref_lib.flux_div(cfg, q.ravel())
    got = port.flux_div(cfg, q.ravel())
    # same operation order as the reference: round-off level (bit-exact for most schemes)
    assert rel_l2(got, want) < 1e-14


def test_flux_div_increment_and_anisotropic(ref_lib):
    from oracle import port
    nb, n, ng = (1, 2, 1), (6, 4, 10), 2
    bounds = [0.0, 3.0, -1.0, 1.0, 0.0, 7.0]
    q = make_state(nb, n, ng, seed=5, bounds=bounds)
    rhs0 = np.random.default_rng(3).normal(size=q.size) * 1e3
    cfg = oracle_cfg(nb, n, ng, scheme=0, bounds=bounds)
    want = ref_lib.flux_div(cfg, q.ravel(), rhs=rhs0, increment=True)
    got = port.flux_div(cfg, q.ravel(), rhs=rhs0, increment=True)
    assert rel_l2(got, want) < 1e-14


@pytest.mark.parametrize("periodic", [(1, 1, 1), (1, 0, 1), (0, 0, 0)])
@pytest.mark.parametrize("nb,n", [((2, 2, 2), (8, 8, 8)), ((1, 3, 2), (8, 4, 6)), ((1, 1, 1), (8, 8, 8)), ((3, 1, 4), (4, 6, 4))])
def test_exchange_bit_exact(ref_lib, periodic, nb, n):
    from oracle import port
    ng = 2
    q = zero_ghosts(make_state(nb, n, ng, seed=2), ng)
    cfg = oracle_cfg(nb, n, ng, periodic=periodic)
    assert np.array_equal(port.exchange(cfg, q.ravel()), ref_lib.exchange(cfg, q.ravel()))


@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("periodic", [(1, 1, 1), (1, 0, 1)])
def test_exchange_tables_bit_exact(ref_lib, nranks, periodic):
    """send/recv transaction lists (order, tags, boxes, ranks, block ids) and per-peer offsets, every rank."""
    from oracle import port
    nb, n, ng = (2, 2, 2), (16, 16, 16), 2
    cfg = oracle_cfg(nb, n, ng, periodic=periodic, nranks=nranks)
    for rank in range(nranks):
        s0, r0, o0 = ref_lib.exchange_tables(cfg, rank)
        s1, r1, o1 = port.exchange_tables(cfg, rank)
        assert np.array_equal(s0, s1) and np.array_equal(r0, r1) and np.array_equal(o0, o1)


def test_exchange_tables_sanity_anchors(ref_lib):
    """SURVEY Appendix B anchors observed on the reference: 2x2x2 blocks of 16^3, g=2."""
    nb, n, ng = (2, 2, 2), (16, 16, 16), 2
    s, r, o = ref_lib.exchange_tables(oracle_cfg(nb, n, ng, periodic=(1, 1, 1), nranks=1), 0)
    assert len(s) == 208 and len(r) == 208 and o[0, 0] == 31232
    assert list(s[:64, 0]) == [16] * 64 and list(s[64:96, 0]) == [12] * 32 and list(s[-16:, 0]) == [1] * 16
    cfg2 = oracle_cfg(nb, n, ng, periodic=(1, 0, 1), nranks=2)
    s, r, o = ref_lib.exchange_tables(cfg2, 0)
    assert len(s) == 68 and len(r) == 68
    assert o[0, 0] == 6656 and o[1, 0] == 5760 and o[0, 3] == 20 and o[1, 3] == 48


@pytest.mark.parametrize("integ", [0, 1, 2, 3, 4, 5])       # rk4, ssprk3_opt, ssprk3, rk2, ssprk34, rk38r (explicit.h:48-116)
def test_rk_trajectory(ref_lib, integ):
    from oracle import port
    nb, n, ng = (2, 2, 1), (8, 4, 4), 2
    q0 = make_state(nb, n, ng, seed=13)
    cfg = oracle_cfg(nb, n, ng, scheme=0, integrator=integ)
    q0 = port.exchange(cfg, q0.ravel())
    dt = 0.2 * (2 * np.pi / 16) / port.reduce_umax(cfg, q0)
    want, _ = ref_lib.advance(cfg, q0, dt, 3)
    got = port.advance(cfg, q0, dt, 3)
    assert rel_l2(got, want) < 1e-14


def test_rk_trajectory_hybrid_multirank_reference(ref_lib):
    """the reference run on 4 thread-ranks gives the same trajectory as the single-rank oracle."""
    from oracle import port
    nb, n, ng = (2, 2, 2), (4, 4, 4), 2
    q0 = make_state(nb, n, ng, seed=17, jump=True)
    cfg1 = oracle_cfg(nb, n, ng, scheme=1, integrator=0)
    cfg4 = oracle_cfg(nb, n, ng, scheme=1, integrator=0, nranks=4)
    q0 = port.exchange(cfg1, q0.ravel())
    dt = 0.2 * (2 * np.pi / 8) / port.reduce_umax(cfg1, q0)
    want, _ = ref_lib.advance(cfg4, q0, dt, 2)
    got = port.advance(cfg1, q0, dt, 2)
    assert rel_l2(got, want) < 1e-14


def test_reduce_and_state_conversion(ref_lib):
    from oracle import port
    nb, n, ng = (2, 1, 2), (8, 4, 6), 2
    q = make_state(nb, n, ng, seed=9)
    cfg = oracle_cfg(nb, n, ng)
    assert port.reduce_umax(cfg, q.ravel()) == ref_lib.reduce_umax(cfg, q.ravel())
    rng = np.random.default_rng(0)
    for _ in range(20):
        p = np.array([1e5 * rng.uniform(0.5, 2), 300 * rng.uniform(0.5, 2), *rng.normal(size=3) * 50])
        w = ref_lib.prim2cons(GAMMA, RGAS, p)
        assert np.array_equal(w, port.prim2cons(GAMMA, RGAS, p))
        assert np.array_equal(ref_lib.cons2prim(GAMMA, RGAS, w), port.cons2prim(GAMMA, RGAS, w))


@pytest.mark.parametrize("nglob,nranks", [(8, 1), (8, 3), (27, 4), (64, 8), (5, 8), (4096, 8)])
def test_partition_matches_reference_tables(ref_lib, nglob, nranks):
    """partition::block_partition_t (partition.h:40-71) seen through the reference's exchange tables:
    every transaction's (rank_send, glob_src) and (rank_recv, glob_dst) agree with spo_partition."""
    from oracle import port
    g2r, g2l = port.partition(nglob, nranks)
    per, extra = divmod(nglob, nranks)
    # contiguous runs, first `extra` ranks get one more
    counts = np.bincount(g2r, minlength=nranks)
    assert list(counts) == [per + (1 if r < extra else 0) for r in range(nranks)]
    assert np.all(np.diff(g2r) >= 0)
    if nglob in (8, 64):
        nb = (2, 2, 2) if nglob == 8 else (4, 4, 4)
        cfg = oracle_cfg(nb, (4, 4, 4), 2, nranks=nranks)
        for rank in range(nranks):
            s, r, _ = ref_lib.exchange_tables(cfg, rank)
            assert np.array_equal(g2r[s[:, 3]], s[:, 1]) and np.array_equal(g2r[s[:, 4]], s[:, 2])
            assert np.array_equal(g2l[s[:, 3]], s[:, 8])          # local id of the source block on the sender
            assert np.array_equal(g2l[r[:, 4]], r[:, 15])


# ---- "next" rows: boundary_fill and source_term (channel runs) --------------------------------------------------
WALL_T = 310.0
BCS = {
    "noslip_isothermal_y": dict(mask=(0, 0, 1, 1, 0, 0), a=(1, -1, -1, -1, -1), b=(0, 2 * WALL_T, 0, 0, 0)),
    "adiabatic_all": dict(mask=(1, 1, 1, 1, 1, 1), a=(1, 1, -1, -1, -1)),
    "symmetry_x_z": dict(mask=(1, 1, 0, 0, 1, 1), a=(1, 1, 1, 1, 1), a_normal=-1.0),
    "extrap1_ymax": dict(mask=(0, 0, 0, 1, 0, 0), kind=1, order=1),
    "extrap2_all": dict(mask=(1, 1, 1, 1, 1, 1), kind=1, order=2),
}


@pytest.mark.parametrize("name", sorted(BCS))
@pytest.mark.parametrize("nranks", [1, 3])
def test_boundary_fill_bit_exact(ref_lib, name, nranks):
    """oracle restatement of algs::boundary_fill (src/grid/boundary_fill.h:32-133) against the reference itself"""
    from oracle import port, ref
    nb, n, ng = (2, 3, 2), (8, 4, 8), 2
    q = make_state(nb, n, ng, seed=41)
    bc = ref.make_bc(**BCS[name])
    cfg = oracle_cfg(nb, n, ng, periodic=(0, 0, 0), nranks=nranks)
    want = ref_lib.boundary_fill(cfg, bc, q.ravel())
    got = port.boundary_fill(cfg, bc, q.ravel())
    assert np.array_equal(got, want)
    assert not np.array_equal(got, q.ravel())


def test_source_term_bit_exact(ref_lib):
    from oracle import port, ref
    nb, n, ng = (2, 1, 2), (8, 8, 4), 2
    q = make_state(nb, n, ng, seed=42)
    rhs0 = np.random.default_rng(3).standard_normal(q.size)
    bc = ref.make_bc(mask=(0,) * 6, force=(3.5, -0.25, 0.125))
    cfg = oracle_cfg(nb, n, ng)
    assert np.array_equal(port.source_term(cfg, bc, q.ravel(), rhs0), ref_lib.source_term(cfg, bc, q.ravel(), rhs0))


@pytest.mark.parametrize("integ", [0, 2])
def test_channel_trajectory(ref_lib, integ):
    """x/z periodic, isothermal no-slip walls in y, body force: exchange + boundary_fill + flux_div + source_term + RK"""
    from oracle import port, ref
    nb, n, ng = (2, 2, 2), (8, 8, 8), 2
    periodic = (1, 0, 1)
    q0 = make_state(nb, n, ng, seed=43)
    cfg = oracle_cfg(nb, n, ng, scheme=0, integrator=integ, periodic=periodic, nranks=2)
    bc = ref.make_bc(force=(40.0, 0.0, 0.0), **BCS["noslip_isothermal_y"])
    q0 = port.boundary_fill(cfg, bc, port.exchange(cfg, q0.ravel()))
    dt = 0.2 * (2 * np.pi / 16) / port.reduce_umax(cfg, q0)
    want, _ = ref_lib.advance_channel(cfg, bc, q0, dt, 3)
    got = port.advance_channel(cfg, bc, q0, dt, 3)
    assert rel_l2(got, want) < 1e-14


@pytest.mark.parametrize("integ,hs", [(0, False), (2, False), (3, False), (2, True), (3, True)])
def test_generic_integrate_advance_bit_exact(ref_lib, integ, hs):
    """integrator_t without a state transform (identity_transform): the generic integrate_advance of advance.h:109-230,
    including the rounding of its scale / add / unscale passes and the high-storage tables rk2hs_t / ssprk3hs_t."""
    from oracle import port
    nb, n, ng = (2, 1, 2), (8, 4, 4), 2
    cfg = oracle_cfg(nb, n, ng, scheme=0, integrator=integ)
    q0 = port.exchange(cfg, make_state(nb, n, ng, seed=3).ravel())
    want = ref_lib.advance_generic(cfg, q0, 1e-5, 2, hs)
    assert np.array_equal(port.advance_generic(cfg, q0, 1e-5, 2, hs), want)
    assert rel_l2(want, q0) > 1e-5


@pytest.mark.parametrize("ng", [1, 3, 4])
@pytest.mark.parametrize("periodic", [(1, 1, 1), (1, 0, 1)])
def test_exchange_with_other_exchange_depths_bit_exact(ref_lib, ng, periodic):
    """1, 3 and 4 exchange cells (visc_lr alone, cent_keep<6>, cent_keep<8>): ghost fill and tables against the reference."""
    from oracle import port
    nb, n = (2, 2, 1), (8, 4, 4)
    q = zero_ghosts(make_state(nb, n, ng, seed=4), ng)
    cfg = oracle_cfg(nb, n, ng, periodic=periodic, nranks=2)
    assert np.array_equal(port.exchange(cfg, q.ravel()), ref_lib.exchange(cfg, q.ravel()))
    for rank in range(2):
        s0, r0, o0 = ref_lib.exchange_tables(cfg, rank)
        s1, r1, o1 = port.exchange_tables(cfg, rank)
        assert np.array_equal(s0, s1) and np.array_equal(r0, r1) and np.array_equal(o0, o1)
