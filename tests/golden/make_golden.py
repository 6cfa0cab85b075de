#!/usr/bin/env python
"""Generates the committed golden vectors of tests/golden/*.npz from the UNMODIFIED reference
(oracle/_ref/libspade_ref.so = /root/reference/src behind oracle/ref_driver.cc). Run in the dev container:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; it checks the CUDA path and the oracle against these files.
Reference paths exercised: flux_div_basic.h:17-77, make_exchange.h:111-410, exchange_config.h:286-419,
advance.h:57-102,236-402, transform_reduce.h:53-191."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from util import amr_state, make_state, oracle_cfg, zero_ghosts  # noqa: E402
from oracle import ref  # noqa: E402

CHANNEL_WALL_T = 310.0
CHANNEL_FORCE = (40.0, -0.5, 0.25)
CHANNEL_BCS = {
    "noslip_isothermal_y": dict(mask=(0, 0, 1, 1, 0, 0), a=(1, -1, -1, -1, -1), b=(0, 2 * CHANNEL_WALL_T, 0, 0, 0)),
    "adiabatic_all": dict(mask=(1, 1, 1, 1, 1, 1), a=(1, 1, -1, -1, -1)),
    "symmetry_x_z": dict(mask=(1, 1, 0, 0, 1, 1), a=(1, 1, 1, 1, 1), a_normal=-1.0),
    "extrap2_all": dict(mask=(1, 1, 1, 1, 1, 1), kind=1, order=2),
}
NB, N, NG = (2, 1, 2), (8, 4, 4), 2   # multiples of 4: the reference's transform_inplace tiles are 4^3 (transform_inplace.h:34-96)


def main():
    assert ref.available(), "build oracle/_ref first (make -C oracle ref)"
    out = {}
    # (i) rhs of one flux_div for every scheme combination
    for scheme in range(9):
        q = make_state(NB, N, NG, seed=100 + scheme, jump=scheme in (1, 6, 8))
        cfg = oracle_cfg(NB, N, NG, scheme=scheme)
        out[f"fdiv_q_{scheme}"] = q
        out[f"fdiv_rhs_{scheme}"] = ref.flux_div(cfg, q.ravel()).reshape(q.shape)
    # (ii) exchange (ghost fill) for three periodicities
    for tag, periodic in (("ppp", (1, 1, 1)), ("pwp", (1, 0, 1)), ("www", (0, 0, 0))):
        q = zero_ghosts(make_state(NB, N, NG, seed=7), NG)
        cfg = oracle_cfg(NB, N, NG, periodic=periodic)
        out[f"exch_in_{tag}"] = q
        out[f"exch_out_{tag}"] = ref.exchange(cfg, q.ravel()).reshape(q.shape)
    # (iii) short RK trajectories (rk4 and the 2-register ssprk3) and the CFL reduction
    for integ, name in ((0, "rk4"), (1, "ssprk3opt")):
        cfg = oracle_cfg(NB, N, NG, scheme=0, integrator=integ)
        q0 = ref.exchange(cfg, make_state(NB, N, NG, seed=13).ravel())
        umax = ref.reduce_umax(cfg, q0)
        dt = 0.2 * (2 * np.pi / 16) / umax
        q1, _ = ref.advance(cfg, q0, dt, 3)
        out[f"adv_q0_{name}"] = q0.reshape((-1,) + (N[2] + 2 * NG, N[1] + 2 * NG, N[0] + 2 * NG, 5))
        out[f"adv_q3_{name}"] = q1.reshape(out[f"adv_q0_{name}"].shape)
        out[f"adv_dt_{name}"] = np.array([dt, umax])
    np.savez_compressed(os.path.join(HERE, "hotpath_small.npz"), **out)
    # (iv) exchange tables (bit-exact integer maps) for 1/2/4/8 ranks, periodic and wall-bounded
    tabs = {}
    for nranks in (1, 2, 4, 8):
        for tag, periodic in (("ppp", (1, 1, 1)), ("pwp", (1, 0, 1))):
            cfg = oracle_cfg((2, 2, 2), (16, 16, 16), 2, periodic=periodic, nranks=nranks)
            for rank in range(nranks):
                s, r, o = ref.exchange_tables(cfg, rank)
                tabs[f"send_{nranks}_{tag}_{rank}"] = s
                tabs[f"recv_{nranks}_{tag}_{rank}"] = r
                tabs[f"offs_{nranks}_{tag}_{rank}"] = o
    np.savez_compressed(os.path.join(HERE, "exchange_tables.npz"), **tabs)
    # (v) channel: boundary_fill kernels, source_term and a wall-bounded forced RK trajectory (boundary_fill.h:32-133,
    #     source_term.h:25-51)
    ch = {}
    nb, n = (2, 2, 1), (8, 4, 4)
    q = make_state(nb, n, NG, seed=51)
    ch["q"] = q
    cfgw = oracle_cfg(nb, n, NG, periodic=(0, 0, 0))
    for name, o in CHANNEL_BCS.items():
        ch[f"fill_{name}"] = ref.boundary_fill(cfgw, ref.make_bc(**o), q.ravel()).reshape(q.shape)
    rhs0 = np.random.default_rng(5).standard_normal(q.shape)
    ch["src_rhs0"] = rhs0
    ch["src_rhs1"] = ref.source_term(cfgw, ref.make_bc(mask=(0,) * 6, force=CHANNEL_FORCE), q.ravel(), rhs0.ravel()).reshape(q.shape)
    cfgc = oracle_cfg(nb, n, NG, scheme=0, integrator=0, periodic=(1, 0, 1))
    bc = ref.make_bc(force=CHANNEL_FORCE, **CHANNEL_BCS["noslip_isothermal_y"])
    q0 = ref.boundary_fill(cfgc, bc, ref.exchange(cfgc, q.ravel()))
    umax = ref.reduce_umax(cfgc, q0)
    dt = 0.2 * (2 * np.pi / 16) / umax
    q3, _ = ref.advance_channel(cfgc, bc, q0, dt, 3)
    ch["adv_q0"], ch["adv_q3"], ch["adv_dt"] = q0.reshape(q.shape), q3.reshape(q.shape), np.array([dt, umax])
    np.savez_compressed(os.path.join(HERE, "channel_small.npz"), **ch)
    # (vi) AMR (config 5): 2x2x2 roots, block 0 refined in all directions (15 blocks; second case: a child refined again,
    #      which drags its coarse neighbours along through the factor-2 constraint). Block boxes, the reference's own
    #      injection + interpolation tables for 1 and 3 ranks, exchange, flux_div on per-block dx, an RK4 trajectory.
    amr = {}
    for case, roots, n_amr, passes in (("a", (2, 2, 2), (8, 4, 4), [[0]]), ("b", (3, 2, 1), (4, 4, 4), [[0], [0]])):
        ref.set_amr(passes)
        cfg1 = oracle_cfg(roots, n_amr, NG, scheme=0, integrator=0, periodic=(1, 1, 1), nranks=1)
        nblk, boxes = ref.block_boxes(cfg1)
        amr[f"{case}_boxes"] = boxes
        q_in = zero_ghosts(amr_state(boxes, n_amr, NG, seed=61), NG)         # regenerated by the tests from the seed
        qe = ref.exchange(cfg1, q_in.ravel())
        amr[f"{case}_q_ex"] = qe.reshape(q_in.shape)
        if case == "a":
            amr[f"{case}_rhs"] = ref.flux_div(cfg1, qe).reshape(q_in.shape)
            umax = ref.reduce_umax(cfg1, qe)
            dt = 0.2 * (2 * np.pi / 64) / umax
            q2, _ = ref.advance(cfg1, qe, dt, 2)
            amr[f"{case}_q_adv"], amr[f"{case}_dt"] = q2.reshape(q_in.shape), np.array([dt, umax])
        for nranks in (1, 3):
            cfg = oracle_cfg(roots, n_amr, NG, periodic=(1, 1, 1), nranks=nranks)
            assert np.array_equal(ref.exchange(cfg, q_in.ravel()), qe)       # the rank count does not matter
            for rank in range(nranks):
                s_, r_, o_ = ref.exchange_tables(cfg, rank)
                si, ri, oi = ref.interp_tables(cfg, rank)
                for nm, arr in (("send", s_), ("recv", r_), ("offs", o_), ("isend", si), ("irecv", ri), ("ioffs", oi)):
                    amr[f"{case}_{nm}_{nranks}_{rank}"] = arr
        ref.set_amr([])
        print("amr case", case, nblk, "blocks")
    np.savez_compressed(os.path.join(HERE, "amr_small.npz"), **amr)
    make_curvilinear()
    for f in ("hotpath_small.npz", "exchange_tables.npz", "channel_small.npz", "amr_small.npz", "curvilinear_small.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


CURV_BOUNDS = [0.0, 2 * np.pi, -1.0, 1.0, 0.5, 2.0]
CURV_MAPS = {"channel": (("scaled", 2.0), ("tanh", -1.0, 1.0, 0.1, 1.3), None),
             "tanh_quad": (None, ("tanh", -1.0, 1.0, 0.1, 1.3), ("quad",))}


def make_curvilinear():
    """(vii) stretched grids (coords::diagonal_coords): the reference's own flux_div(basic) for the convective functors,
    from oracle/_ref/libspade_ref_curv.so (the reference compiled with the two-declaration repair of
    core/coord_system.h:255,274 described in oracle/ref_driver_curv.cc); its Jacobian / metric tables; an RK4 trajectory."""
    assert ref.curv_available(), "build oracle/_ref first (make -C oracle ref)"
    nb, n = (2, 2, 1), (8, 4, 4)
    cv = {}
    for name, maps in CURV_MAPS.items():
        cd = ref.make_coords(maps)
        for scheme in (3, 5, 7):
            q = make_state(nb, n, NG, seed=200 + scheme, bounds=CURV_BOUNDS)
            cfg = oracle_cfg(nb, n, NG, scheme=scheme, bounds=CURV_BOUNDS)
            # the input is regenerated by the tests from the seed (make_state(nb, n, NG, seed=200 + scheme, bounds=CURV_BOUNDS))
            cv[f"{name}_rhs_{scheme}"] = ref.curv_flux_div(cfg, cd, q.ravel()).reshape(q.shape)
        jac, nrm, xyz = ref.curv_geometry(oracle_cfg(nb, n, NG, scheme=3, bounds=CURV_BOUNDS), cd, 3)
        cv[f"{name}_jac"], cv[f"{name}_nrm"], cv[f"{name}_xyz"] = jac, nrm, xyz
    cd = ref.make_coords(CURV_MAPS["channel"])
    cfg = oracle_cfg(nb, n, NG, scheme=3, integrator=0, bounds=CURV_BOUNDS)
    q0 = ref.exchange(cfg, make_state(nb, n, NG, seed=77, bounds=CURV_BOUNDS).ravel())
    dt = 2e-5
    q2 = ref.curv_advance(cfg, cd, q0, dt, 2)
    shape = cv["channel_rhs_3"].shape
    cv["adv_q2"], cv["adv_dt"] = q2.reshape(shape), np.array([dt])      # q0 = exchange(make_state(seed=77)), regenerated
    np.savez_compressed(os.path.join(HERE, "curvilinear_small.npz"), **cv)


if __name__ == "__main__":
    main()
