"""AMR grids (BASELINE config 5; reference src/grid/get_transaction.h:100-263, transactions.h:136-234,
make_exchange.h:208-314). The refinement logic stays in SPADE; what crosses the boundary is the block boxes and the
transaction tables of SPADE's own exchange_config_t. The golden fixture (tests/golden/amr_small.npz, written by
make_golden.py from the unmodified reference) holds exactly that for two refined grids, 1 and 3 ranks, plus the reference's
exchange / flux_div / RK4 results. CPU half: the oracle and the host side of the plan; GPU half: the kernels."""
import ctypes as C
import os

import numpy as np
import pytest

from util import GAMMA, RGAS, amr_state, oracle_cfg, product_flux, rel_l2, zero_ghosts

HERE = os.path.dirname(os.path.abspath(__file__))
NG = 2
CASES = {"a": ((2, 2, 2), (8, 4, 4)), "b": ((3, 2, 1), (4, 4, 4))}


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "amr_small.npz"))


def _q_in(gold, case):
    return zero_ghosts(amr_state(gold[f"{case}_boxes"], CASES[case][1], NG, seed=61), NG)


def _all_ranks(gold, case, name, nranks=3):
    return np.concatenate([gold[f"{case}_{name}_{nranks}_{r}"] for r in range(nranks)], axis=0)


# ---------------------------------------------------------------- CPU: oracle and plan bookkeeping
@pytest.mark.parametrize("case", ["a", "b"])
def test_oracle_amr_exchange_bit_exact(gold, case):
    from oracle import port
    roots, n = CASES[case]
    cfg = oracle_cfg(roots, n, NG)
    q = _q_in(gold, case)
    for nranks in (1, 3):        # the union of the per-rank send lists is the whole exchange, whatever the partition
        port.set_amr(gold[f"{case}_boxes"], _all_ranks(gold, case, "send", nranks), _all_ranks(gold, case, "isend", nranks))
        try:
            got = port.exchange(cfg, q.ravel()).reshape(q.shape)
        finally:
            port.set_amr()
        assert np.array_equal(got, gold[f"{case}_q_ex"])


def test_oracle_amr_flux_div_and_trajectory(gold):
    from oracle import port
    roots, n = CASES["a"]
    cfg = oracle_cfg(roots, n, NG, scheme=0, integrator=0)
    port.set_amr(gold["a_boxes"], gold["a_send_1_0"], gold["a_isend_1_0"])
    try:
        qe = gold["a_q_ex"]
        assert rel_l2(port.flux_div(cfg, qe.ravel()).reshape(qe.shape), gold["a_rhs"]) < 1e-14
        dt, umax = gold["a_dt"]
        assert port.reduce_umax(cfg, qe.ravel()) == umax
        assert rel_l2(port.advance(cfg, qe.ravel(), float(dt), 2).reshape(qe.shape), gold["a_q_adv"]) < 1e-14
    finally:
        port.set_amr()


@pytest.mark.parametrize("case", ["a", "b"])
@pytest.mark.parametrize("nranks", [1, 3])
def test_plan_from_tables_message_layout(gold, case, nranks):
    """host side of spb_exchange_create_from_tables + spb_exchange_add_interp (no GPU): message sizes per peer are the
    reference's injec_offsets + intrp_offsets, and the lists keep the reference's order"""
    from spade_b200._lib import lib, check, int3
    roots, n = CASES[case]
    i64 = C.POINTER(C.c_int64)
    for rank in range(nranks):
        t = {k: np.ascontiguousarray(gold[f"{case}_{k}_{nranks}_{rank}"]) for k in ("send", "recv", "isend", "irecv", "offs", "ioffs")}
        h = C.c_void_p()
        check(lib().spb_exchange_create_from_tables(C.byref(h), int3(n), int3((NG,) * 3), rank, nranks, t["send"].ctypes.data_as(i64),
                                                    len(t["send"]), t["recv"].ctypes.data_as(i64), len(t["recv"])))
        check(lib().spb_exchange_add_interp(h, t["isend"].ctypes.data_as(i64), len(t["isend"]), t["irecv"].ctypes.data_as(i64), len(t["irecv"])))
        assert lib().spb_exchange_num_interp_send(h) == len(t["isend"]) and lib().spb_exchange_num_interp_recv(h) == len(t["irecv"])
        for p in range(nranks):
            assert lib().spb_exchange_send_cells(h, p) == t["offs"][p, 0] + t["ioffs"][p, 0]
            assert lib().spb_exchange_recv_cells(h, p) == t["offs"][p, 1] + t["ioffs"][p, 1]
        send = np.zeros_like(t["send"]); recv = np.zeros_like(t["recv"])
        check(lib().spb_exchange_tables(h, send.ctypes.data_as(i64), recv.ctypes.data_as(i64), None))
        assert np.array_equal(send, t["send"]) and np.array_equal(recv, t["recv"])       # already in the reference's order
        lib().spb_exchange_destroy(h)


@pytest.mark.parametrize("pre,nranks", [("p", 2), ("p", 4), ("b", 4), ("b", 8)])
def test_boundary_blocks_include_the_donors_of_off_rank_interpolation_sends(pre, nranks):
    """Host side of the overlapped schedule (no GPU): spb_exchange_boundary_blocks must mark the source block of EVERY off-rank
    send transaction, injection and patch_fill_t alike — a donor block that only feeds an off-rank interpolation would otherwise
    be advanced while its message is being packed. Tables: the reference's, for the parity grid (p) and the bench
    grid (b) of config 5; p on 2 ranks and b on 4 / 8 ranks hold such interpolation-only donors (15 / 152 / 168 blocks)."""
    from spade_b200._lib import lib, check, int3
    fix = np.load(os.path.join(HERE, "golden", "config5_amr.npz"))
    n = tuple(int(x) for x in fix[f"{pre}_cells"])
    nblocks = len(fix[f"{pre}_boxes"])
    i64 = C.POINTER(C.c_int64)
    seen_interp_only = False
    for rank in range(nranks):
        t = {k: np.ascontiguousarray(fix[f"{pre}_{k}_{nranks}_{rank}"], dtype=np.int64) for k in ("send", "recv", "isend", "irecv")}
        per, extra = divmod(nblocks, nranks)
        nloc = per + (1 if rank < extra else 0)
        h = C.c_void_p()
        check(lib().spb_exchange_create_from_tables(C.byref(h), int3(n), int3((NG,) * 3), rank, nranks, t["send"].ctypes.data_as(i64),
                                                    len(t["send"]), t["recv"].ctypes.data_as(i64), len(t["recv"])))
        check(lib().spb_exchange_add_interp(h, t["isend"].ctypes.data_as(i64), len(t["isend"]), t["irecv"].ctypes.data_as(i64), len(t["irecv"])))
        mask = np.zeros(max(nloc, 1), dtype=np.uint8)
        check(lib().spb_exchange_boundary_blocks(h, nloc, mask.ctypes.data_as(C.POINTER(C.c_ubyte))))
        inj = {int(r[8]) for r in t["send"] if r[2] != rank}
        itp = {int(r[8]) for r in t["isend"] if r[2] != rank}
        assert set(np.flatnonzero(mask[:nloc])) == inj | itp
        seen_interp_only = seen_interp_only or bool(itp - inj)
        lib().spb_exchange_destroy(h)
    assert seen_interp_only == ((pre, nranks) != ("p", 4))


# ---------------------------------------------------------------- GPU
def _rank_slices(nblocks, nranks):
    from oracle import port
    g2r, _ = port.partition(nblocks, nranks)
    return [(int(np.argmax(g2r == r)), int((g2r == r).sum())) for r in range(nranks)]


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["a", "b"])
def test_gpu_amr_exchange_one_rank_bit_exact(gold, case):
    import spade_b200.api as sp
    roots, n = CASES[case]
    grid = sp.cartesian_grid_t.from_boxes(n, gold[f"{case}_boxes"])
    qa = sp.grid_array.from_host(grid, _q_in(gold, case))
    ex = sp.make_exchange(qa, (1, 1, 1), tables=tuple(gold[f"{case}_{k}_1_0"] for k in ("send", "recv", "isend", "irecv")))
    ex.exchange(qa)
    assert np.array_equal(qa.to_host(), gold[f"{case}_q_ex"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["a", "b"])
def test_gpu_amr_exchange_three_simulated_ranks_bit_exact(gold, case):
    """pack (injection + interpolation sections) on each sender, hand the message over, unpack on the receiver"""
    import torch
    import spade_b200.api as sp
    roots, n = CASES[case]
    boxes, q = gold[f"{case}_boxes"], _q_in(gold, case)
    ranks = []
    for r, (lo, cnt) in enumerate(_rank_slices(len(boxes), 3)):
        grid = sp.cartesian_grid_t.from_boxes(n, boxes[lo:lo + cnt], sp.pool_t(r, 3), first_block=lo)
        qa = sp.grid_array.from_host(grid, q[lo:lo + cnt])
        ranks.append((qa, sp.make_exchange(qa, (1, 1, 1), tables=tuple(gold[f"{case}_{k}_3_{r}"] for k in ("send", "recv", "isend", "irecv")))))
    lib, bufs = sp.lib(), {}
    for r, (qa, ex) in enumerate(ranks):
        for p in range(3):
            if p != r and ex.send_cells[p]:
                b = torch.empty(5 * ex.send_cells[p], dtype=torch.float64, device="cuda")
                sp.check(lib.spb_exchange_pack(ex._h, C.c_void_p(qa.data.data_ptr()), p, C.c_void_p(b.data_ptr()), None))
                bufs[(r, p)] = b
        sp.check(lib.spb_exchange_local(ex._h, C.c_void_p(qa.data.data_ptr()), None))
    for r, (qa, ex) in enumerate(ranks):
        for p in range(3):
            if p != r and ex.recv_cells[p]:
                assert ex.recv_cells[p] == ranks[p][1].send_cells[r]
                sp.check(lib.spb_exchange_unpack(ex._h, C.c_void_p(qa.data.data_ptr()), p, C.c_void_p(bufs[(p, r)].data_ptr()), None))
    torch.cuda.synchronize()
    assert np.array_equal(np.concatenate([qa.to_host() for qa, _ in ranks], axis=0), gold[f"{case}_q_ex"])


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
def test_gpu_amr_flux_div_and_rk4_trajectory(gold, fused):
    """per-block spacing (fine blocks next to coarse ones) in the RHS kernel, and two RK4 steps with the AMR exchange;
    the fused stage kernel stores the same-level (injection) ghosts itself and leaves only the interpolation transactions
    to a separate kernel (spb_exchange_local_interp)"""
    import spade_b200.api as sp
    roots, n = CASES["a"]
    grid = sp.cartesian_grid_t.from_boxes(n, gold["a_boxes"])
    gas = sp.ideal_gas_t(GAMMA, RGAS)
    flux = sp.flux_desc(product_flux(0))
    qa, ra = sp.grid_array.from_host(grid, gold["a_q_ex"]), sp.grid_array(grid, 0.0)
    sp.flux_div(qa, ra, flux, sp.overwrite)
    assert rel_l2(ra.to_host(), gold["a_rhs"]) < 1e-12
    assert sp.transform_reduce(qa, sp.FN_WAVESPEED, sp.RED_MAX, gas) == pytest.approx(float(gold["a_dt"][1]), rel=1e-15)
    ex = sp.make_exchange(qa, (1, 1, 1), tables=tuple(gold[f"a_{k}_1_0"] for k in ("send", "recv", "isend", "irecv")))
    rhs = sp.flux_div_rhs_t(flux, sp.overwrite) if fused else (lambda r, qq, t: sp.flux_div(qq, r, flux, sp.overwrite))
    bc = sp.exchange_bc_t(ex) if fused else (lambda qq, t: ex.exchange(qq))
    ti = sp.integrator_t(sp.time_axis_t(0.0, float(gold["a_dt"][0])), sp.rk4_t, sp.integrator_data_t(qa, ra, sp.rk4_t), rhs, bc,
                         sp.state_transform_t(gas))
    for _ in range(2):
        ti.advance()
    assert (ti._plan is not None) == fused
    if fused:
        assert ti._fuse_exchange
    assert rel_l2(ti.solution().to_host(), gold["a_q_adv"]) < 1e-12
