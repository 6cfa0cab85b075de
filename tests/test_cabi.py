"""The C-ABI shared library loads on a machine without a GPU, exports every symbol include/spade_b200.h declares,
and refuses compute calls loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "spade_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    from spade_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 25
    lib = _lib.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/spade_b200.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), "python binding table and header disagree"


def test_no_oracle_in_product_path():
    """the product package must not import or link anything under oracle/"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "spade_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle", "").replace("oracle restatement", ""), f
    out = os.popen(f"ldd {os.path.join(ROOT, 'spade_b200', 'libspade_b200.so')}").read()
    assert "oracle" not in out and "spade_ref" not in out


def test_version_and_error_convention():
    from spade_b200 import _lib
    lib = _lib.lib()
    assert b"sm_100a" in lib.spb_version()
    # bad arguments give a nonzero code and a message, never a crash
    rc = lib.spb_exchange_create(None, _lib.int3((1, 1, 1)), _lib.int3((4, 4, 4)), _lib.int3((2, 2, 2)), _lib.int3((1, 1, 1)), 0, 1)
    assert rc != 0 and b"spb_exchange_create" in lib.spb_last_error()
    with pytest.raises(_lib.SpbError):
        _lib.check(rc)


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from spade_b200 import _lib
    lib = _lib.lib()
    assert lib.spb_device_count() == 0
    h = C.c_void_p()
    bbox = np.array([0, 1, 0, 1, 0, 1], dtype=np.float64)
    rc = lib.spb_grid_create(C.byref(h), _lib.int3((8, 8, 8)), _lib.int3((2, 2, 2)), 1, bbox.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 10003 and b"no CUDA device" in lib.spb_last_error()      # SPB_ERR_NO_DEVICE


def test_host_mirror_tables_match_reference_rk_tables():
    """rk_t tables (explicit.h:27-116) and the coefficient differences formed like advance.h:47-55,84-92."""
    import spade_b200.api as sp
    rk4 = sp.rk4_t
    assert [[float(x) for x in r] for r in rk4.table] == [[0, 0, 0, 0], [0.5, 0, 0, 0], [0, 0.5, 0, 0], [0, 0, 1, 0]]
    assert [float(x) for x in rk4.accum] == [1 / 6, 1 / 3, 1 / 3, 1 / 6]
    diffs = [[sp._ratio_diff_value(c, p) for c, p in zip(cur, prev)] for prev, cur in
             zip(rk4.table, rk4.table[1:] + [rk4.accum])]
    assert [sum(1 for d in row if d != 0.0) for row in diffs] == [1, 2, 2, 4]      # 9 residual reads per step
    assert diffs[3] == [1 / 6, 1 / 3, -2 / 3, 1 / 6]       # exact ratio difference, then one rounding (advance.h:47-55)
    for alg in (sp.rk2_t, sp.ssprk3_t, sp.ssprk34_t, sp.rk38r_t, rk4):
        assert abs(float(sum(alg.accum)) - 1.0) < 1e-15                              # consistency of every table
        for row, c in zip(alg.table, alg.dt):
            assert sum(row) == c


def test_shim_recognises_every_functor_type_of_the_implemented_set():
    """integration/api_surface_check.cc (compiled in the dev container against the unmodified reference headers + the shim):
    spade::b200::flux_desc over every functor type; the POD descriptor must carry the functor's own members. Host-only."""
    import subprocess
    exe = os.path.join(ROOT, "integration", "_build", "api_surface_check")
    if not os.path.exists(exe):
        pytest.skip("integration/_build/api_surface_check not built (needs /root/reference at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60, cwd=os.path.dirname(exe))
    assert out.returncode == 0, out.stdout + out.stderr
    rows = {}
    for line in out.stdout.strip().splitlines():
        name, rest = line[:44].strip(), line[44:].split()
        rows[name] = dict(kv.split("=") for kv in rest)
    assert len(rows) == 15
    assert rows["fweno_t"]["linear"] == "0" and rows["fweno_t<disable_smooth>"]["linear"] == "1"
    assert rows["weno_t<rusanov_t, disable_smooth>"]["linear"] == "1" and rows["weno_t<rusanov_t, disable_smooth>"]["conv"] == "3"
    assert rows["totani_lr"]["conv"] == "1" and rows["cent_keep<2>"]["conv"] == "1"
    assert rows["cent_keep<6> + visc_lr"]["conv"] == "4" and rows["cent_keep<8>"]["conv"] == "5"
    assert rows["fweno_t"]["conv"] == "3" and rows["weno_t<rusanov_t>"]["conv"] == "3"
    w = rows["hybrid(totani, fweno, ducros, full) + wale"]
    assert (w["diss"], w["blend"], w["visc"], w["sgs"], w["cw"], w["delta"], w["prt"], w["eps"]) == ("1", "0", "1", "1", "0.55", "0.1", "0.9", "0.01")
    assert rows["hybrid(cent_keep<4>, fweno, ducros, diss)"]["blend"] == "1"
    assert rows["hybrid(totani, weno_t<rusanov_t>, ducros)"]["diss"] == "1"
    assert float(rows["visc_lr"]["beta"]) == pytest.approx(-2.0 * 1.8e-5 / 3.0, rel=1e-5)


def test_fused_stage_plans_reproduce_every_rk_table():
    """Host logic of integrator_t: the per-stage plan of the fused kernel (at most two residual inputs and one output per stage,
    the final combination prepared one stage early) must be algebraically the Butcher table (explicit.h:27-116) for every
    scheme it accepts. Emulated with scalars: w' = lambda(t) w, conserved variable w, no kernels involved."""
    import spade_b200.api as sp

    def rhs(w, t):
        return (-0.7 + 0.3 * np.cos(t)) * w + 0.2 * np.sin(3 * t)

    for alg in (sp.rk2_t, sp.rk4_t, sp.ssprk3_t, sp.ssprk34_t, sp.rk38r_t, sp.rk2hs_t, sp.ssprk3hs_t):
        plan = sp.integrator_t._fused_plan(alg)
        n, dt, t0, w0 = alg.rows(), 0.05, 0.3, 1.7
        # classical form: k_i = f(w0 + dt sum_j a_ij k_j, t0 + c_i dt); w1 = w0 + dt sum_i b_i k_i
        k = []
        for i in range(n):
            wi = w0 + dt * sum(float(alg.table[i][j]) * k[j] for j in range(i))
            k.append(rhs(wi, t0 + float(alg.dt[i]) * dt))
        want = w0 + dt * sum(float(alg.accum[i]) * k[i] for i in range(n))
        if plan is None:
            continue
        # the plan: stage i sees w (the running state), evaluates r = f(w, t_i), then
        #   w <- w + dt (cq_self r + sum cq_a reg[in_a]);   reg[out] <- co_self r + sum co_a reg[in_a]
        w, reg = w0, {}
        for i, st in enumerate(plan):
            r = rhs(w, t0 + float(alg.dt[i]) * dt)
            ins = [reg[key] for key in st["in"]]
            new_w = w + dt * (st["cq_self"] * r + sum(c * x for c, x in zip(st["cq"], ins)))
            if st["out"] is not None:
                reg[st["out"]] = st["co_self"] * r + sum(c * x for c, x in zip(st["co"], ins))
            w = new_w
        assert abs(w - want) < 1e-14 * abs(want), alg.name
    assert sp.integrator_t._fused_plan(sp.rk4_t) is not None and sp.integrator_t._fused_plan(sp.ssprk3_t) is not None


def _levels(nx, boxes):
    from spade_b200 import _lib
    lib = _lib.lib()
    boxes = np.ascontiguousarray(boxes, dtype=np.float64)
    nlb = boxes.shape[0]
    lev_n = (C.c_int * 3)()
    lev_inv = np.zeros((3, 16))
    lev = np.zeros(nlb, dtype=np.int32)
    tol = np.zeros(3)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    _lib.check(lib.spb_grid_spacing_levels(_lib.int3(nx), nlb, dp(boxes), lev_n, dp(lev_inv), lev.ctypes.data_as(C.POINTER(C.c_int)), dp(tol)))
    return list(lev_n), lev_inv, lev, tol


def test_spacing_levels_of_weak_scaled_slabs_amr_grids_and_exotic_lattices():
    """spb_grid_spacing_levels (host-only): every rank of a weak-scaled TGV box is ONE level per direction although the reference's
    box.size/num_cell (cartesian_grid.h:134-135) wobbles by 10-20 ulp on the ranks away from the origin — the case that sent
    ranks 1 and 2 to a slower kernel variant in rounds 1 and 2; an AMR grid has one level per refinement; more than 16 distinct
    spacings along a direction are reported as -1 (the RHS calls then refuse the grid)."""
    import spade_b200.api as sp
    L, eps = 2 * np.pi, 2.220446049250313e-16
    n = 8
    blocks = sp.cartesian_blocks_t((2, 2, 16 * n), [0.0, L, 0.0, L, 0.0, L * n])
    wobble = []
    for r in range(n):
        boxes = np.array([blocks.get_block_box(lb) for lb in range(r * 64, (r + 1) * 64)])
        inv = 1.0 / ((boxes[:, 5] - boxes[:, 4]) / 32)
        wobble.append(np.ptp(inv) / inv[0] / eps)
        lev_n, lev_inv, lev, tol = _levels((32, 32, 32), boxes)
        assert lev_n == [1, 1, 1] and not lev.any()
        assert lev_inv[2, 0] == inv[0] and tol[2] >= wobble[-1] * eps
    assert max(wobble) > 8                                   # the old fixed 8-ulp test failed on such ranks
    # three refinement levels along x and z, two along y, blocks far from the origin
    rng = np.random.default_rng(0)
    boxes, want = [], []
    for _ in range(200):
        l = rng.integers(0, 3, size=3) % np.array([3, 2, 3])
        lo = rng.uniform(-40.0, 40.0, size=3)
        size = 1.0 / 2.0 ** l
        boxes.append([lo[0], lo[0] + size[0], lo[1], lo[1] + size[1], lo[2], lo[2] + size[2]])
        want.append(l)
    lev_n, lev_inv, lev, tol = _levels((16, 16, 16), np.array(boxes))
    assert lev_n == [3, 2, 3]
    for b, l in zip(range(200), want):
        for d in range(3):
            got = lev_inv[d, (lev[b] >> (8 * d)) & 255]
            assert abs(got - 16.0 * 2.0 ** l[d]) <= 1e-13 * got
    # 17 distinct block sizes along y
    boxes = np.array([[0.0, 1.0, 0.0, 1.0 + 0.1 * k, 0.0, 1.0] for k in range(17)])
    lev_n, _, _, _ = _levels((8, 8, 8), boxes)
    assert lev_n == [1, -1, 1]
