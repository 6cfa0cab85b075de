"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/liboracle.so (the plain-C restatement,
oracle/spade_oracle.c). Built on demand with gcc (a few seconds)."""
import ctypes as C
import os
import subprocess
import numpy as np
from .ref import RefBc, RefCfg, RefCoords, make_bc, make_cfg, make_coords  # same struct layouts (spo_cfg == ref_cfg, spo_bc == ref_bc, spo_coords == RefCoords)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "spade_oracle.c")
    hdr = os.path.join(_HERE, "spade_oracle.h")
    if (force or not os.path.exists(LIB_PATH)
            or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                               "-o", LIB_PATH, src, "-lm"])


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.spo_array_size.restype = C.c_int64
        _lib.spo_offset.restype = C.c_int64
        _lib.spo_offset.argtypes = [C.POINTER(RefCfg), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64]
        for f in (_lib.spo_coord_map, _lib.spo_coord_deriv):
            f.restype = C.c_double
            f.argtypes = [C.POINTER(RefCoords), C.c_int, C.c_double]
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def array_size(cfg):
    return int(lib().spo_array_size(C.byref(cfg)))


def offset(cfg, v, i, j, k, lb):
    return int(lib().spo_offset(C.byref(cfg), v, i, j, k, lb))


def flux_div(cfg, q, rhs=None, increment=False):
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = np.zeros_like(q) if rhs is None else np.array(rhs, dtype=np.float64, copy=True)
    assert lib().spo_flux_div(C.byref(cfg), _ptr(q), _ptr(out), int(increment)) == 0
    return out


def exchange(cfg, q):
    out = np.array(q, dtype=np.float64, copy=True)
    assert lib().spo_exchange(C.byref(cfg), _ptr(out)) == 0
    return out


def reduce_umax(cfg, q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = C.c_double(0.0)
    assert lib().spo_reduce_umax(C.byref(cfg), _ptr(q), C.byref(out)) == 0
    return out.value


def advance(cfg, q, dt, nsteps):
    out = np.array(q, dtype=np.float64, copy=True)
    assert lib().spo_advance(C.byref(cfg), _ptr(out), C.c_double(dt), int(nsteps)) == 0
    return out


def advance_generic(cfg, q, dt, nsteps, high_storage=False):
    out = np.array(q, dtype=np.float64, copy=True)
    assert lib().spo_advance_generic(C.byref(cfg), _ptr(out), C.c_double(dt), int(nsteps), int(bool(high_storage))) == 0
    return out


def boundary_fill(cfg, bc, q):
    out = np.array(q, dtype=np.float64, copy=True)
    assert lib().spo_boundary_fill(C.byref(cfg), C.byref(bc), _ptr(out)) == 0
    return out


def source_term(cfg, bc, q, rhs):
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = np.array(rhs, dtype=np.float64, copy=True)
    assert lib().spo_source_term(C.byref(cfg), C.byref(bc), _ptr(q), _ptr(out)) == 0
    return out


def advance_channel(cfg, bc, q, dt, nsteps):
    out = np.array(q, dtype=np.float64, copy=True)
    assert lib().spo_advance_channel(C.byref(cfg), C.byref(bc), _ptr(out), C.c_double(dt), int(nsteps)) == 0
    return out


_coords_keep = None


def set_coords(cd=None):
    """General coordinates (coords::diagonal_coords) for the following calls; set_coords() returns to coords::identity."""
    global _coords_keep
    _coords_keep = cd
    lib().spo_set_coords(None if cd is None else C.byref(cd))


def coord_map(cd, d, x):
    return float(lib().spo_coord_map(C.byref(cd), int(d), float(x)))


def coord_deriv(cd, d, x):
    return float(lib().spo_coord_deriv(C.byref(cd), int(d), float(x)))


class SpoAmr(C.Structure):
    _fields_ = [("nblocks", C.c_int64), ("bbox", C.POINTER(C.c_double)), ("inj", C.POINTER(C.c_int64)), ("ninj", C.c_int64),
                ("itp", C.POINTER(C.c_int64)), ("nitp", C.c_int64)]


_amr_keep = None


def set_amr(boxes=None, inj=None, itp=None):
    """AMR grid description for the following calls (block boxes + the reference's own transaction tables, all ranks'
    send lists concatenated); set_amr() returns to the uniform lattice."""
    global _amr_keep
    if boxes is None:
        lib().spo_set_amr(None)
        _amr_keep = None
        return
    boxes = np.ascontiguousarray(boxes, dtype=np.float64)
    inj = np.ascontiguousarray(inj, dtype=np.int64).reshape(-1, 16)
    itp = np.ascontiguousarray(itp, dtype=np.int64).reshape(-1, 26)
    a = SpoAmr()
    a.nblocks = boxes.shape[0]
    a.bbox = boxes.ctypes.data_as(C.POINTER(C.c_double))
    a.inj, a.ninj = inj.ctypes.data_as(C.POINTER(C.c_int64)), inj.shape[0]
    a.itp, a.nitp = itp.ctypes.data_as(C.POINTER(C.c_int64)), itp.shape[0]
    _amr_keep = (boxes, inj, itp, a)
    lib().spo_set_amr(C.byref(a))


def exchange_tables(cfg, rank, cap=1 << 16):
    send = np.zeros((cap, 16), dtype=np.int64)
    recv = np.zeros((cap, 16), dtype=np.int64)
    ns, nr = C.c_int64(0), C.c_int64(0)
    offs = np.zeros((cfg.nranks, 6), dtype=np.int64)
    i64 = C.POINTER(C.c_int64)
    assert lib().spo_exchange_tables(C.byref(cfg), int(rank), send.ctypes.data_as(i64), recv.ctypes.data_as(i64),
                                     C.c_int64(cap), C.byref(ns), C.byref(nr), offs.ctypes.data_as(i64)) == 0
    assert ns.value <= cap and nr.value <= cap
    return send[:ns.value].copy(), recv[:nr.value].copy(), offs


def partition(nglob, nranks):
    g2r = np.zeros(nglob, dtype=np.int64)
    g2l = np.zeros(nglob, dtype=np.int64)
    i64 = C.POINTER(C.c_int64)
    lib().spo_partition(C.c_int64(nglob), int(nranks), g2r.ctypes.data_as(i64), g2l.ctypes.data_as(i64))
    return g2r, g2l


def prim2cons(gamma, R, p):
    p = np.ascontiguousarray(p, dtype=np.float64)
    w = np.zeros(5)
    lib().spo_prim2cons(C.c_double(gamma), C.c_double(R), _ptr(p), _ptr(w))
    return w


def cons2prim(gamma, R, w):
    w = np.ascontiguousarray(w, dtype=np.float64)
    p = np.zeros(5)
    lib().spo_cons2prim(C.c_double(gamma), C.c_double(R), _ptr(w), _ptr(p))
    return p
