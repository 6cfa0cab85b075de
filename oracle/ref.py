"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libspade_ref.so (the unmodified
SPADE reference, built by oracle/Makefile from /root/reference/src + oracle/ref_driver.cc)."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libspade_ref.so")


class RefCfg(C.Structure):
    _fields_ = [("nblocks", C.c_int * 3), ("ncells", C.c_int * 3), ("ng", C.c_int),
                ("bounds", C.c_double * 6), ("periodic", C.c_int * 3), ("scheme", C.c_int),
                ("gamma", C.c_double), ("R", C.c_double), ("mu", C.c_double),
                ("prandtl", C.c_double), ("sensor_eps", C.c_double), ("nranks", C.c_int),
                ("integrator", C.c_int), ("sgs_cw", C.c_double), ("sgs_delta", C.c_double), ("sgs_prt", C.c_double)]


class RefBc(C.Structure):
    """ref_bc / spo_bc: boundary_fill kernel + body force of a channel run."""
    _fields_ = [("mask", C.c_int * 6), ("kind", C.c_int), ("order", C.c_int), ("a", C.c_double * 5), ("b", C.c_double * 5),
                ("use_normal", C.c_int), ("a_normal", C.c_double), ("force", C.c_double * 3)]


def make_bc(mask, kind=0, order=0, a=(1, 1, 1, 1, 1), b=(0, 0, 0, 0, 0), a_normal=None, force=(0, 0, 0)):
    bc = RefBc()
    bc.mask[:] = [int(m) for m in mask]
    bc.kind, bc.order = int(kind), int(order)
    bc.a[:] = [float(x) for x in a]
    bc.b[:] = [float(x) for x in b]
    bc.use_normal = 0 if a_normal is None else 1
    bc.a_normal = 0.0 if a_normal is None else float(a_normal)
    bc.force[:] = [float(x) for x in force]
    return bc


LIB_PATH_O3 = os.path.join(_HERE, "_ref", "libspade_ref_o3.so")
BUILD_FLAGS = {LIB_PATH: "g++ -std=c++20 -O2 -ffp-contract=off", LIB_PATH_O3: "g++ -std=c++20 -O3"}


def available():
    return os.path.exists(LIB_PATH)


_lib = None
_lib_path = LIB_PATH


def use_timing_build():
    """bench.py's CPU legs time the -O3 build of the same driver (oracle/Makefile) when it exists; returns the flags of
    the library that will be loaded. Must be called before the first lib() of the process."""
    global _lib_path
    assert _lib is None, "use_timing_build() after the library was loaded"
    if os.path.exists(LIB_PATH_O3):
        _lib_path = LIB_PATH_O3
    return BUILD_FLAGS[_lib_path]


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_lib_path)
        _lib.ref_last_error.restype = C.c_char_p
        _lib.ref_array_size.restype = C.c_int64
    return _lib


def make_cfg(nblocks, ncells, ng=2, bounds=None, periodic=(1, 1, 1), scheme=0, gamma=1.4, R=287.15,
             mu=1.0e-3, prandtl=0.72, sensor_eps=1e-2, nranks=1, integrator=0, sgs=(0.55, 0.1, 0.9)):
    c = RefCfg()
    c.nblocks[:] = list(nblocks)
    c.ncells[:] = list(ncells)
    c.ng = ng
    if bounds is None:
        bounds = [0.0, 2 * np.pi] * 3
    c.bounds[:] = list(bounds)
    c.periodic[:] = [int(p) for p in periodic]
    c.scheme = scheme
    c.gamma, c.R, c.mu, c.prandtl, c.sensor_eps = gamma, R, mu, prandtl, sensor_eps
    c.nranks = nranks
    c.integrator = integrator
    c.sgs_cw, c.sgs_delta, c.sgs_prt = [float(x) for x in sgs]
    return c


def _check(rc):
    if rc != 0:
        raise RuntimeError("reference driver failed: " + lib().ref_last_error().decode())


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def array_size(cfg):
    return int(lib().ref_array_size(C.byref(cfg)))


def flux_div(cfg, q, rhs=None, increment=False):
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = np.zeros_like(q) if rhs is None else np.array(rhs, dtype=np.float64, copy=True)
    _check(lib().ref_flux_div(C.byref(cfg), _ptr(q), _ptr(out), int(increment)))
    return out


def exchange(cfg, q):
    out = np.array(q, dtype=np.float64, copy=True)
    _check(lib().ref_exchange(C.byref(cfg), _ptr(out)))
    return out


def reduce_umax(cfg, q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = C.c_double(0.0)
    _check(lib().ref_reduce_umax(C.byref(cfg), _ptr(q), C.byref(out)))
    return out.value


def advance(cfg, q, dt, nsteps):
    out = np.array(q, dtype=np.float64, copy=True)
    sec = C.c_double(0.0)
    _check(lib().ref_advance(C.byref(cfg), _ptr(out), C.c_double(dt), int(nsteps), C.byref(sec)))
    return out, sec.value


def output_vtk(cfg, q, out_dir, basename):
    """the reference's io::output_vtk of the array holding q (io/io_vtk.h:276-288)"""
    q = np.ascontiguousarray(q, dtype=np.float64)
    _check(lib().ref_output_vtk(C.byref(cfg), _ptr(q), str(out_dir).encode(), str(basename).encode()))


def advance_generic(cfg, q, dt, nsteps, high_storage=False):
    out = np.array(q, dtype=np.float64, copy=True)
    _check(lib().ref_advance_generic(C.byref(cfg), _ptr(out), C.c_double(dt), int(nsteps), int(bool(high_storage))))
    return out


def boundary_fill(cfg, bc, q):
    out = np.array(q, dtype=np.float64, copy=True)
    _check(lib().ref_boundary_fill(C.byref(cfg), C.byref(bc), _ptr(out)))
    return out


def source_term(cfg, bc, q, rhs):
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = np.array(rhs, dtype=np.float64, copy=True)
    _check(lib().ref_source_term(C.byref(cfg), C.byref(bc), _ptr(q), _ptr(out)))
    return out


def advance_channel(cfg, bc, q, dt, nsteps):
    out = np.array(q, dtype=np.float64, copy=True)
    sec = C.c_double(0.0)
    _check(lib().ref_advance_channel(C.byref(cfg), C.byref(bc), _ptr(out), C.c_double(dt), int(nsteps), C.byref(sec)))
    return out, sec.value


def set_amr(passes):
    """AMR: the following calls build the grid on amr_blocks_t and refine, pass by pass, the listed global blocks in all
    directions (constraint factor2). `passes` = list of lists of global block ids; [] returns to the uniform lattice.
    Buffers are then sized by block_boxes()[0] blocks in global block order."""
    counts = np.array([len(p) for p in passes], dtype=np.int64)
    ids = np.array([b for p in passes for b in p], dtype=np.int64)
    i64 = C.POINTER(C.c_int64)
    lib().ref_set_amr(len(passes), counts.ctypes.data_as(i64), ids.ctypes.data_as(i64))


def block_boxes(cfg, cap=4096):
    n = C.c_int64(0)
    boxes = np.zeros((cap, 6))
    _check(lib().ref_block_boxes(C.byref(cfg), C.byref(n), _ptr(boxes), C.c_int64(cap)))
    assert n.value <= cap
    return int(n.value), boxes[:n.value].copy()


def interp_tables(cfg, rank, cap=1 << 16):
    send = np.zeros((cap, 26), dtype=np.int64)
    recv = np.zeros((cap, 26), dtype=np.int64)
    ns, nr = C.c_int64(0), C.c_int64(0)
    offs = np.zeros((cfg.nranks, 6), dtype=np.int64)
    i64 = C.POINTER(C.c_int64)
    _check(lib().ref_interp_tables(C.byref(cfg), int(rank), send.ctypes.data_as(i64), recv.ctypes.data_as(i64),
                                   C.c_int64(cap), C.byref(ns), C.byref(nr), offs.ctypes.data_as(i64)))
    assert ns.value <= cap and nr.value <= cap
    return send[:ns.value].copy(), recv[:nr.value].copy(), offs


def exchange_tables(cfg, rank, cap=1 << 16):
    send = np.zeros((cap, 16), dtype=np.int64)
    recv = np.zeros((cap, 16), dtype=np.int64)
    ns, nr = C.c_int64(0), C.c_int64(0)
    offs = np.zeros((cfg.nranks, 6), dtype=np.int64)
    i64 = C.POINTER(C.c_int64)
    _check(lib().ref_exchange_tables(C.byref(cfg), int(rank), send.ctypes.data_as(i64), recv.ctypes.data_as(i64),
                                     C.c_int64(cap), C.byref(ns), C.byref(nr), offs.ctypes.data_as(i64)))
    assert ns.value <= cap and nr.value <= cap
    return send[:ns.value].copy(), recv[:nr.value].copy(), offs


def prim2cons(gamma, R, p):
    p = np.ascontiguousarray(p, dtype=np.float64)
    w = np.zeros(5)
    lib().ref_prim2cons(C.c_double(gamma), C.c_double(R), _ptr(p), _ptr(w))
    return w


def cons2prim(gamma, R, w):
    w = np.ascontiguousarray(w, dtype=np.float64)
    p = np.zeros(5)
    lib().ref_cons2prim(C.c_double(gamma), C.c_double(R), _ptr(w), _ptr(p))
    return p


# ---- curvilinear (coords::diagonal_coords) convective path: oracle/_ref/libspade_ref_curv.so, see ref_driver_curv.cc ----
CURV_LIB_PATH = os.path.join(_HERE, "_ref", "libspade_ref_curv.so")
COORD_IDENTITY, COORD_SCALED, COORD_TANH, COORD_QUAD = 0, 1, 2, 3


class RefCoords(C.Structure):
    """ref_coords (and the head of spo_coords): one 1-D mapping per direction."""
    _fields_ = [("kind", C.c_int * 3), ("par", (C.c_double * 4) * 3), ("metric_at_physical", C.c_int)]


def make_coords(maps, metric_at_physical=True):
    """maps: three entries, each None / ("scaled", k) / ("tanh", y0, y1, inflation, rate) / ("quad",)."""
    names = {"identity": COORD_IDENTITY, "scaled": COORD_SCALED, "tanh": COORD_TANH, "quad": COORD_QUAD}
    cd = RefCoords()
    for d, m in enumerate(maps):
        m = ("identity",) if m is None else tuple(m)
        cd.kind[d] = names[m[0]]
        for i, x in enumerate(m[1:]):
            cd.par[d][i] = float(x)
    cd.metric_at_physical = int(bool(metric_at_physical))
    return cd


def curv_available():
    return os.path.exists(CURV_LIB_PATH)


_clib = None


def curv_lib():
    global _clib
    if _clib is None:
        _clib = C.CDLL(CURV_LIB_PATH)
        _clib.refc_last_error.restype = C.c_char_p
        _clib.refc_map.restype = C.c_double
        _clib.refc_deriv.restype = C.c_double
        _clib.refc_map.argtypes = [C.POINTER(RefCoords), C.c_int, C.c_double]
        _clib.refc_deriv.argtypes = [C.POINTER(RefCoords), C.c_int, C.c_double]
    return _clib


def _ccheck(rc):
    if rc != 0:
        raise RuntimeError("curvilinear reference driver failed: " + curv_lib().refc_last_error().decode())


def curv_map(cd, d, x):
    return float(curv_lib().refc_map(C.byref(cd), int(d), float(x)))


def curv_deriv(cd, d, x):
    return float(curv_lib().refc_deriv(C.byref(cd), int(d), float(x)))


def curv_flux_div(cfg, cd, q, rhs=None, increment=False):
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = np.zeros_like(q) if rhs is None else np.array(rhs, dtype=np.float64, copy=True)
    _ccheck(curv_lib().refc_flux_div(C.byref(cfg), C.byref(cd), _ptr(q), _ptr(out), int(increment)))
    return out


def curv_geometry(cfg, cd, lb):
    n = [int(x) for x in cfg.ncells]
    jac = np.zeros((n[2], n[1], n[0]))
    nrm = np.zeros((n[2], n[1], n[0], 3))
    xyz = np.zeros((n[2], n[1], n[0], 3))
    _ccheck(curv_lib().refc_geometry(C.byref(cfg), C.byref(cd), C.c_int64(lb), _ptr(jac), _ptr(nrm), _ptr(xyz)))
    return jac, nrm, xyz


def curv_advance(cfg, cd, q, dt, nsteps):
    out = np.array(q, dtype=np.float64, copy=True)
    _ccheck(curv_lib().refc_advance(C.byref(cfg), C.byref(cd), _ptr(out), C.c_double(dt), int(nsteps)))
    return out
