"""ctypes loader of libspade_b200.so (the C ABI declared in include/spade_b200.h).

There is no CPU fallback: if the shared library is missing the import of any compute entry point
fails loudly, and every compute call fails if no CUDA device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPB_B200_LIB: development override (timing experiments build differently named copies of the same library)
LIB_PATH = os.environ.get("SPB_B200_LIB") or os.path.join(_HERE, "libspade_b200.so")


class FluxDesc(C.Structure):
    """spb_flux_desc (include/spade_b200.h)."""
    _fields_ = [("conv", C.c_int), ("diss", C.c_int), ("blend", C.c_int), ("visc", C.c_int),
                ("gamma", C.c_double), ("R", C.c_double), ("mu", C.c_double), ("beta", C.c_double),
                ("prandtl_inv", C.c_double), ("sensor_eps", C.c_double),
                ("sgs", C.c_int), ("sgs_cw", C.c_double), ("sgs_delta", C.c_double), ("sgs_prt", C.c_double),
                ("weno_linear", C.c_int)]


class StageDesc(C.Structure):
    """spb_stage_desc (include/spade_b200.h)."""
    _fields_ = [("nin", C.c_int), ("inp", C.c_void_p * 2), ("cq_self", C.c_double), ("cq", C.c_double * 2),
                ("out", C.c_void_p), ("co_self", C.c_double), ("co", C.c_double * 2)]


class StagePlan(C.Structure):
    """spb_stage_plan (include/spade_b200.h)."""
    _fields_ = [("nin", C.c_int), ("inp", C.c_int * 2), ("cq", C.c_double * 2), ("co", C.c_double * 2),
                ("cq_self", C.c_double), ("co_self", C.c_double), ("out", C.c_int)]


SPB_ERR_BAD_ARG, SPB_ERR_UNSUPPORTED, SPB_ERR_NO_DEVICE, SPB_ERR_DRIVER = 10001, 10002, 10003, 10004
SPB_PART_ALL, SPB_PART_BOUNDARY, SPB_PART_INTERIOR = 0, 1, 2


class BcDesc(C.Structure):
    """spb_bc_desc (include/spade_b200.h)."""
    _fields_ = [("kind", C.c_int), ("order", C.c_int), ("a", C.c_double * 5), ("b", C.c_double * 5),
                ("use_normal", C.c_int), ("a_normal", C.c_double)]


class SourceDesc(C.Structure):
    """spb_source_desc (include/spade_b200.h)."""
    _fields_ = [("kind", C.c_int), ("f", C.c_double * 5)]


class MetricDesc(C.Structure):
    """spb_metric_desc (include/spade_b200.h)."""
    _fields_ = [("area", C.POINTER(C.c_double) * 3), ("jac", C.POINTER(C.c_double) * 3), ("face", C.POINTER(C.c_double) * 3)]


class SpbError(RuntimeError):
    pass


_lib = None
_i64p = C.POINTER(C.c_int64)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

# name -> (restype, argtypes); every symbol include/spade_b200.h declares
SYMBOLS = {
    "spb_grid_create": (C.c_int, [C.POINTER(C.c_void_p), _ip, _ip, C.c_int64, _dp]),
    "spb_grid_spacing_levels": (C.c_int, [_ip, C.c_int64, _dp, _ip, _dp, _ip, _dp]),
    "spb_grid_destroy": (None, [C.c_void_p]),
    "spb_grid_array_size": (C.c_int64, [C.c_void_p]),
    "spb_grid_offset": (C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64]),
    "spb_grid_set_metric": (C.c_int, [C.c_void_p, C.POINTER(MetricDesc)]),
    "spb_grid_has_metric": (C.c_int, [C.c_void_p]),
    "spb_flux_div": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(FluxDesc), C.c_int, C.c_void_p]),
    "spb_flux_div_blocks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(FluxDesc), C.c_int,
                                      C.c_int64, C.c_int64, C.c_void_p]),
    "spb_flux_div_rk_stage": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(FluxDesc), C.POINTER(StageDesc),
                                        C.c_int64, C.c_int64, C.c_void_p]),
    "spb_flux_div_rk_stage_exchange": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(FluxDesc), C.POINTER(StageDesc),
                                                 C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "spb_rk_update": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, _dp, C.c_double,
                                C.c_double, C.c_void_p]),
    "spb_flux_div_rk_stage_supported": (C.c_int, [C.c_void_p]),
    "spb_flux_div_rk_stage_part": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "spb_rk_fused_plan": (C.c_int, [C.c_int, _dp, C.POINTER(StagePlan)]),
    "spb_axpy_roundtrip": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p]),
    "spb_ssprk3_stage": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                   C.c_double, C.c_double, C.c_void_p]),
    "spb_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _dp,
                             C.c_void_p]),
    "spb_exchange_create": (C.c_int, [C.POINTER(C.c_void_p), _ip, _ip, _ip, _ip, C.c_int, C.c_int]),
    "spb_exchange_create_from_tables": (C.c_int, [C.POINTER(C.c_void_p), _ip, _ip, C.c_int, C.c_int, _i64p,
                                                  C.c_int64, _i64p, C.c_int64]),
    "spb_exchange_add_interp": (C.c_int, [C.c_void_p, _i64p, C.c_int64, _i64p, C.c_int64]),
    "spb_exchange_num_interp_send": (C.c_int64, [C.c_void_p]),
    "spb_exchange_num_interp_recv": (C.c_int64, [C.c_void_p]),
    "spb_exchange_destroy": (None, [C.c_void_p]),
    "spb_exchange_num_send": (C.c_int64, [C.c_void_p]),
    "spb_exchange_num_recv": (C.c_int64, [C.c_void_p]),
    "spb_exchange_local_interp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "spb_exchange_boundary_blocks": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_ubyte)]),
    "spb_exchange_tables": (C.c_int, [C.c_void_p, _i64p, _i64p, _i64p]),
    "spb_exchange_local_blocks": (C.c_int64, [C.c_void_p]),
    "spb_exchange_first_block": (C.c_int64, [C.c_void_p]),
    "spb_exchange_send_cells": (C.c_int64, [C.c_void_p, C.c_int]),
    "spb_exchange_recv_cells": (C.c_int64, [C.c_void_p, C.c_int]),
    "spb_exchange_local": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "spb_exchange_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "spb_exchange_unpack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "spb_exchange_pack_peer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "spb_dev_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "spb_dev_free": (C.c_int, [C.c_void_p]),
    "spb_ipc_export": (C.c_int, [C.c_void_p, C.c_char_p]),
    "spb_ipc_import": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "spb_ipc_close": (C.c_int, [C.c_void_p]),
    "spb_flag_signal": (C.c_int, [C.c_void_p, C.c_ulonglong, C.c_void_p]),
    "spb_flag_wait": (C.c_int, [C.c_void_p, C.c_ulonglong, C.c_void_p]),
    "spb_boundary_fill": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, _i64p, C.c_int64, C.POINTER(BcDesc), C.c_void_p]),
    "spb_source_term": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SourceDesc), C.c_void_p]),
    "spb_last_error": (C.c_char_p, []),
    "spb_device_count": (C.c_int, []),
    "spb_sync": (C.c_int, [C.c_void_p]),
    "spb_launch_count": (C.c_int64, []),
    "spb_version": (C.c_char_p, []),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SpbError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(spade_b200 has no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise SpbError(f"libspade_b200 error {rc}: {lib().spb_last_error().decode()}")


def int3(v):
    return (C.c_int * 3)(*[int(x) for x in v])
