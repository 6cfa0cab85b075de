"""spade_b200 — B200 (sm_100a) implementation of SPADE's RHS hot path behind the reference's operator API.

`spade_b200.api` mirrors the reference interface (grid_array, make_exchange, pde_algs.flux_div,
time_integration.integrator_t, algs.transform_reduce) on top of the C ABI in include/spade_b200.h;
all compute is in hand-written CUDA kernels in libspade_b200.so (spade_b200/csrc). No CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["api", "_lib"]
