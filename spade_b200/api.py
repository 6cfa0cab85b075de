"""Host-side mirror of the SPADE operator API for the RHS hot path, on top of the C ABI.

Names, argument meaning and error behaviour follow the reference (wvannoordt/spade, C++20 headers):
    cartesian_blocks_t / cartesian_grid_t   src/grid/cartesian_blocks.h:24-117, cartesian_grid.h:55-384
    grid_array                              src/grid/grid_array.h:182-422
    make_exchange / arr_exchange_t.exchange src/grid/make_exchange.h:111-421
    pde_algs.flux_div                       src/pde-algs/flux-div/flux_div.h:23-56
    time_integration.integrator_t, rk_t     src/time-integration/integrator.h:25-59, explicit.h:27-116,
                                            advance.h:236-280, 359-402
    algs.transform_reduce                   src/algs/transform_reduce.h:43-191
    flux functors                           src/navier-stokes/*.h
The C++ shim (include/spade_b200_shim.hpp) is the drop-in for an existing SPADE solver; this module is
the same thing for Python drivers, the tests and bench.py. torch supplies device memory, streams and
torch.distributed — every kernel is in libspade_b200.so.
"""
import ctypes as C
import os
from fractions import Fraction

import numpy as np
import torch

from . import _lib
from ._lib import BcDesc, FluxDesc, SourceDesc, SpbError, StageDesc, StagePlan, check, int3, lib

NVAR = 5

# ---- enums of include/spade_b200.h ---------------------------------------------------------------
CONV_NONE, CONV_TOTANI, CONV_CENT_KEEP4, CONV_FWENO, CONV_CENT_KEEP6, CONV_CENT_KEEP8 = 0, 1, 2, 3, 4, 5
DISS_NONE, DISS_FWENO = 0, 1
BLEND_FULL_FLUX, BLEND_DISS_FLUX = 0, 1
RED_MAX, RED_SUM = 0, 1
FN_WAVESPEED, FN_VAR, FN_ABSVAR, FN_KINETIC = 0, 1, 2, 3


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dptr(t):
    return C.c_void_p(t.data_ptr())


# ---- parallel group (reference parallel::pool_t, src/parallel/compute_pool.h:141-365) --------------
class pool_t:
    """rank/size view of the process group; one process per GPU (torch.distributed) instead of the
    reference's one std::thread per GPU."""

    def __init__(self, rank=0, size=1, group=None):
        self._rank, self._size, self.group = int(rank), int(size), group

    @staticmethod
    def from_torch(group=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return pool_t(dist.get_rank(group), dist.get_world_size(group), group)
        return pool_t()

    def rank(self):
        return self._rank

    def size(self):
        return self._size

    def isroot(self):
        return self._rank == 0

    def sync(self):
        if self._size > 1:
            import torch.distributed as dist
            dist.barrier(self.group)

    def reduce(self, value, op):
        """pool_t::reduce (compute_pool.h:247-284): all ranks get the reduced scalar."""
        if self._size == 1:
            return value
        import torch.distributed as dist
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        if op == RED_MAX:
            # a NaN on any rank must survive (a diverged field has to stop the CFL logic): reduce (value, is-NaN) together
            bad = value != value
            t = torch.tensor([-np.inf if bad else value, 1.0 if bad else 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            return float("nan") if float(t[1].item()) > 0.0 else float(t[0].item())
        t = torch.tensor([value], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return float(t.item())


# ---- grid ---------------------------------------------------------------------------------------------
class cartesian_blocks_t:
    """Uniform lattice of blocks; global id lb = bi + nb0*(bj + nb1*bk) (cartesian_blocks.h:54-55,96)."""

    def __init__(self, num_blocks, bounds):
        self.num_blocks = [int(x) for x in num_blocks]
        self.bounds = [float(x) for x in bounds]            # xmin xmax ymin ymax zmin zmax
        assert len(self.num_blocks) == 3 and len(self.bounds) == 6
        self.total_blocks = self.num_blocks[0] * self.num_blocks[1] * self.num_blocks[2]

    def total_num_blocks(self):
        return self.total_blocks

    def block_index(self, lb):
        nb = self.num_blocks
        return (lb % nb[0], (lb // nb[0]) % nb[1], lb // (nb[0] * nb[1]))

    def get_block_box(self, lb):
        """cartesian_blocks.h:49,66-71: min = bounds.min + b*bsize; max = min + bsize."""
        b = self.block_index(lb)
        out = []
        for d in range(3):
            bsize = (self.bounds[2 * d + 1] - self.bounds[2 * d]) / self.num_blocks[d]
            lo = self.bounds[2 * d] + b[d] * bsize
            out += [lo, lo + bsize]
        return out


class identity:
    """coords::identity (core/coord_system.h:51): the only coordinate system the reference's
    gradient-based fluxes accept (info_gradient.h:83)."""


class identity_1D:
    """coords::identity_1D (core/coord_system.h:54-63)."""

    def map(self, x):
        return np.asarray(x, dtype=np.float64)

    def coord_deriv(self, x):
        return np.ones_like(np.asarray(x, dtype=np.float64))


class scaled_coord_1D:
    """coords::scaled_coord_1D(k) (core/coord_system.h:142-158): x = k xi."""

    def __init__(self, k):
        self.k = float(k)

    def map(self, x):
        return self.k * np.asarray(x, dtype=np.float64)

    def coord_deriv(self, x):
        return np.full_like(np.asarray(x, dtype=np.float64), self.k)


class quad_1D:
    """coords::quad_1D (core/coord_system.h:160-173): x = xi^2."""

    def map(self, x):
        x = np.asarray(x, dtype=np.float64)
        return x * x

    def coord_deriv(self, x):
        return 2.0 * np.asarray(x, dtype=np.float64)


class integrated_tanh_1D:
    """coords::integrated_tanh_1D(y0, y1, inflation, rate) (core/coord_system.h:103-140): wall-clustered stretching of a
    channel; the constructor's assignments (eta1 = y0, eta0 = y1) are kept as written there."""

    def __init__(self, y0, y1, inflation, rate):
        self.eta1, self.eta0 = float(y0), float(y1)
        self.deta = self.eta1 - self.eta0
        k, strch = float(rate), float(inflation)
        self.alpha0, self.alpha1 = k / self.deta, -k / self.deta
        self.beta0 = self.alpha0 * (strch * self.deta - self.eta0)
        self.beta1 = self.alpha0 * (self.eta1 + strch * self.deta)
        self.f0 = float(self._func(self.eta0))
        self.normInv = 1.0 / (float(self._func(self.eta1)) - self.f0)

    def _func(self, eta):
        eta = np.asarray(eta, dtype=np.float64)
        return (np.log(np.abs(np.cosh(self.alpha0 * eta + self.beta0))) / self.alpha0
                + np.log(np.abs(np.cosh(self.alpha1 * eta + self.beta1))) / self.alpha1 - eta)

    def map(self, x):
        return self.eta0 + self.deta * (self._func(x) - self.f0) * self.normInv

    def coord_deriv(self, x):
        x = np.asarray(x, dtype=np.float64)
        return (self.deta * (np.tanh(self.alpha0 * x + self.beta0) + np.tanh(self.alpha1 * x + self.beta1) - 1.0)) * self.normInv


class diagonal_coords:
    """coords::diagonal_coords(xcoord, ycoord, zcoord) (core/coord_system.h:65-90): one 1-D mapping per direction.
    `metric_at` selects where info::metric evaluates coord_deriv: "physical" = the mapped position, which is what the
    reference does (omni/infos/info_metric.h:31 passes grid.get_coords(idx)); "computational" = the consistent choice.
    calc_jacobian always uses the computational position (flux_div_basic.h:49-50)."""

    def __init__(self, xcoord=None, ycoord=None, zcoord=None, metric_at="physical"):
        self.maps = [m if m is not None else identity_1D() for m in (xcoord, ycoord, zcoord)]
        if metric_at not in ("physical", "computational"):
            raise SpbError("diagonal_coords: metric_at is 'physical' (reference) or 'computational'")
        self.metric_at = metric_at

    def map(self, x):
        return [self.maps[d].map(x[d]) for d in range(3)]


class cartesian_grid_t:
    """cartesian_grid_t(cells_in_block, blocks, coords, group), cartesian_grid.h:84-89. coords: identity (default) or
    diagonal_coords; dense coordinate systems (coords::cyl_coords) are not implemented."""

    def __init__(self, cells_in_block, blocks, coords=None, group=None):
        if coords is not None and not isinstance(coords, (identity, diagonal_coords)):
            raise SpbError("cartesian_grid_t: coords must be identity or diagonal_coords")
        self.coords = coords if isinstance(coords, diagonal_coords) else None
        self.num_cell = [int(x) for x in cells_in_block]
        self.blocks = blocks
        self._group = group if group is not None else pool_t()
        # partition::block_partition_t (partition.h:27-84) — taken from the library's own plan builder
        plan = C.c_void_p()
        check(lib().spb_exchange_create(C.byref(plan), int3(blocks.num_blocks), int3(self.num_cell), int3([1, 1, 1]),
                                        int3([1, 1, 1]), self._group.rank(), self._group.size()))
        self.num_local_blocks = int(lib().spb_exchange_local_blocks(plan))
        self.first_block = int(lib().spb_exchange_first_block(plan))
        lib().spb_exchange_destroy(plan)
        self._bbox = np.array([blocks.get_block_box(self.first_block + l) for l in range(self.num_local_blocks)],
                              dtype=np.float64).reshape(-1)
        self._handles = {}
        self._bnd = {}

    @staticmethod
    def from_boxes(cells_in_block, boxes, group=None, first_block=0):
        """A grid given by the bounding boxes of this rank's blocks ([nlb][6], local order): what SPADE's
        cartesian_grid_t on amr::amr_blocks_t hands over after refine_blocks (cartesian_grid.h:331-368; every block keeps
        the same cell count, only its box and so its dx differ). The refinement logic itself stays in SPADE."""
        g = cartesian_grid_t.__new__(cartesian_grid_t)
        g.num_cell = [int(x) for x in cells_in_block]
        g.blocks = None
        g._group = group if group is not None else pool_t()
        g.coords = None
        g._boxes = np.ascontiguousarray(boxes, dtype=np.float64).reshape(-1, 6)
        g.num_local_blocks = g._boxes.shape[0]
        g.first_block = int(first_block)
        g._bbox = g._boxes.reshape(-1).copy()
        g._handles = {}
        g._bnd = {}
        return g

    def handle(self, num_exch):
        """spb_grid for arrays with `num_exch` exchange cells (device image of grid_geometry_t)."""
        key = tuple(int(x) for x in num_exch)
        if key not in self._handles:
            h = C.c_void_p()
            check(lib().spb_grid_create(C.byref(h), int3(self.num_cell), int3(key), self.num_local_blocks,
                                        self._bbox.ctypes.data_as(C.POINTER(C.c_double))))
            self._handles[key] = h
            if self.coords is not None:
                self._set_metric(h, key)
        return self._handles[key]

    def metric_tables(self, num_exch):
        """The three 1-D tables per direction of spb_grid_set_metric for this rank's blocks: (area, jac, face), each a
        list over directions of arrays [nlb, n_d + 2 g_d (+ 1)]; positions as grid_geometry.h:57-70."""
        area, jac, face = [], [], []
        box = self._bbox.reshape(-1, 6)
        for d in range(3):
            m = self.coords.maps[d]
            n, g = self.num_cell[d], int(num_exch[d])
            dx = (box[:, 2 * d + 1] - box[:, 2 * d]) / n
            xc = box[:, 2 * d, None] + (np.arange(-g, n + g)[None, :] + 0.5) * dx[:, None]
            xf = box[:, 2 * d, None] + (np.arange(-g, n + g + 1)[None, :] + 0.5 - 0.5) * dx[:, None]
            jac.append(np.ascontiguousarray(m.coord_deriv(xc)))
            face.append(np.ascontiguousarray(m.coord_deriv(xf)))
            area.append(np.ascontiguousarray(m.coord_deriv(m.map(xc)) if self.coords.metric_at == "physical" else jac[-1].copy()))
        return area, jac, face

    def _set_metric(self, h, num_exch):
        area, jac, face = self.metric_tables(num_exch)
        md = _lib.MetricDesc()
        dp = C.POINTER(C.c_double)
        for d in range(3):
            md.area[d], md.jac[d], md.face[d] = area[d].ctypes.data_as(dp), jac[d].ctypes.data_as(dp), face[d].ctypes.data_as(dp)
        check(lib().spb_grid_set_metric(h, C.byref(md)))

    def __del__(self):
        try:
            for h in self._handles.values():
                lib().spb_grid_destroy(h)
            self._handles = {}
        except Exception:
            pass

    def group(self):
        return self._group

    def boundary_blocks(self, ibndy):
        """grid_geometry_t::boundary_blocks[ibndy] (cartesian_grid.h:139-146): local blocks on face ibndy of the lattice."""
        if ibndy not in self._bnd:
            idir, pm = ibndy // 2, ibndy % 2
            want = self.blocks.num_blocks[idir] - 1 if pm else 0
            self._bnd[ibndy] = np.array([l for l in range(self.num_local_blocks)
                                         if self.blocks.block_index(self.first_block + l)[idir] == want], dtype=np.int64)
        return self._bnd[ibndy]

    def get_num_local_blocks(self):
        return self.num_local_blocks

    def get_num_global_blocks(self):
        return self.blocks.total_blocks

    def get_num_cells(self, d=None):
        return self.num_cell if d is None else self.num_cell[d]

    def get_dx(self, d, lb=0):
        box = self.blocks.get_block_box(self.first_block + lb) if self.blocks is not None else self._boxes[lb]
        return (box[2 * d + 1] - box[2 * d]) / self.num_cell[d]

    def get_grid_size(self):
        return self.num_cell[0] * self.num_cell[1] * self.num_cell[2] * self.blocks.total_blocks

    def local_cells(self):
        return self.num_cell[0] * self.num_cell[1] * self.num_cell[2] * self.num_local_blocks

    def cell_centers(self, lb, num_exch):
        """x,y,z of all padded cells of local block lb (grid_geometry.h:58-70), numpy arrays."""
        box = self.blocks.get_block_box(self.first_block + lb)
        out = []
        for d in range(3):
            dx = (box[2 * d + 1] - box[2 * d]) / self.num_cell[d]
            idx = np.arange(-num_exch[d], self.num_cell[d] + num_exch[d])
            out.append(box[2 * d] + (idx + 0.5) * dx)
        return out


class grid_array:
    """grid_array(grid, fill, num_exch, device) — 5-variable cell-centred array on the device in the
    reference's memory order off = v + 5*(i' + ni*(j' + nj*(k' + nk*lb))) (mem_map.h:484-496): a contiguous
    float64 cuda tensor of shape [nlb, nk', nj', ni', 5]."""

    def __init__(self, grid, fill=0.0, num_exch=(2, 2, 2), data=None):
        self.grid = grid
        self.num_exch = [int(x) for x in num_exch]
        self.h = grid.handle(self.num_exch)
        p = [n + 2 * g for n, g in zip(grid.num_cell, self.num_exch)]
        self.shape = (grid.num_local_blocks, p[2], p[1], p[0], NVAR)
        if data is None:
            self.data = torch.full(self.shape, float(fill), dtype=torch.float64, device="cuda")
        else:
            assert tuple(data.shape) == self.shape and data.dtype == torch.float64 and data.is_cuda
            self.data = data.contiguous()

    @staticmethod
    def from_host(grid, host, num_exch=(2, 2, 2)):
        """host: numpy array in reference order; copied through pinned memory."""
        p = [n + 2 * g for n, g in zip(grid.num_cell, num_exch)]
        shape = (grid.num_local_blocks, p[2], p[1], p[0], NVAR)
        h = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float64).reshape(shape))
        return grid_array(grid, num_exch=num_exch, data=h.pin_memory().to("cuda", non_blocking=True))

    def to_host(self):
        return self.data.cpu().numpy()

    def get_grid(self):
        return self.grid

    def get_num_exchange(self):
        return self.num_exch

    def interior(self):
        g, n = self.num_exch, self.grid.num_cell
        return self.data[:, g[2]:g[2] + n[2], g[1]:g[1] + n[1], g[0]:g[0] + n[0], :]

    def clone(self):
        return grid_array(self.grid, num_exch=self.num_exch, data=self.data.clone())


# ---- flux functors (closed set; each mirrors a reference type) ---------------------------------------------
class ideal_gas_t:
    """fluid_state::ideal_gas_t(gamma, R), gas.h:35-49"""

    def __init__(self, gamma, R):
        self.gamma, self.R = float(gamma), float(R)


class constant_viscosity_t:
    """viscous_laws::constant_viscosity_t(visc, prandtl), viscous_laws.h:62-100"""

    def __init__(self, visc, prandtl):
        self.visc = float(visc)
        self.beta = -2.0 * self.visc / 3.0
        self.prandtl_inv = 1.0 / float(prandtl)


class _functor:
    def desc(self):
        raise NotImplementedError


class totani_lr(_functor):
    """convective::totani_lr, convective.h:54-94"""

    def __init__(self, gas):
        self.gas = gas

    def _fill(self, d):
        d.conv = CONV_TOTANI
        d.gamma, d.R = self.gas.gamma, self.gas.R


class cent_keep(_functor):
    """convective::cent_keep<order>(gas), convective.h:97-192; order 2 == totani_lr arithmetic; orders 6 and 8 need 3 and 4
    exchange cells."""

    def __init__(self, order, gas):
        if order not in (2, 4, 6, 8):
            raise SpbError("cent_keep: orders 2, 4, 6, 8 (convective.h:100)")
        self.order, self.gas = order, gas

    def _fill(self, d):
        d.conv = {2: CONV_TOTANI, 4: CONV_CENT_KEEP4, 6: CONV_CENT_KEEP6, 8: CONV_CENT_KEEP8}[self.order]
        d.gamma, d.R = self.gas.gamma, self.gas.R


enable_smooth, disable_smooth = "enable_smooth", "disable_smooth"      # convective::weno_smooth_indicator, convective.h:248-252


class fweno_t(_functor):
    """convective::fweno_t<gas, use_smooth>, convective.h:336-497; disable_smooth = the linear weights 1/3, 2/3, 2/3, 1/3
    (convective.h:397-401, "use this for MMS")"""

    def __init__(self, gas, use_smooth=enable_smooth):
        if use_smooth not in (enable_smooth, disable_smooth):
            raise SpbError("fweno_t: use_smooth is enable_smooth or disable_smooth")
        self.gas, self.use_smooth = gas, use_smooth

    def _fill(self, d):
        d.conv = CONV_FWENO
        d.gamma, d.R = self.gas.gamma, self.gas.R
        d.weno_linear = 1 if self.use_smooth == disable_smooth else 0


class rusanov_t:
    """convective::rusanov_t(gas), flux_funcs.h:9-53: the local Lax-Friedrichs split fluxes consumed by weno_t."""

    def __init__(self, gas):
        self.gas = gas


class weno_t(fweno_t):
    """convective::weno_t(rusanov_t(gas)) with enable_smooth, convective.h:256-333: the same reconstruction as fweno_t
    (the nonlinear weights a0/(a0+a1) and be1^2/(2 be0^2 + be1^2) are the same number), applied to precomputed split
    fluxes; it runs on the fweno_t kernel and agrees with the reference's weno_t to round-off."""

    def __init__(self, flux_func, use_smooth=enable_smooth):
        if not isinstance(flux_func, rusanov_t):
            raise SpbError("weno_t: implemented for the rusanov_t flux function")
        super().__init__(flux_func.gas, use_smooth)


class ducros_t:
    """state_sensor::ducros_t(epsilon), state_sensor.h:21-43"""

    def __init__(self, epsilon):
        self.epsilon = float(epsilon)


full_flux, diss_flux = BLEND_FULL_FLUX, BLEND_DISS_FLUX


class hybrid_scheme_t(_functor):
    """convective::hybrid_scheme_t(scheme0, scheme1, blender, tag), hybrid_scheme.h:15-47"""

    def __init__(self, scheme0, scheme1, blender, tag=full_flux):
        if (not isinstance(scheme1, fweno_t) or not isinstance(blender, ducros_t) or isinstance(scheme0, fweno_t)
                or (isinstance(scheme0, cent_keep) and scheme0.order > 4)):
            raise SpbError("hybrid_scheme_t: implemented for (totani_lr|cent_keep, fweno_t, ducros_t)")
        self.scheme0, self.scheme1, self.blender, self.tag = scheme0, scheme1, blender, tag

    def _fill(self, d):
        self.scheme0._fill(d)
        d.diss = DISS_FWENO
        d.weno_linear = 1 if self.scheme1.use_smooth == disable_smooth else 0
        d.blend = self.tag
        d.sensor_eps = self.blender.epsilon


class wale_t:
    """subgrid_scale::wale_t(gas, cw, delta, prt), subgrid_scale.h:25-91 (Nicoud & Ducros eddy viscosity)."""

    def __init__(self, gas, cw, delta, prt):
        self.gas, self.cw, self.delta, self.prt = gas, float(cw), float(delta), float(prt)


class sgs_visc_t:
    """viscous_laws::sgs_visc_t(laminar, turb), viscous_laws.h:175-216: mu + mu_t, beta - 0.66666666667 mu_t, alpha + mu_t/Pr_t."""

    def __init__(self, lam, turb):
        if not isinstance(lam, constant_viscosity_t) or not isinstance(turb, wale_t):
            raise SpbError("sgs_visc_t: implemented for (constant_viscosity_t, wale_t)")
        self.lam, self.turb = lam, turb


class visc_lr(_functor):
    """viscous::visc_lr(vlaw, gas), viscous.h:14-113; vlaw = constant_viscosity_t or sgs_visc_t(constant_viscosity_t, wale_t)"""

    def __init__(self, vlaw, gas):
        if not isinstance(vlaw, (constant_viscosity_t, sgs_visc_t)):
            raise SpbError("visc_lr: constant_viscosity_t or sgs_visc_t (power_law_t has no get_all() in the reference, viscous_laws.h:102-135)")
        self.vlaw, self.gas = vlaw, gas

    def _fill(self, d):
        lam = self.vlaw.lam if isinstance(self.vlaw, sgs_visc_t) else self.vlaw
        d.visc = 1
        d.mu, d.beta, d.prandtl_inv = lam.visc, lam.beta, lam.prandtl_inv
        d.gamma, d.R = self.gas.gamma, self.gas.R
        if isinstance(self.vlaw, sgs_visc_t):
            t = self.vlaw.turb
            d.sgs, d.sgs_cw, d.sgs_delta, d.sgs_prt = 1, t.cw, t.delta, t.prt


class composite_kernel_t(_functor):
    """omni::compose(k0, k1, ...), omni/compose.h:11-45: the sum of the kernels on the union stencil."""

    def __init__(self, *kernels):
        self.kernels = kernels
        nconv = sum(1 for k in kernels if not isinstance(k, visc_lr))
        nvisc = sum(1 for k in kernels if isinstance(k, visc_lr))
        if nconv > 1 or nvisc > 1:
            raise SpbError("compose: at most one convective and one viscous functor")

    def _fill(self, d):
        for k in self.kernels:
            k._fill(d)


def compose(*kernels):
    return composite_kernel_t(*kernels)


def flux_desc(flux_func):
    d = FluxDesc()
    d.conv, d.diss, d.blend, d.visc = CONV_NONE, DISS_NONE, BLEND_FULL_FLUX, 0
    d.gamma, d.R, d.mu, d.beta, d.prandtl_inv, d.sensor_eps = 1.4, 287.15, 0.0, 0.0, 1.0, 0.0
    d.sgs, d.sgs_cw, d.sgs_delta, d.sgs_prt = 0, 0.0, 0.0, 1.0
    flux_func._fill(d)
    return d


# ---- pde_algs ------------------------------------------------------------------------------------------------
overwrite, increment = 0, 1


def flux_div(prims, rhs, flux_func, traits=increment, blocks=None):
    """pde_algs::flux_div(prims, rhs, flux_func, traits); default trait is `increment` like
    flux_div_basic.h:32-35. `blocks=(b0,b1)` restricts to a local block range."""
    d = flux_func if isinstance(flux_func, FluxDesc) else flux_desc(flux_func)
    if blocks is None:
        check(lib().spb_flux_div(prims.h, _dptr(prims.data), _dptr(rhs.data), C.byref(d), int(traits), _stream_ptr()))
    else:
        check(lib().spb_flux_div_blocks(prims.h, _dptr(prims.data), _dptr(rhs.data), C.byref(d), int(traits),
                                        int(blocks[0]), int(blocks[1]), _stream_ptr()))


class flux_div_rhs_t:
    """The usual rhs callback of a SPADE solver, `[&](auto& rhs, const auto& q, const auto& t) { flux_div(q, rhs, f, traits); }`
    (development/cuda-tgv/main.cc:174-186), as an object the integrator can recognise: integrator_t then runs the
    fused flux_div + stage-update kernel (spb_flux_div_rk_stage) instead of two passes over q."""

    def __init__(self, flux_func, traits=overwrite):
        self.flux = flux_func if isinstance(flux_func, FluxDesc) else flux_desc(flux_func)
        self.traits = traits

    def __call__(self, rhs, q, t):
        flux_div(q, rhs, self.flux, self.traits)


# ---- exchange ------------------------------------------------------------------------------------------------
class arr_exchange_t:
    """make_exchange(array, periodic) -> handle; handle.exchange(array, pool) (make_exchange.h:111-421)."""

    def __init__(self, grid, num_exch, periodic, tables=None):
        self.grid = grid
        self.pool = grid.group()
        self._h = C.c_void_p()
        i64 = C.POINTER(C.c_int64)
        if tables is None:
            check(lib().spb_exchange_create(C.byref(self._h), int3(grid.blocks.num_blocks), int3(grid.num_cell),
                                            int3(num_exch), int3([int(bool(p)) for p in periodic]),
                                            self.pool.rank(), self.pool.size()))
        else:
            # transaction tables marshalled from SPADE's exchange_config_t (the C++ shim does the same): injection lists,
            # and for AMR grids the interpolation lists
            send, recv = (np.ascontiguousarray(t, dtype=np.int64).reshape(-1, 16) for t in tables[:2])
            check(lib().spb_exchange_create_from_tables(C.byref(self._h), int3(grid.num_cell), int3(num_exch), self.pool.rank(),
                                                        self.pool.size(), send.ctypes.data_as(i64), len(send),
                                                        recv.ctypes.data_as(i64), len(recv)))
            if len(tables) > 2:
                isend, irecv = (np.ascontiguousarray(t, dtype=np.int64).reshape(-1, 26) for t in tables[2:4])
                check(lib().spb_exchange_add_interp(self._h, isend.ctypes.data_as(i64), len(isend), irecv.ctypes.data_as(i64), len(irecv)))
        self.send_cells = [int(lib().spb_exchange_send_cells(self._h, p)) for p in range(self.pool.size())]
        self.recv_cells = [int(lib().spb_exchange_recv_cells(self._h, p)) for p in range(self.pool.size())]
        self._sendbuf, self._recvbuf = {}, {}
        self._runs, self._reqs = None, []
        self._p2p = None                # decided at the first exchange (collective): peer-memory path or NCCL send/recv

    def __del__(self):
        try:
            if getattr(self, "_p2p", None):
                for bufs, flags in self._remote.values():
                    for b in list(bufs) + [flags]:
                        lib().spb_ipc_close(b)
                for bufs, flags in self._own.values():
                    for b in list(bufs) + [flags]:
                        lib().spb_dev_free(b)
                self._p2p = False
            if getattr(self, "_h", None):
                lib().spb_exchange_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def tables(self):
        ns, nr = int(lib().spb_exchange_num_send(self._h)), int(lib().spb_exchange_num_recv(self._h))
        send = np.zeros((ns, 16), dtype=np.int64)
        recv = np.zeros((nr, 16), dtype=np.int64)
        offs = np.zeros((self.pool.size(), 6), dtype=np.int64)
        i64 = C.POINTER(C.c_int64)
        check(lib().spb_exchange_tables(self._h, send.ctypes.data_as(i64), recv.ctypes.data_as(i64), offs.ctypes.data_as(i64)))
        return send, recv, offs

    def num_interp(self):
        return int(lib().spb_exchange_num_interp_send(self._h)) + int(lib().spb_exchange_num_interp_recv(self._h))

    def local_injection_cells(self):
        """cells of the same-rank injection transactions (the ghost cells the fused stage kernel writes itself)"""
        if getattr(self, "_linj", None) is None:
            send, _, _ = self.tables()
            mine = send[send[:, 2] == self.pool.rank()]
            self._linj = int((mine[:, 9] * mine[:, 10] * mine[:, 11]).sum())
        return self._linj

    def _buffers(self):
        me = self.pool.rank()
        for p in range(self.pool.size()):
            if p == me:
                continue
            if self.send_cells[p] and p not in self._sendbuf:
                self._sendbuf[p] = torch.empty(NVAR * self.send_cells[p], dtype=torch.float64, device="cuda")
            if self.recv_cells[p] and p not in self._recvbuf:
                self._recvbuf[p] = torch.empty(NVAR * self.recv_cells[p], dtype=torch.float64, device="cuda")

    def sendrecv(self, sendbufs, recvbufs):
        """exchange_message_t::send_all (exchange_message.h:14-56) over torch.distributed: one message per peer in each
        direction, posted as one batch (NCCL groups them; gloo runs them as they come). Returns the requests."""
        import torch.distributed as dist
        ops = []
        for p in sorted(recvbufs):
            ops.append(dist.P2POp(dist.irecv, recvbufs[p], p, group=self.pool.group))
        for p in sorted(sendbufs):
            ops.append(dist.P2POp(dist.isend, sendbufs[p], p, group=self.pool.group))
        return dist.batch_isend_irecv(ops) if ops else []

    def boundary_block_runs(self):
        """Local block ranges [b0, b1) that own a cell another rank needs — the source blocks of the off-rank send
        transactions, injection AND interpolation (AMR donors), from spb_exchange_boundary_blocks — as sorted disjoint
        runs, and the complementary runs: the stage kernel runs on the first set, its messages leave, and the second set is
        computed while they are in flight."""
        if self._runs is None:
            nlb = self.grid.num_local_blocks
            mark = np.zeros(max(nlb, 1), dtype=np.uint8)
            check(lib().spb_exchange_boundary_blocks(self._h, nlb, mark.ctypes.data_as(C.POINTER(C.c_ubyte))))
            mark = mark[:nlb].astype(bool)

            def runs(flag):
                idx = np.flatnonzero(mark == flag)
                if len(idx) == 0:
                    return []
                cuts = np.flatnonzero(np.diff(idx) > 1)
                starts = np.concatenate(([idx[0]], idx[cuts + 1]))
                ends = np.concatenate((idx[cuts], [idx[-1]])) + 1
                return [(int(a), int(b)) for a, b in zip(starts, ends)]
            self._runs = (runs(True), runs(False))
        return self._runs

    # ---- peer-memory path (one process per GPU on one node): no communication kernel holds an SM -----------------
    def _setup_p2p(self):
        """Receive buffers (two, used alternately: a neighbour may already pack its next message while this rank still unpacks
        the previous one) and arrival flags in raw device memory, exported through CUDA IPC and mapped by the neighbours.
        Returns False (and the NCCL path stays) if the backend is not NCCL, SPB_P2P=0, or any rank fails to map a handle."""
        import torch.distributed as dist
        self._p2p = False
        if self.pool.size() == 1 or os.environ.get("SPB_P2P", "1") == "0" or dist.get_backend(self.pool.group) != "nccl":
            return False
        me, n = self.pool.rank(), self.pool.size()
        own, export, ok = {}, {}, True
        try:
            for p in range(n):
                if p == me or not self.recv_cells[p]:
                    continue
                ptrs = []
                for _ in range(2):
                    b = C.c_void_p()
                    check(lib().spb_dev_alloc(C.byref(b), NVAR * 8 * self.recv_cells[p]))
                    ptrs.append(b)
                f = C.c_void_p()
                check(lib().spb_dev_alloc(C.byref(f), 16))                       # two 8-byte flags, zero-filled
                own[p] = (ptrs, f)
                hs = []
                for b in ptrs + [f]:
                    h = C.create_string_buffer(64)
                    check(lib().spb_ipc_export(b, h))
                    hs.append(h.raw)
                export[p] = hs
        except SpbError:
            ok = False
        gathered = [None] * n
        dist.all_gather_object(gathered, (ok, export), group=self.pool.group)
        ok = all(g[0] for g in gathered)
        remote = {}
        if ok:
            try:
                for p in range(n):
                    if p == me or not self.send_cells[p]:
                        continue
                    hs = gathered[p][1][me]                                  # rank p's buffers for messages from me
                    mapped = []
                    for h in hs:
                        out = C.c_void_p()
                        check(lib().spb_ipc_import(h, C.byref(out)))
                        mapped.append(out)
                    remote[p] = (mapped[:2], mapped[2])
            except (SpbError, KeyError):
                ok = False
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.pool.group)
        if int(flag.item()) == 0:
            # some rank could not allocate, export or map: every rank takes the NCCL path; give back what was set up here
            for bufs, f in remote.values():
                for b in list(bufs) + [f]:
                    lib().spb_ipc_close(b)
            dist.barrier(group=self.pool.group)             # no peer still maps the buffers freed below
            for bufs, f in own.values():
                for b in list(bufs) + [f]:
                    lib().spb_dev_free(b)
            return False
        self._own, self._remote, self._seq, self._p2p = own, remote, 0, True
        return True

    def begin(self, array):
        """First half of exchange(): the off-rank messages leave. Peer-memory path: each message is packed STRAIGHT INTO the
        neighbour's receive buffer over NVLink and the neighbour's arrival flag is raised behind it (stream order). NCCL
        path: pack into a send buffer and post the sends/receives (NCCL runs them on its own stream, ordered after the
        packs). Kernels launched after this call overlap the transfers."""
        self._reqs = []
        if self.pool.size() > 1:
            st = _stream_ptr()
            if self._p2p is None:
                self._setup_p2p()
            if self._p2p:
                self._seq += 1
                par = self._seq & 1
                for p, (bufs, flags) in self._remote.items():
                    check(lib().spb_exchange_pack_peer(self._h, _dptr(array.data), p, bufs[par], st))
                    check(lib().spb_flag_signal(C.c_void_p(flags.value + 8 * par), self._seq, st))
                return
            self._buffers()
            for p, buf in self._sendbuf.items():
                check(lib().spb_exchange_pack(self._h, _dptr(array.data), p, _dptr(buf), st))
            self._reqs = self.sendrecv(self._sendbuf, self._recvbuf)

    def finish(self, array, local=True):
        """Second half: same-rank ghost copies, then wait for the messages and unpack them."""
        st = _stream_ptr()
        if local:
            check(lib().spb_exchange_local(self._h, _dptr(array.data), st))
        else:
            check(lib().spb_exchange_local_interp(self._h, _dptr(array.data), st))    # AMR: what the stage kernel left (no-op otherwise)
        if self.pool.size() > 1 and self._p2p:
            par = self._seq & 1
            trace = getattr(self, "finish_trace", None)        # diagnosis: CUDA events around every wait and unpack
            if trace is not None:
                evs = [torch.cuda.Event(enable_timing=True)]
                evs[0].record()
            for p, (bufs, flags) in self._own.items():
                check(lib().spb_flag_wait(C.c_void_p(flags.value + 8 * par), self._seq, st))
                if trace is not None:
                    evs.append(torch.cuda.Event(enable_timing=True)); evs[-1].record()
                check(lib().spb_exchange_unpack(self._h, _dptr(array.data), p, bufs[par], st))
                if trace is not None:
                    evs.append(torch.cuda.Event(enable_timing=True)); evs[-1].record()
            if trace is not None:
                trace.append(evs)
            return
        for r in self._reqs:
            r.wait()
        self._reqs = []
        if self.pool.size() > 1:
            for p, buf in self._recvbuf.items():
                check(lib().spb_exchange_unpack(self._h, _dptr(array.data), p, _dptr(buf), st))

    def exchange(self, array, pool=None):
        self.begin(array)
        self.finish(array)          # the same-rank copies overlap the NVLink transfers


class exchange_bc_t:
    """The usual boundary callback of a periodic SPADE solver, `[&](auto& q, const auto& t) { handle.exchange(q, pool); }`
    (development/cuda-tgv/main.cc:188-191), as an object the integrator can recognise: with a fused stage plan the
    rank-boundary blocks are advanced first, their ghost messages leave over NVLink, and the rank-interior blocks are
    advanced while the messages are in flight."""

    def __init__(self, handle, boundaries=None, kern=None):
        """boundaries / kern: the other half of a wall-bounded solver's callback, algs::boundary_fill(q, boundaries, kern)
        after the exchange (bc = exchange + boundary_fill)."""
        self.handle, self.boundaries, self.kern = handle, boundaries, kern

    def after(self, q):
        if self.boundaries is not None:
            boundary_fill(q, self.boundaries, self.kern)

    def __call__(self, q, t):
        self.handle.exchange(q)
        self.after(q)


def make_exchange(array, periodic, tables=None):
    """make_exchange(array, periodic); `tables` = (send, recv[, interp_send, interp_recv]) builds the plan from transaction
    tables produced by SPADE's own exchange_config_t (needed for AMR grids, whose topology stays in SPADE)."""
    return arr_exchange_t(array.grid, array.num_exch, periodic, tables)


# ---- domain boundaries (reference src/grid/boundary_fill.h) -------------------------------------------------------
class identifier_t:
    """boundary::identifier_t = bound_box_t<bool, 3>: which of xmin xmax ymin ymax zmin zmax; `a || b` is `a | b` here."""

    def __init__(self, *flags):
        self.flags = tuple(bool(f) for f in flags)
        assert len(self.flags) == 6

    def __or__(self, other):
        return identifier_t(*[a or b for a, b in zip(self.flags, other.flags)])

    def __call__(self, idir, pm):
        return self.flags[2 * idir + pm]


class boundary:
    xmin = identifier_t(1, 0, 0, 0, 0, 0)
    xmax = identifier_t(0, 1, 0, 0, 0, 0)
    ymin = identifier_t(0, 0, 1, 0, 0, 0)
    ymax = identifier_t(0, 0, 0, 1, 0, 0)
    zmin = identifier_t(0, 0, 0, 0, 1, 0)
    zmax = identifier_t(0, 0, 0, 0, 0, 1)

    class extrap_t:
        """boundary::extrapolate<order> (boundary_fill.h:17-27)"""

        def __init__(self, order):
            self.order = int(order)

        def desc(self):
            d = BcDesc()
            d.kind, d.order = 1, self.order
            return d

    @staticmethod
    def extrapolate(order):
        return boundary.extrap_t(order)


class mirror_kernel_t:
    """The boundary_fill kernels `[=](const prim_t& q_image, int idir) -> prim_t` that are linear per variable:
    ghost[v] = a[v]*image[v] + b[v]; with a_normal the velocity component along idir uses that factor instead."""

    def __init__(self, a, b=(0.0,) * 5, a_normal=None):
        self.a, self.b, self.a_normal = [float(x) for x in a], [float(x) for x in b], a_normal

    def desc(self):
        d = BcDesc()
        d.kind, d.order = 0, 0
        d.a[:] = self.a
        d.b[:] = self.b
        d.use_normal = 0 if self.a_normal is None else 1
        d.a_normal = 0.0 if self.a_normal is None else float(self.a_normal)
        return d


def noslip_isothermal_wall(t_wall):
    """ghost = (p, 2 T_wall - T, -u, -v, -w)"""
    return mirror_kernel_t((1, -1, -1, -1, -1), (0, 2.0 * t_wall, 0, 0, 0))


def noslip_adiabatic_wall():
    return mirror_kernel_t((1, 1, -1, -1, -1))


def symmetry_plane():
    return mirror_kernel_t((1, 1, 1, 1, 1), a_normal=-1.0)


def boundary_fill(arr, boundaries, kern):
    """algs::boundary_fill(arr, boundaries, kern), boundary_fill.h:32-133: the listed boundaries in the order
    xmin xmax ymin ymax zmin zmax, each over the local blocks on that face of the block lattice."""
    d = kern.desc()
    grid = arr.grid
    for ib in range(6):
        idir, pm = ib // 2, ib % 2
        if not boundaries(idir, pm):
            continue
        blocks = grid.boundary_blocks(ib)
        if len(blocks) == 0:
            continue
        check(lib().spb_boundary_fill(arr.h, _dptr(arr.data), idir, pm, blocks.ctypes.data_as(C.POINTER(C.c_int64)),
                                      len(blocks), C.byref(d), _stream_ptr()))


# ---- source term (reference src/pde-algs/source_term.h) ------------------------------------------------------------
class body_force_t:
    """source_term_func `[=](const prim_t& q) -> flux_t {0, f.u, fx, fy, fz}`: the forcing of a channel run"""

    def __init__(self, fx, fy=0.0, fz=0.0):
        self.f = (float(fx), float(fy), float(fz))

    def desc(self):
        d = SourceDesc()
        d.kind = 0
        d.f[:] = list(self.f) + [0.0, 0.0]
        return d


class constant_source_t:
    def __init__(self, values):
        self.values = [float(x) for x in values]

    def desc(self):
        d = SourceDesc()
        d.kind = 1
        d.f[:] = self.values
        return d


def source_term(q, rhs, source_term_func):
    """pde_algs::source_term(q, rhs, source_term_func): rhs += S(q)/jac on interior cells"""
    d = source_term_func.desc()
    check(lib().spb_source_term(q.h, _dptr(q.data), _dptr(rhs.data), C.byref(d), _stream_ptr()))


# ---- algs ----------------------------------------------------------------------------------------------------
def transform_reduce(array, fn=FN_WAVESPEED, op=RED_MAX, gas=None, ivar=0):
    """algs::transform_reduce(array, make_reduction(array, f, op)) incl. the cross-rank pool.reduce
    (transform_reduce.h:171-190)."""
    gas = gas or ideal_gas_t(1.4, 287.15)
    out = C.c_double(0.0)
    check(lib().spb_reduce(array.h, _dptr(array.data), int(op), int(fn), int(ivar), gas.gamma, gas.R,
                           C.byref(out), _stream_ptr()))
    return array.grid.group().reduce(out.value, op)


# ---- time integration ---------------------------------------------------------------------------------------------
class rk_t:
    """Butcher table as exact ratios, explicit.h:27-116."""

    def __init__(self, table, accum, dt, name, high_storage=False):
        self.table = [[Fraction(x) for x in row] for row in table]
        self.accum = [Fraction(x) for x in accum]
        self.dt = [Fraction(x) for x in dt]
        self.name = name
        self.high_storage = bool(high_storage)

    def var_size(self):
        """explicit.h:33: the high-storage variants keep a second copy of the solution (same arithmetic in the fused
        prim/cons path, advance.h:236-280, which never touches the second copy)."""
        return 2 if self.high_storage else 1

    def rhs_size(self):
        return len(self.table[0])

    def rows(self):
        return len(self.table)


F = Fraction
rk2_t = rk_t([[0, 0], [F(1, 2), 0]], [0, 1], [0, F(1, 2)], "rk2")
rk2hs_t = rk_t([[0, 0], [F(1, 2), 0]], [0, 1], [0, F(1, 2)], "rk2hs", high_storage=True)
rk4_t = rk_t([[0, 0, 0, 0], [F(1, 2), 0, 0, 0], [0, F(1, 2), 0, 0], [0, 0, 1, 0]],
             [F(1, 6), F(1, 3), F(1, 3), F(1, 6)], [0, F(1, 2), F(1, 2), 1], "rk4")
ssprk3_t = rk_t([[0, 0, 0], [1, 0, 0], [F(1, 4), F(1, 4), 0]], [F(1, 6), F(1, 6), F(2, 3)], [0, 1, F(1, 2)], "ssprk3")
ssprk3hs_t = rk_t([[0, 0, 0], [1, 0, 0], [F(1, 4), F(1, 4), 0]], [F(1, 6), F(1, 6), F(2, 3)], [0, 1, F(1, 2)], "ssprk3hs",
                  high_storage=True)
ssprk34_t = rk_t([[0, 0, 0, 0], [F(1, 2), 0, 0, 0], [F(1, 2), F(1, 2), 0, 0], [F(1, 6), F(1, 6), F(1, 6), 0]],
                 [F(1, 6), F(1, 6), F(1, 6), F(1, 2)], [0, F(1, 2), 1, F(1, 2)], "ssprk34")
rk38r_t = rk_t([[0, 0, 0, 0], [F(1, 3), 0, 0, 0], [F(-1, 3), 1, 0, 0], [1, -1, 1, 0]],
               [F(1, 8), F(3, 8), F(3, 8), F(1, 8)], [0, F(1, 3), F(2, 3), 1], "rk38r")


class tspecial_rk3_t:
    name = "ssprk3_opt"

    @staticmethod
    def rhs_size():
        return 2


ssprk3_opt = tspecial_rk3_t()


def _ratio_diff_value(a, b):
    """ratio_diff_t + coeff_value_t (advance.h:47-55): (d1*n0 - d0*n1)/(d0*d1) evaluated in double."""
    n0, d0, n1, d1 = a.numerator, a.denominator, b.numerator, b.denominator
    num, den = d1 * n0 - d0 * n1, d0 * d1
    return float(num) / float(den) if num != 0 else 0.0


class time_axis_t:
    def __init__(self, t0, dt):
        self.t, self.dt = float(t0), float(dt)

    def time(self):
        return self.t

    def timestep(self):
        return self.dt


class integrator_data_t:
    """integrator_data_t(q, rhs, scheme): one solution array and rhs_size() residual registers."""

    def __init__(self, q, rhs, scheme):
        nvar = scheme.var_size() if hasattr(scheme, "var_size") else 1
        self.solution_data = [q] + [q.clone() for _ in range(nvar - 1)]          # high-storage tables: a second copy (explicit.h:33)
        self.residual_data = [rhs] + [rhs.clone() for _ in range(scheme.rhs_size() - 1)]

    def solution(self, i=0):
        return self.solution_data[i]

    def residual(self, i=0):
        return self.residual_data[i]


class state_transform_t:
    """fluid_state::state_transform_t(cons_t(), gas): selects the fused prim<->cons update."""

    def __init__(self, gas):
        self.gas = gas


class identity_transform_t:
    """time_integration::identity_transform (integrator.h:8-14): the array itself is integrated — integrator_t then takes
    the generic integrate_advance path (advance.h:109-230)."""


identity_transform = identity_transform_t()


class integrator_t:
    """integrator_t(axis, scheme, data, rhs_calc, boundary_cond, trans).advance()
    — the fused prim/cons path of advance.h:236-280 and the ssprk3_opt path of advance.h:359-402.
    If rhs_calc is a flux_div_rhs_t (overwrite trait) and the functor set is supported, every stage is ONE kernel
    (flux_div + stage update, spb_flux_div_rk_stage) writing into a second solution buffer; `fused=False` forces the
    two-kernel path."""

    def __init__(self, axis, scheme, data, rhs_calc, boundary_cond, trans=identity_transform, fused=True, fuse_exchange=True):
        if not isinstance(trans, (state_transform_t, identity_transform_t)):
            raise SpbError("integrator_t: trans is a state_transform_t or identity_transform")
        if isinstance(trans, identity_transform_t):
            if not isinstance(scheme, rk_t):
                raise SpbError("integrator_t: ssprk3_opt needs a state_transform_t (advance.h:359)")
            fused = False
        self.axis, self.scheme, self.data = axis, scheme, data
        self.rhs_calc, self.boundary_cond, self.trans = rhs_calc, boundary_cond, trans
        self._plan = None
        self._scratch = None
        self._fuse_exchange = bool(fuse_exchange)
        self._two_streams = os.environ.get("SPB_TWO_STREAMS", "1") != "0"
        self._block_runs = os.environ.get("SPB_BLOCK_RUNS", "0") == "1"
        self._defer = os.environ.get("SPB_DEFER_UNPACK", "1") != "0"
        self._pending = self._ev_b = self._ev_i = None
        self._side = None
        self.stage_events = None        # bench.py: a list collects (start, stop, algorithmic bytes per cell) per stage kernel
        self.phase_events = [] if os.environ.get("SPB_PHASE_EVENTS") else None   # diagnosis: per-phase CUDA events of every stage
        self.join_events = []
        self._boundary_delay = int(float(os.environ.get("SPB_BOUNDARY_DELAY_US", "0")) * 1900)   # development A/B (SM cycles)
        if fused and isinstance(rhs_calc, flux_div_rhs_t) and rhs_calc.traits == overwrite and isinstance(scheme, rk_t):
            if lib().spb_flux_div_rk_stage_supported(C.byref(rhs_calc.flux)):
                self._plan = self._fused_plan(scheme)

    def solution(self):
        return self.data.solution(0)

    def time(self):
        return self.axis.t

    @staticmethod
    def _fused_plan(s):
        """Per stage: which residual registers are read, which one is written, and the coefficients (without dt), from
        spb_rk_fused_plan — the one planner both host sides share. diffs[i][j] is the coefficient difference
        (a_{i+1,j} - a_{i,j}; last row: b_j - a_{n-1,j}) exactly as advance.h:47-55,84-92 forms it. A final update that needs
        more than two earlier residuals gets their combination C prepared by the stage before it in register n-2
        (rk4: C = k0/6 + k1/3 - 2 k2/3). None if a stage would need more than two inputs."""
        n = s.rows()
        rows = s.table + [s.accum]
        diffs = (C.c_double * (n * n))(*[_ratio_diff_value(c, p) for i in range(n) for c, p in zip(rows[i + 1], rows[i])])
        raw = (StagePlan * n)()
        rc = lib().spb_rk_fused_plan(n, diffs, raw)
        if rc == _lib.SPB_ERR_UNSUPPORTED:
            return None
        check(rc)
        # registers are named by their index ("k", j); register n-2 may hold the combination C instead of k_{n-2}
        return [{"cq_self": st.cq_self, "in": [("k", st.inp[a]) for a in range(st.nin)], "cq": [st.cq[a] for a in range(st.nin)],
                 "co": [st.co[a] for a in range(st.nin)], "out": ("k", st.out) if st.out >= 0 else None, "co_self": st.co_self}
                for st in raw]

    def _advance_fused(self):
        ax, dt, d = self.axis, self.axis.dt, self.data
        f = self.rhs_calc.flux
        if self._scratch is None:
            self._scratch = d.solution(0).clone()
        cur, nxt = d.solution(0), self._scratch
        s = self.scheme
        for i, st in enumerate(self._plan):
            sd = StageDesc()
            sd.nin = len(st["in"])
            for a, (_, j) in enumerate(st["in"]):
                sd.inp[a] = d.residual(j).data.data_ptr()
                sd.cq[a] = st["cq"][a] * dt
                sd.co[a] = st["co"][a]
            sd.cq_self = st["cq_self"] * dt
            sd.out = d.residual(st["out"][1]).data.data_ptr() if st["out"] else None
            sd.co_self = st["co_self"]
            ex = self.boundary_cond.handle if isinstance(self.boundary_cond, exchange_bc_t) else None
            overlap = ex is not None and ex.pool.size() > 1

            def launch_runs(part):
                # development A/B (SPB_BLOCK_RUNS=1): the round-1 form, one launch per contiguous run of blocks
                first, second = ex.boundary_block_runs()
                for b0, b1 in (first if part == _lib.SPB_PART_BOUNDARY else second):
                    check(lib().spb_flux_div_rk_stage_exchange(cur.h, _dptr(cur.data), _dptr(nxt.data), C.byref(f), C.byref(sd),
                                                               ex._h if self._fuse_exchange else None, b0, b1, _stream_ptr()))

            def launch(part):
                if self._block_runs and part != _lib.SPB_PART_ALL:
                    return launch_runs(part)
                # one launch per part of the local blocks (spb_flux_div_rk_stage_part: boundary / interior / all), whatever their
                # order in memory; with a recognised exchange handle the kernel also fills the same-rank injection ghosts of q_out
                exh = ex._h if ex is not None else None
                if ex is not None and self._fuse_exchange:
                    rc = lib().spb_flux_div_rk_stage_part(cur.h, _dptr(cur.data), _dptr(nxt.data), C.byref(f), C.byref(sd), exh, 1, part, _stream_ptr())
                    if rc != _lib.SPB_ERR_UNSUPPORTED:
                        return check(rc)
                    self._fuse_exchange = False                  # plan not canonical: separate same-rank copy from now on
                check(lib().spb_flux_div_rk_stage_part(cur.h, _dptr(cur.data), _dptr(nxt.data), C.byref(f), C.byref(sd), exh, 0, part, _stream_ptr()))

            # Deferred unpack (periodic multi-rank runs without AMR interpolation or wall fills): the off-rank ghost cells of stage s
            # are only read by the rank-boundary blocks of stage s+1, so their wait + unpack moves to the side stream in front of
            # that boundary kernel, and the rank-interior kernel of stage s+1 starts on the main stream without waiting for any
            # message: a neighbour that runs late (eight power-capped GPUs never run at the same pace) or a slow link stalls the
            # short side-stream chain unpack -> boundary kernel -> pack, which has the whole interior kernel to catch up.
            defer = (overlap and self._two_streams and self._defer and self._fuse_exchange and not self._block_runs
                     and self.boundary_cond.boundaries is None and ex.num_interp() == 0)
            if self.stage_events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if defer:
                if self._side is None:
                    self._side = torch.cuda.Stream(priority=-1)
                main, side = torch.cuda.current_stream(), self._side
                if i == 0:
                    side.wait_stream(main)                      # everything enqueued before this step
                    self._ev_b = self._ev_i = None
                pe = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if self.phase_events is not None else None
                if pe:
                    pe[0].record(main)
                with torch.cuda.stream(side):
                    if self._pending is not None:
                        ex.finish(self._pending, local=False)   # messages of the previous stage -> off-rank ghost cells of `cur`
                        self._pending = None
                    if pe:
                        pe[4].record(side)                      # (deferred schedule: the unpack of the PREVIOUS stage's messages)
                    if self._ev_i is not None:
                        side.wait_event(self._ev_i)             # the previous interior kernel wrote same-rank ghosts of the boundary blocks
                    if self._boundary_delay:
                        torch.cuda._sleep(self._boundary_delay)     # development A/B: let the interior kernel fill the SMs first
                    launch(_lib.SPB_PART_BOUNDARY)
                    ev_b = torch.cuda.Event()
                    ev_b.record(side)
                    if pe:
                        pe[1].record(side)
                    if self._fuse_exchange:
                        ex.begin(nxt)
                    if pe:
                        pe[2].record(side)
                if self._ev_b is not None:
                    main.wait_event(self._ev_b)                 # the previous boundary kernel wrote same-rank ghosts of the interior blocks
                if self._fuse_exchange:
                    launch(_lib.SPB_PART_INTERIOR)
                    ev_i = torch.cuda.Event()
                    ev_i.record(main)
                    if pe:
                        pe[3].record(main)
                        self.phase_events.append(pe)
                    self._ev_b, self._ev_i, self._pending = ev_b, ev_i, nxt
                else:
                    # the plan refused the ghost fusion at the first launch: finish this stage the undeferred way
                    main.wait_stream(side)
                    launch(_lib.SPB_PART_INTERIOR)
                    ex.begin(nxt)
                    self._ev_b = self._ev_i = None
                    defer = False
            elif overlap and self._two_streams:
                # rank-boundary blocks on a high-priority side stream, their messages packed and posted from it; the rank-interior
                # blocks run on the main stream AT THE SAME TIME (both read q_in only and write disjoint cells), so the small
                # boundary launch leaves no partial wave behind and the messages fly under the interior kernel
                if self._side is None:
                    self._side = torch.cuda.Stream(priority=-1)
                main, side = torch.cuda.current_stream(), self._side
                pe = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if self.phase_events is not None else None
                if pe:
                    pe[0].record(main)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    launch(_lib.SPB_PART_BOUNDARY)
                    if pe:
                        pe[1].record(side)
                    ex.begin(nxt)
                    if pe:
                        pe[2].record(side)
                launch(_lib.SPB_PART_INTERIOR)
                if pe:
                    pe[3].record(main)
                main.wait_stream(side)
            elif overlap:
                launch(_lib.SPB_PART_BOUNDARY)
                ex.begin(nxt)
                launch(_lib.SPB_PART_INTERIOR)
            else:
                launch(_lib.SPB_PART_ALL)
            if self.stage_events is not None:
                e1.record()
                # algorithmic bytes per interior cell of this launch: q in, q out, residual registers read / written, and one
                # 40-byte write per same-rank injection ghost cell when the ghost exchange is fused into the kernel
                ghost = 0.0
                if ex is not None and self._fuse_exchange:
                    ghost = 40.0 * ex.local_injection_cells() / max(1, cur.grid.local_cells())
                self.stage_events.append((e0, e1, 80.0 + 40.0 * sd.nin + (40.0 if st["out"] else 0.0) + ghost))
            cur, nxt = nxt, cur
            tnext = ax.t + (float(s.dt[i + 1]) * dt if i + 1 < s.rows() else dt)
            if i + 1 == s.rows():
                ax.t += dt
                tnext = ax.t
            if defer:
                if i + 1 == s.rows():
                    # end of the step: the last messages are unpacked and the main stream sees the whole state
                    main, side = torch.cuda.current_stream(), self._side
                    with torch.cuda.stream(side):
                        ex.finish(self._pending, local=False)
                        self._pending = None
                    if self.phase_events is not None:
                        # diagnosis: the join of a step (last interior kernel done -> last messages unpacked on the side stream)
                        ej = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                        ej[0].record(main)
                        main.wait_stream(side)
                        ej[1].record(main)
                        self.join_events.append(ej)
                    else:
                        main.wait_stream(side)
            elif ex is not None:
                ex.finish(cur, local=not self._fuse_exchange)
                self.boundary_cond.after(cur)
                if self.phase_events is not None and overlap and self._two_streams:
                    pe[4].record(torch.cuda.current_stream())
                    self.phase_events.append(pe)
            else:
                self.boundary_cond(cur, tnext)
        if cur is not d.solution(0):                       # odd number of stages: the result sits in the scratch buffer
            d.solution(0).data, self._scratch.data = self._scratch.data, d.solution(0).data

    def _update(self, prev_row, curr_row):
        q = self.data.solution(0)
        dt = self.axis.dt
        nk = len(curr_row)
        coeff = (C.c_double * nk)(*[_ratio_diff_value(c, p) * dt for c, p in zip(curr_row, prev_row)])
        ks = (C.c_void_p * nk)(*[r.data.data_ptr() for r in self.data.residual_data[:nk]])
        check(lib().spb_rk_update(q.h, _dptr(q.data), ks, nk, coeff, self.trans.gas.gamma, self.trans.gas.R, _stream_ptr()))

    def _advance_generic(self):
        """integrate_advance for a trans that is not a state_transform_t (advance.h:109-230) with identity_transform: per
        stage the solution is augmented with dt a_ij k_j, the residual evaluated, and (without a second copy of the
        solution) the augmentation taken out again; each `resid *= c; sol += resid; resid *= 1/c` triple is one kernel."""
        ax, dt, d, s = self.axis, self.axis.dt, self.data, self.scheme
        second = len(d.solution_data) > 1

        def axpy(sol, resid, c, subtract):
            check(lib().spb_axpy_roundtrip(sol.h, _dptr(sol.data), _dptr(resid.data), float(c), int(subtract), _stream_ptr()))

        for i in range(s.rows()):
            sol = d.solution(1 if second else 0)
            if second:
                sol.data.copy_(d.solution(0).data)
            used = [j for j in range(i) if s.table[i][j] != 0]
            for j in used:
                axpy(sol, d.residual(j), dt * (s.table[i][j].numerator / s.table[i][j].denominator), 0)
            t = ax.t + float(s.dt[i]) * dt
            if i > 0:
                self.boundary_cond(sol, t)
            self.rhs_calc(d.residual(i), sol, t)
            if not second:
                for j in used:
                    axpy(sol, d.residual(j), dt * (s.table[i][j].numerator / s.table[i][j].denominator), 1)
        q = d.solution(0)
        for i in range(s.rows()):
            if s.accum[i] != 0:
                axpy(q, d.residual(i), (s.accum[i].numerator / s.accum[i].denominator) * dt, 0)
        ax.t += dt
        self.boundary_cond(q, ax.t)

    def advance(self):
        if isinstance(self.trans, identity_transform_t):
            return self._advance_generic()
        if self._plan is not None:
            return self._advance_fused()
        ax, q, dt = self.axis, self.data.solution(0), self.axis.dt
        gas = self.trans.gas
        if isinstance(self.scheme, tspecial_rk3_t):
            r0, r1 = self.data.residual(0), self.data.residual(1)
            tc = [0.0, 1.0, 0.5]
            self.rhs_calc(r0, q, ax.t + tc[0] * dt)
            for stage in range(3):
                check(lib().spb_ssprk3_stage(q.h, stage, _dptr(q.data), _dptr(r0.data), _dptr(r1.data), dt,
                                             gas.gamma, gas.R, _stream_ptr()))
                if stage < 2:
                    self.boundary_cond(q, ax.t + tc[stage + 1] * dt)
                    self.rhs_calc(r1, q, ax.t + tc[stage + 1] * dt)
            ax.t += dt
            self.boundary_cond(q, ax.t)
            return
        s = self.scheme
        self.rhs_calc(self.data.residual(0), q, ax.t + float(s.dt[0]) * dt)
        for i in range(1, s.rows()):
            self._update(s.table[i - 1], s.table[i])
            self.boundary_cond(q, ax.t + float(s.dt[i]) * dt)
            self.rhs_calc(self.data.residual(i), q, ax.t + float(s.dt[i]) * dt)
        self._update(s.table[s.rows() - 1], s.accum)
        ax.t += dt
        self.boundary_cond(q, ax.t)


# ---- checkpoint files in the reference's byte order (reference src/io/io_native.h:18-56) --------------------------
class io:
    """io::binary_write / io::binary_read: a headerless file in which global block lb_glob occupies the bytes
    [lb_glob * B, (lb_glob + 1) * B), B = bytes of one padded block (exchange cells included) in the array's own memory
    order. The device layout here IS that order, so a rank's share is one contiguous slab at offset first_block * B:
    checkpoints written by a SPADE solver restart here and the other way round."""

    @staticmethod
    def _slab(array):
        per_block = int(np.prod(array.shape[1:])) * 8
        return per_block, array.grid.first_block * per_block

    @staticmethod
    def binary_write(filename, array):
        per_block, off = io._slab(array)
        host = array.data.cpu().numpy()
        pool = array.grid.group()
        if pool.isroot():
            total = (array.grid.get_num_global_blocks() if array.grid.blocks is not None else array.shape[0]) * per_block
            with open(filename, "ab"):
                pass
            os.truncate(filename, total)
        pool.sync()
        with open(filename, "r+b") as f:
            f.seek(off)
            f.write(host.tobytes())
        pool.sync()

    @staticmethod
    def binary_read(filename, array):
        per_block, off = io._slab(array)
        nbytes = per_block * array.shape[0]
        with open(filename, "rb") as f:
            f.seek(off)
            raw = f.read(nbytes)
        if len(raw) != nbytes:
            raise SpbError(f"binary_read: {filename} holds {len(raw)} of the {nbytes} bytes of this rank's blocks")
        host = torch.from_numpy(np.frombuffer(raw, dtype=np.float64).reshape(array.shape).copy())
        array.data.copy_(host.pin_memory(), non_blocking=True)
        torch.cuda.current_stream().synchronize()


    # ---- visualisation files (reference src/io/io_vtk.h:276-288) -----------------------------------------------------
    VAR_NAMES = {"prim": ("P", "T", "U", "V", "W"), "cons": ("rho", "rhoH", "rhoU", "rhoV", "rhoW"),
                 "flux": ("continuity", "energy", "x_momentum", "y_momentum", "z_momentum")}       # fluid_state.h:28-77

    @staticmethod
    def _b64(raw):
        """spade::detail::stream_base_64 of a vector (core/base_64.h:68-75): the uint32 byte count and the payload are
        encoded separately, each with its own '=' padding."""
        import base64
        import struct
        return base64.b64encode(struct.pack("<I", len(raw))) + base64.b64encode(raw)

    @staticmethod
    def write_vtk_files(out_dir, basename, host, num_cell, num_exch, boxes, global_ids, num_global_blocks, coords=None,
                        names=("P", "T", "U", "V", "W"), write_base=True):
        """The files of io::output_vtk for blocks held on the host: `<basename>.pvts` (written if write_base: the root rank's
        job) and one `data_<basename>/b<9-digit global id>.vts` per block — StructuredGrid pieces with the interior cells
        of every variable as base-64 "binary" CellData in i-fastest order and the (ni+1)(nj+1)(nk+1) node positions
        coords.map(box.min + index*dx) (grid_geometry.h:52-70), byte for byte the reference's files
        (tests/golden/vtk_small, written by the unmodified reference)."""
        ni, nj, nk = (int(x) for x in num_cell)
        g = [int(x) for x in num_exch]
        os.makedirs(os.path.join(out_dir, "data_" + basename), exist_ok=True)
        if write_base:
            with open(os.path.join(out_dir, basename + ".pvts"), "w") as f:
                f.write('<?xml version="1.0"?>\n<VTKFile type="PStructuredGrid" version="0.1" byte_order="LittleEndian">\n')
                f.write(f'<PStructuredGrid WholeExtent="0 {ni} 0 {nj} 0 {nk * num_global_blocks}" GhostLevel="0">\n')
                f.write(f'<PCellData Scalars = "{",".join(names)}">\n')
                for nm in names:
                    f.write(f'<PDataArray type="Float64" Name="{nm}"/>\n')
                f.write('</PCellData>\n<PPoints>\n<PDataArray type="Float64" NumberOfComponents="3"/>\n</PPoints>\n')
                for lb in range(num_global_blocks):
                    f.write(f'<Piece Extent="0 {ni} 0 {nj} {lb * nk} {(lb + 1) * nk}" Source="data_{basename}/b{lb:09d}.vts"/>\n')
                f.write('</PStructuredGrid>\n</VTKFile>\n')
        for l, lb_glob in enumerate(global_ids):
            blk = host[l, g[2]:g[2] + nk, g[1]:g[1] + nj, g[0]:g[0] + ni, :]
            bx = boxes[l]
            dx = [(bx[2 * d + 1] - bx[2 * d]) / n for d, n in enumerate((ni, nj, nk))]      # bbx.size(d)/num_cell[d]
            ax = [bx[2 * d] + np.arange(n + 1, dtype=np.float64) * dx[d] for d, n in enumerate((ni, nj, nk))]
            if coords is not None and not isinstance(coords, identity):
                ax = [np.array([c.map(float(x)) for x in a]) for c, a in zip((coords.xcoord, coords.ycoord, coords.zcoord), ax)]
            pts = np.empty((nk + 1, nj + 1, ni + 1, 3))
            pts[..., 0], pts[..., 1], pts[..., 2] = ax[0][None, None, :], ax[1][None, :, None], ax[2][:, None, None]
            with open(os.path.join(out_dir, "data_" + basename, f"b{int(lb_glob):09d}.vts"), "wb") as f:
                f.write(b'<?xml version="1.0"?>\n<VTKFile type="StructuredGrid" version="0.1" byte_order="LittleEndian">\n')
                f.write(f'<StructuredGrid WholeExtent="0 {ni} 0 {nj} 0 {nk}">\n<Piece Extent="0 {ni} 0 {nj} 0 {nk}">\n'.encode())
                f.write(f'<CellData Scalars="{",".join(names)}">\n'.encode())
                for v, nm in enumerate(names):
                    f.write(f'<DataArray type="Float64" Name="{nm}" format="binary">\n'.encode())
                    f.write(io._b64(np.ascontiguousarray(blk[..., v]).tobytes()))
                    f.write(b'\n</DataArray>\n')
                f.write(b'</CellData>\n<Points>\n<DataArray type="Float64" NumberOfComponents="3" format="binary">\n')
                f.write(io._b64(pts.tobytes()))
                f.write(b'\n</DataArray>\n</Points>\n</Piece>\n</StructuredGrid>\n</VTKFile>\n')

    @staticmethod
    def output_vtk(out_dir, basename, array, names=None):
        """io::output_vtk(out_dir, basename, array): every rank writes the pieces of its own blocks, the root the .pvts."""
        grid, pool = array.grid, array.grid.group()
        nglob = grid.get_num_global_blocks() if grid.blocks is not None else int(pool.reduce(float(grid.num_local_blocks), RED_SUM))
        boxes = (np.array([grid.blocks.get_block_box(grid.first_block + l) for l in range(grid.num_local_blocks)])
                 if grid.blocks is not None else grid._boxes)
        if pool.isroot():
            os.makedirs(out_dir, exist_ok=True)
        pool.sync()
        io.write_vtk_files(out_dir, basename, array.to_host(), grid.num_cell, array.num_exch, boxes,
                           range(grid.first_block, grid.first_block + grid.num_local_blocks), nglob, grid.coords,
                           names or io.VAR_NAMES["prim"], write_base=pool.isroot())
        pool.sync()


def launch_count():
    return int(lib().spb_launch_count())


def device_count():
    return int(lib().spb_device_count())
