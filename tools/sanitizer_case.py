"""compute-sanitizer case (tools/gpu_visit.sh sanitizer): one RK4 step, an incremented flux_div and a source term for a spread of
functor sets, block shapes, exchange depths and both coordinate systems on small wall-bounded grids, then a generic-advance step."""

import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from util import GAMMA, RGAS, make_state, product_flux
import spade_b200.api as sp
bounds = [0.0, 2 * np.pi, -1.0, 1.0, 0.5, 2.0]
gas = sp.ideal_gas_t(GAMMA, RGAS)
for coords_on in (False, True):
    for nb, n, ng, schemes in (((2, 1, 2), (40, 12, 8), 2, (0, 1, 11)), ((1, 2, 1), (16, 16, 8), 2, (0, 3, 12)), ((1, 1, 2), (12, 20, 6), 4, (13, 14))):
        coords = sp.diagonal_coords(sp.scaled_coord_1D(2.0), sp.integrated_tanh_1D(-1.0, 1.0, 0.1, 4.0), sp.quad_1D()) if coords_on else sp.identity()
        grid = sp.cartesian_grid_t(n, sp.cartesian_blocks_t(nb, bounds), coords, sp.pool_t(0, 1))
        q0 = make_state(nb, n, ng, seed=3, bounds=bounds)
        for scheme in schemes:
            flux = sp.flux_desc(product_flux(scheme))
            qa, ra = sp.grid_array.from_host(grid, q0, (ng,) * 3), sp.grid_array(grid, 0.0, (ng,) * 3)
            ex = sp.make_exchange(qa, (1, 0, 1))
            bc = sp.exchange_bc_t(ex, sp.boundary.ymin | sp.boundary.ymax, sp.noslip_isothermal_wall(300.0))
            bc(qa, 0.0)
            ti = sp.integrator_t(sp.time_axis_t(0.0, 1e-7), sp.rk4_t, sp.integrator_data_t(qa, ra, sp.rk4_t), sp.flux_div_rhs_t(flux, sp.overwrite),
                                 bc, sp.state_transform_t(gas))
            ti.advance()
            sp.flux_div(qa, ra, flux, sp.increment)
            sp.source_term(qa, ra, sp.body_force_t(1.0, 0.5, 0.25))
            print("ok", coords_on, nb, n, ng, scheme, float(ti.solution().data.abs().max()), flush=True)
    tg = sp.integrator_t(sp.time_axis_t(0.0, 1e-7), sp.ssprk3hs_t, sp.integrator_data_t(qa, ra, sp.ssprk3hs_t),
                         lambda r, qq, t: sp.flux_div(qq, r, flux, sp.overwrite), lambda qq, t: ex.exchange(qq))
    tg.advance()
    print("ok generic", float(tg.solution().data.abs().max()), flush=True)
