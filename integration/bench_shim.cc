// Weak-scaling benchmark of the drop-in THROUGH THE C++ SHIM, in the reference's own multi-GPU model: one process, one host
// thread per GPU (compute_env_t::exec, compute_pool.h:497-514). A TGV solver written only against SPADE's API — grid,
// grid_array, make_exchange, integrator_t — with the two callbacks passed as the shim's named types, so that
// integrator_t::advance() runs one kernel per stage and block range, the rank-boundary blocks first on a side stream, their
// messages packed straight into the neighbour GPU's buffer behind stream-ordered flags (include/spade_b200_shim.hpp).
// Per GPU: blocks_xy x blocks_xy x blocks_z blocks of cells^3 (default 8 x 8 x 8 of 32^3 = 256^3, BASELINE config 4); rank r
// owns z-slab r of the 8 x 8 x 8G lattice (SPADE's contiguous partition).
// Prints one JSON line: cell-stage-updates/s of the whole job (max over ranks of the host time around device-synchronous
// advance() calls between pool barriers).
// Usage: bench_shim [gpus=1] [steps=20] [blocks_xy=8] [blocks_z_per_gpu=8] [cells=32] [scheme: 0 central+visc | 1 hybrid+visc]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "spade.h"
#include "spade_b200_shim.hpp"

using real_t = double;
using prim_t = spade::fluid_state::prim_t<real_t>;
using cons_t = spade::fluid_state::cons_t<real_t>;
using flux_t = spade::fluid_state::flux_t<real_t>;

int main(int argc, char** argv)
{
    const int ngpu   = argc > 1 ? std::atoi(argv[1]) : 1;
    const int nsteps = argc > 2 ? std::atoi(argv[2]) : 20;
    const int nbxy   = argc > 3 ? std::atoi(argv[3]) : 8;
    const int nbz    = argc > 4 ? std::atoi(argv[4]) : 8;
    const int nc     = argc > 5 ? std::atoi(argv[5]) : 32;
    const int scheme = argc > 6 ? std::atoi(argv[6]) : 0;
    std::vector<int> devices;
    for (int d = 0; d < ngpu; ++d) devices.push_back(d);
    spade::parallel::compute_env_t env(&argc, &argv, devices);
    env.exec([&](spade::parallel::pool_t& pool)
    {
        const real_t gamma = 1.4, rgas = 287.15, p0 = 101325.0, t0 = 300.0, u0 = 34.7, pi = 3.14159265358979323846;
        const real_t mu = (p0/(rgas*t0))*u0/1600.0;
        spade::ctrs::array<int, 3> num_blocks(nbxy, nbxy, nbz*ngpu), cells(nc, nc, nc), exch(2, 2, 2);
        spade::bound_box_t<real_t, 3> bounds;
        for (int d = 0; d < 3; ++d) { bounds.min(d) = 0.0; bounds.max(d) = 2.0*pi; }
        bounds.max(2) = 2.0*pi*ngpu*real_t(nbz)/real_t(nbxy);
        spade::coords::identity<real_t> coords;
        spade::grid::cartesian_blocks_t blocks(num_blocks, bounds);
        spade::grid::cartesian_grid_t grid(cells, blocks, coords, pool);
        spade::ctrs::array<bool, 3> periodic(true, true, true);
        spade::fluid_state::ideal_gas_t<real_t> air(gamma, rgas);
        spade::viscous_laws::constant_viscosity_t<real_t> vlaw(mu, 0.72);
        spade::convective::totani_lr tscheme(air);
        spade::convective::fweno_t<decltype(air)> wscheme(air);
        spade::state_sensor::ducros_t<real_t> ducr(1e-2);
        spade::viscous::visc_lr vscheme(vlaw, air);
        const auto ic = [=] _sp_hybrid (const spade::coords::point_t<real_t>& x)
        {
            prim_t q;
            q.p() = p0 + (p0/(rgas*t0))*u0*u0/16.0*(cos(2.0*x[0]) + cos(2.0*x[1]))*(cos(2.0*x[2]) + 2.0);
            q.T() = t0;
            q.u() = u0*sin(x[0])*cos(x[1])*cos(x[2]);
            q.v() = -u0*cos(x[0])*sin(x[1])*cos(x[2]);
            q.w() = 0.0;
            return q;
        };
        const real_t dx = 2.0*pi/(nbxy*nc);
        const real_t dt = 0.2*dx/(std::sqrt(gamma*rgas*t0) + u0);

        const auto run = [&](const auto& flux_func)
        {
            prim_t fill1 = 0.0; flux_t fill2 = 0.0;
            spade::grid::grid_array prim(grid, fill1, exch, spade::device::gpu);
            spade::grid::grid_array rhs (grid, fill2, exch, spade::device::gpu);
            spade::algs::fill_array(prim, ic);
            cons_t cstate;
            spade::fluid_state::state_transform_t trans(cstate, air);
            spade::time_integration::time_axis_t axis(real_t(0.0), dt);
            spade::time_integration::rk4_t alg;
            auto handle = spade::b200::make_exchange(prim, periodic);
            const auto calc_rhs = spade::b200::flux_div_rhs(flux_func);
            const auto bc = spade::b200::exchange_bc(handle, pool);
            bc(prim, real_t(0.0));
            const double umax0 = spade::b200::transform_reduce(prim, spade::b200::wavespeed<decltype(air)>{air}, spade::algs::max);
            spade::time_integration::integrator_data_t qd(std::move(prim), std::move(rhs), alg);
            spade::time_integration::integrator_t ti(axis, alg, qd, calc_rhs, bc, trans);
            for (int n = 0; n < 3; ++n) ti.advance();
            cudaDeviceSynchronize();
            pool.sync();
            const auto t0w = std::chrono::steady_clock::now();
            for (int n = 0; n < nsteps; ++n) ti.advance();
            cudaDeviceSynchronize();
            pool.sync();
            double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0w).count();
            sec = pool.reduce(sec, [](const double& a, const double& b) { return a > b ? a : b; });
            const double umax = spade::b200::transform_reduce(ti.solution(), spade::b200::wavespeed<decltype(air)>{air}, spade::algs::max);
            if (!pool.isroot()) return;
            const double cells_total = double(nbxy)*nbxy*double(nbz)*ngpu*double(nc)*nc*nc;
            std::printf("{\"solver\": \"bench_shim\", \"host\": \"C++20 shim, one host thread per GPU in one process\", \"n_gpus\": %d, \"steps\": %d, "
                        "\"workload\": \"TGV %dx%dx%d cells (%dx%dx%d blocks of %d^3), %s + visc_lr, rk4_t\", "
                        "\"value\": %.6e, \"unit\": \"cell-stage-updates/s\", \"ms_per_step\": %.4f, \"umax_start\": %.6f, \"umax_end\": %.6f, \"finite\": %s}\n",
                        ngpu, nsteps, nbxy*nc, nbxy*nc, nbz*ngpu*nc, nbxy, nbxy, nbz*ngpu, nc, scheme ? "hybrid(totani_lr,fweno_t,ducros_t)" : "totani_lr",
                        cells_total*4*nsteps/sec, 1e3*sec/nsteps, umax0, umax, (umax == umax && umax < 10.0*umax0) ? "true" : "false");
        };
        if (scheme == 0) run(spade::omni::compose(tscheme, vscheme));
        else
        {
            spade::convective::hybrid_scheme_t hyb(tscheme, wscheme, ducr, spade::convective::full_flux);
            run(spade::omni::compose(hyb, vscheme));
        }
    });
    return 0;
}
