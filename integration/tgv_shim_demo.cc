// TGV-style solver written against SPADE's own API (the reference headers, unmodified, are only #included at build
// time: nvcc -x cu -I/root/reference/src). It runs the SAME solver twice on device::gpu arrays:
//   A. the reference's generic CUDA path: flux_div tag `basic`, grid::make_exchange, integrator_t (advance.h:236-280)
//   B. the drop-in: tag `b200`, b200::make_exchange, the same integrator_t call (include/spade_b200_shim.hpp)
//   C. the drop-in with the two callbacks passed as named objects (b200::flux_div_rhs, b200::exchange_bc) instead of
//      lambdas, which lets the same integrator_t call run ONE kernel per stage (RHS + stage update + ghost exchange)
// and prints one JSON line with the timings and the relative L2 differences of the final states.
// With gpus > 1 the solver runs the reference's own multi-GPU model — one host thread per GPU inside this process
// (compute_env_t::exec, compute_pool.h:497-514) — and the drop-in exchanges ghost cells by packing straight into the peer GPU's
// receive buffer over NVLink (b200::arr_exchange_t::exchange_messages).
// Usage: tgv_shim_demo [blocks_per_dim=4] [cells_per_block=32] [steps=2] [scheme: 0 central+visc | 1 hybrid+visc] [gpus=1]
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
    const int nb = argc > 1 ? std::atoi(argv[1]) : 4;
    const int nc = argc > 2 ? std::atoi(argv[2]) : 32;
    const int nsteps = argc > 3 ? std::atoi(argv[3]) : 2;
    const int scheme = argc > 4 ? std::atoi(argv[4]) : 0;
    const int ngpu = argc > 5 ? std::atoi(argv[5]) : 1;
    std::vector<int> devices;
    for (int d = 0; d < ngpu; ++d) devices.push_back(d);
    spade::parallel::compute_env_t env(&argc, &argv, devices);
    env.exec([&](spade::parallel::pool_t& pool)
    {
        const real_t gamma = 1.4, rgas = 287.15, p0 = 101325.0, t0 = 300.0, u0 = 34.7, pi = 3.14159265358979323846;
        const real_t mu = (p0/(rgas*t0))*u0/1600.0;
        spade::ctrs::array<int, 3> num_blocks(nb, nb, nb), cells(nc, nc, nc), exch(2, 2, 2);
        spade::bound_box_t<real_t, 3> bounds;
        for (int d = 0; d < 3; ++d) { bounds.min(d) = 0.0; bounds.max(d) = 2.0*pi; }
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
            q.T() = t0*(1.0 + 0.02*sin(x[0] + 2.0*x[1] - x[2]));
            q.u() = u0*sin(x[0])*cos(x[1])*cos(x[2]);
            q.v() = -u0*cos(x[0])*sin(x[1])*cos(x[2]);
            q.w() = 0.3*u0*sin(x[2])*cos(x[0] + x[1]);
            return q;
        };
        const real_t dx = 2.0*pi/(nb*nc);
        const real_t dt = 0.2*dx/(std::sqrt(gamma*rgas*t0*1.02) + 1.5*u0);

        auto run = [&](const auto& flux_func, const int use_b200, std::vector<real_t>& out, double& seconds, double& umax)
        {
            prim_t fill1 = 0.0; flux_t fill2 = 0.0;
            spade::grid::grid_array prim(grid, fill1, exch, spade::device::gpu);
            spade::grid::grid_array rhs (grid, fill2, exch, spade::device::gpu);
            spade::algs::fill_array(prim, ic);
            cons_t cstate;
            spade::fluid_state::state_transform_t trans(cstate, air);
            spade::time_integration::time_axis_t axis(real_t(0.0), dt);
            spade::time_integration::rk4_t alg;
            auto ref_handle = spade::grid::make_exchange(prim, periodic);
            auto new_handle = spade::b200::make_exchange(prim, periodic);
            auto go = [&](const auto& calc_rhs, const auto& bc)
            {
                bc(prim, real_t(0.0));
                spade::time_integration::integrator_data_t qd(std::move(prim), std::move(rhs), alg);
                spade::time_integration::integrator_t ti(axis, alg, qd, calc_rhs, bc, trans);
                ti.advance();                                           // warm-up step (also part of the trajectory)
                cudaDeviceSynchronize();
                const auto t0w = std::chrono::steady_clock::now();
                for (int n = 0; n < nsteps; ++n) ti.advance();
                cudaDeviceSynchronize();
                seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0w).count();
                auto& sol = ti.solution();
                if (use_b200) umax = spade::b200::transform_reduce(sol, spade::b200::wavespeed<decltype(air)>{air}, spade::algs::max);
                out.resize(sol.data.size());
                cudaMemcpy(out.data(), &sol.data[0], sizeof(real_t)*out.size(), cudaMemcpyDeviceToHost);
            };
            if (use_b200 == 2)
                go(spade::b200::flux_div_rhs(flux_func), spade::b200::exchange_bc(new_handle, pool));
            else if (use_b200)
                go([&](auto& rr, const auto& qq, const auto&) { spade::pde_algs::flux_div(qq, rr, flux_func, spade::algs::make_traits(spade::pde_algs::b200, spade::pde_algs::overwrite)); },
                   [&](auto& qq, const auto&) { new_handle.exchange(qq, pool); });
            else
                go([&](auto& rr, const auto& qq, const auto&) { spade::pde_algs::flux_div(qq, rr, flux_func, spade::algs::make_traits(spade::pde_algs::basic, spade::pde_algs::overwrite)); },
                   [&](auto& qq, const auto&) { ref_handle.exchange(qq, pool); });
        };

        auto both = [&](const auto& flux_func)
        {
            std::vector<real_t> qa, qb, qc;
            double ta = 0.0, tb = 0.0, tc = 0.0, umax = 0.0, dummy = 0.0;
            run(flux_func, 0, qa, ta, dummy);
            run(flux_func, 1, qb, tb, umax);
            run(flux_func, 2, qc, tc, dummy);
            double num = 0.0, numc = 0.0, den = 0.0;
            for (std::size_t i = 0; i < qa.size(); ++i)
            {
                const double d = qa[i] - qb[i], dc = qa[i] - qc[i];
                num += d*d; numc += dc*dc; den += qa[i]*qa[i];
            }
            // every rank holds its own blocks: norms and times over the whole job
            num = pool.sum(num); numc = pool.sum(numc); den = pool.sum(den);
            const auto mx = [](const double& a, const double& b) { return a > b ? a : b; };
            ta = pool.reduce(ta, mx); tb = pool.reduce(tb, mx); tc = pool.reduce(tc, mx);
            if (!pool.isroot()) return;
            const double cells_total = double(nb)*nb*nb*double(nc)*nc*nc;
            std::printf("{\"solver\": \"tgv_shim_demo\", \"blocks\": %d, \"cells_per_block\": %d, \"steps\": %d, \"scheme\": %d, "
                        "\"reference_gpu_basic_cell_stage_updates_per_s\": %.6e, \"b200_cell_stage_updates_per_s\": %.6e, "
                        "\"b200_fused_cell_stage_updates_per_s\": %.6e, \"speedup\": %.2f, \"speedup_fused\": %.2f, "
                        "\"rel_l2\": %.3e, \"rel_l2_fused\": %.3e, \"umax\": %.6f, \"gpus\": %d}\n",
                        nb, nc, nsteps, scheme, cells_total*4*nsteps/ta, cells_total*4*nsteps/tb, cells_total*4*nsteps/tc, ta/tb, ta/tc,
                        std::sqrt(num/den), std::sqrt(numc/den), umax, ngpu);
        };
        if (scheme == 0) both(spade::omni::compose(tscheme, vscheme));
        else
        {
            spade::convective::hybrid_scheme_t hyb(tscheme, wscheme, ducr, spade::convective::full_flux);
            both(spade::omni::compose(hyb, vscheme));
        }
    });
    return 0;
}
