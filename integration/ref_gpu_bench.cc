// The reference's OWN CUDA path on this GPU, next to the drop-in, on the same solver:
// a TGV-style solver written against SPADE's API, compiled against the UNMODIFIED reference headers with nvcc -arch=sm_100a,
// run once per valid flux_div algorithm tag of the reference
//   basic  (flux_div_basic.h:17-77, always valid)
//   longf  (flux_div_longf.h)
//   fldbc  (flux_div_fldbc.h:21-564: shared-memory variant for stencils with >= 2 exchange cells and a gradient)
//   fused  (flux_div_fused.h:20-268: WENO-width stencils)
// and once through the drop-in (tag b200 + b200::flux_div_rhs / exchange_bc: one kernel per stage). A variant counts as valid
// when its final state agrees with tag `basic` to 1e-10 relative L2 (the shared-memory variants are not valid for every
// functor set). Prints one JSON line: cell-stage-updates/s per variant, the best valid reference variant, and the
// drop-in's speed-up over it. bench.py runs this binary for its `ref_gpu_baseline` record.
// Usage: ref_gpu_bench [blocks_per_dim=8] [cells_per_block=32] [steps=2] [scheme: 0 totani_lr+visc_lr | 1 hybrid(totani,fweno,ducros)+visc_lr | 2 cent_keep<4>+visc_lr]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <string>
#include <vector>
#include "spade.h"
#include "spade_b200_shim.hpp"

using real_t = double;
using prim_t = spade::fluid_state::prim_t<real_t>;
using cons_t = spade::fluid_state::cons_t<real_t>;
using flux_t = spade::fluid_state::flux_t<real_t>;

struct result_t { std::string tag; double value = 0.0, rel_l2 = -1.0, seconds = 0.0; bool ran = false; };

int main(int argc, char** argv)
{
    const int nb = argc > 1 ? std::atoi(argv[1]) : 8;
    const int nc = argc > 2 ? std::atoi(argv[2]) : 32;
    const int nsteps = argc > 3 ? std::atoi(argv[3]) : 2;
    const int scheme = argc > 4 ? std::atoi(argv[4]) : 0;
    std::vector<int> devices{0};
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
        const auto c4scheme = spade::convective::cent_keep<4>(air);
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
        const double cells_total = double(nb)*nb*nb*double(nc)*nc*nc;

        // one solver run: warm-up step, then nsteps timed (host clock around device-synchronous work, like the reference's own drivers)
        auto solve = [&](const auto& calc_rhs_of, const auto& bc_of, std::vector<real_t>& out, double& seconds)
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
            const auto calc_rhs = calc_rhs_of();
            const auto bc = bc_of(ref_handle, new_handle);
            bc(prim, real_t(0.0));
            spade::time_integration::integrator_data_t qd(std::move(prim), std::move(rhs), alg);
            spade::time_integration::integrator_t ti(axis, alg, qd, calc_rhs, bc, trans);
            ti.advance();
            cudaDeviceSynchronize();
            const auto t0w = std::chrono::steady_clock::now();
            for (int n = 0; n < nsteps; ++n) ti.advance();
            cudaDeviceSynchronize();
            seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0w).count();
            auto& sol = ti.solution();
            out.resize(sol.data.size());
            cudaMemcpy(out.data(), &sol.data[0], sizeof(real_t)*out.size(), cudaMemcpyDeviceToHost);
        };

        std::vector<real_t> base;
        std::vector<result_t> results;
        const auto rel_l2 = [&](const std::vector<real_t>& a)
        {
            double num = 0.0, den = 0.0;
            for (std::size_t i = 0; i < a.size(); ++i) { const double d = a[i] - base[i]; num += d*d; den += base[i]*base[i]; }
            return std::sqrt(num/den);
        };
        const auto ref_bc = [&](auto& ref_handle, auto&) { return [&](auto& qq, const auto&) { ref_handle.exchange(qq, pool); }; };
        const auto run_tag = [&](const auto& flux_func, const auto& tag, const char* name)
        {
            result_t r; r.tag = name;
            std::vector<real_t> q;
            try
            {
                solve([&] { return [&](auto& rr, const auto& qq, const auto&) { spade::pde_algs::flux_div(qq, rr, flux_func, spade::algs::make_traits(tag, spade::pde_algs::overwrite)); }; },
                      ref_bc, q, r.seconds);
                if (cudaGetLastError() != cudaSuccess) throw std::runtime_error("cuda error");
                if (base.empty()) base = q;
                r.rel_l2 = rel_l2(q);
                r.value = cells_total*4*nsteps/r.seconds;
                r.ran = true;
            }
            catch (const std::exception& e) { std::fprintf(stderr, "ref_gpu_bench: tag %s failed: %s\n", name, e.what()); }
            results.push_back(r);
        };
        const auto run_b200 = [&](const auto& flux_func)
        {
            result_t r; r.tag = "b200";
            std::vector<real_t> q;
            solve([&] { return spade::b200::flux_div_rhs(flux_func); },
                  [&](auto&, auto& new_handle) { return spade::b200::exchange_bc(new_handle, pool); }, q, r.seconds);
            r.rel_l2 = rel_l2(q);
            r.value = cells_total*4*nsteps/r.seconds;
            r.ran = true;
            results.push_back(r);
        };

        const char* scheme_name = "";
        if (scheme == 0)
        {
            scheme_name = "totani_lr + visc_lr";
            const auto f = spade::omni::compose(tscheme, vscheme);
            run_tag(f, spade::pde_algs::basic, "basic");
            run_tag(f, spade::pde_algs::longf, "longf");
            run_b200(f);
        }
        else if (scheme == 1)
        {
            scheme_name = "hybrid(totani_lr, fweno_t, ducros_t, full_flux) + visc_lr";
            spade::convective::hybrid_scheme_t hyb(tscheme, wscheme, ducr, spade::convective::full_flux);
            const auto f = spade::omni::compose(hyb, vscheme);
            run_tag(f, spade::pde_algs::basic, "basic");
            run_tag(f, spade::pde_algs::fldbc, "fldbc");
            run_tag(f, spade::pde_algs::fused, "fused");
            run_b200(f);
        }
        else
        {
            scheme_name = "cent_keep<4> + visc_lr";
            const auto f = spade::omni::compose(c4scheme, vscheme);
            run_tag(f, spade::pde_algs::basic, "basic");
            run_tag(f, spade::pde_algs::fldbc, "fldbc");
            run_b200(f);
        }
        if (!pool.isroot()) return;
        const result_t* best = nullptr;
        const result_t* ours = nullptr;
        for (const auto& r: results)
        {
            if (r.tag == "b200") { ours = &r; continue; }
            if (r.ran && r.rel_l2 >= 0.0 && r.rel_l2 < 1e-10 && (!best || r.value > best->value)) best = &r;
        }
        std::printf("{\"solver\": \"ref_gpu_bench\", \"functor_set\": \"%s\", \"grid\": \"%dx%dx%d blocks of %d^3\", \"steps\": %d, \"variants\": {",
                    scheme_name, nb, nb, nb, nc, nsteps);
        bool first = true;
        for (const auto& r: results)
        {
            if (r.tag == "b200") continue;
            std::printf("%s\"%s\": {\"value\": %.6e, \"rel_l2_vs_basic\": %.3e, \"valid\": %s}", first ? "" : ", ", r.tag.c_str(), r.value, r.rel_l2,
                        (r.ran && r.rel_l2 >= 0.0 && r.rel_l2 < 1e-10) ? "true" : "false");
            first = false;
        }
        std::printf("}, \"unit\": \"cell-stage-updates/s\", \"tag\": \"%s\", \"value\": %.6e, \"b200_value\": %.6e, \"b200_rel_l2_vs_basic\": %.3e, "
                    "\"speedup_over_best_valid\": %.2f}\n",
                    best ? best->tag.c_str() : "none", best ? best->value : 0.0, ours ? ours->value : 0.0, ours ? ours->rel_l2 : -1.0,
                    (best && ours && best->value > 0.0) ? ours->value/best->value : 0.0);
    });
    return 0;
}
