// Compile-only check of the shim's functor recognition (include/spade_b200_shim.hpp): every functor type of the implemented set
// must be accepted by spade::b200::flux_desc and fill the POD descriptor with the functor's own members. Built by
// integration/Makefile (no GPU needed to build; running it only prints the descriptors).
#include <cstdio>
#include "spade.h"
#include "spade_b200_shim.hpp"

using real_t = double;

static void show(const char* name, const spb_flux_desc& d)
{
    std::printf("%-44s conv=%d diss=%d blend=%d visc=%d sgs=%d linear=%d gamma=%g R=%g mu=%g beta=%g prinv=%g eps=%g cw=%g delta=%g prt=%g\n", name, d.conv, d.diss,
                d.blend, d.visc, d.sgs, d.weno_linear, d.gamma, d.R, d.mu, d.beta, d.prandtl_inv, d.sensor_eps, d.sgs_cw, d.sgs_delta, d.sgs_prt);
}

// Instantiated but never executed without a GPU: the rest of the shim's entry points on a device::gpu array
static void compile_only(spade::parallel::pool_t& pool)
{
    using prim_t = spade::fluid_state::prim_t<real_t>;
    using flux_t = spade::fluid_state::flux_t<real_t>;
    spade::ctrs::array<int, 3> num_blocks(2, 2, 2), cells(16, 16, 16), exch(2, 2, 2);
    spade::bound_box_t<real_t, 3> bounds;
    for (int d = 0; d < 3; ++d) { bounds.min(d) = 0.0; bounds.max(d) = 1.0; }
    spade::coords::identity<real_t> coords;
    spade::grid::cartesian_blocks_t blocks(num_blocks, bounds);
    spade::grid::cartesian_grid_t grid(cells, blocks, coords, pool);
    prim_t f1 = 0.0; flux_t f2 = 0.0;
    spade::grid::grid_array q(grid, f1, exch, spade::device::gpu);
    spade::grid::grid_array r(grid, f2, exch, spade::device::gpu);
    spade::ctrs::array<bool, 3> periodic(true, false, true);
    auto handle = spade::b200::make_exchange(q, periodic);
    handle.exchange(q, pool);
    const auto walls = spade::boundary::ymin || spade::boundary::ymax;
    spade::b200::boundary_fill(q, walls, spade::b200::mirror_kernel::noslip_adiabatic());
    spade::b200::boundary_fill(q, walls, spade::b200::mirror_kernel::symmetry());
    spade::b200::boundary_fill(q, spade::boundary::ymax, spade::boundary::extrapolate<2>);
    spade::b200::source_term(q, r, spade::b200::body_force{{1.0, 0.0, 0.0}});
    spade::fluid_state::ideal_gas_t<real_t> air(1.4, 287.15);
    (void)spade::b200::transform_reduce(q, spade::b200::wavespeed<decltype(air)>{air}, spade::algs::max);
    (void)spade::b200::transform_reduce(q, spade::b200::kinetic_energy<decltype(air)>{air}, spade::algs::sum);
    (void)spade::b200::transform_reduce(q, spade::b200::variable{0}, spade::algs::sum);
    (void)spade::b200::transform_reduce(q, spade::b200::abs_variable{2}, spade::algs::max);
    spade::b200::binary_write("/tmp/spade_b200_check.bin", q);
    spade::b200::binary_read("/tmp/spade_b200_check.bin", q);
    spade::b200::invalidate(grid);
    {
        // every integrate_advance overload of the shim: ssprk3_opt (advance.h:359-402), the generic path with
        // identity_transform (advance.h:109-230) incl. a high-storage table, the fused prim/cons path with opaque callbacks
        // and with the named callbacks (one kernel per stage, overlapped schedule)
        using cons_t = spade::fluid_state::cons_t<real_t>;
        cons_t cstate;
        spade::fluid_state::state_transform_t trans(cstate, air);
        spade::viscous_laws::constant_viscosity_t<real_t> vl(1.8e-5, 0.72);
        const auto fl = spade::omni::compose(spade::convective::totani_lr(air), spade::viscous::visc_lr(vl, air));
        const auto rhs_l = [&](auto& rr, const auto& qq, const auto&) { spade::pde_algs::flux_div(qq, rr, fl, spade::algs::make_traits(spade::pde_algs::b200, spade::pde_algs::overwrite)); };
        const auto bc_l  = [&](auto& qq, const auto&) { handle.exchange(qq, pool); };
        spade::time_integration::time_axis_t axis(real_t(0.0), real_t(1e-6));
        {
            spade::time_integration::integrator_data_t qd(q, r, spade::time_integration::ssprk3_opt);
            spade::time_integration::integrator_t ti(axis, spade::time_integration::ssprk3_opt, qd, rhs_l, bc_l, trans);
            ti.advance();
        }
        {
            spade::time_integration::rk4_t alg;
            spade::time_integration::integrator_data_t qd(q, r, alg);
            spade::time_integration::integrator_t ti(axis, alg, qd, rhs_l, bc_l);
            ti.advance();
        }
        {
            spade::time_integration::ssprk3hs_t alg;
            spade::time_integration::integrator_data_t qd(q, r, alg);
            spade::time_integration::integrator_t ti(axis, alg, qd, rhs_l, bc_l);
            ti.advance();
        }
        {
            spade::time_integration::ssprk34_t alg;
            spade::time_integration::integrator_data_t qd(q, r, alg);
            spade::time_integration::integrator_t t1(axis, alg, qd, rhs_l, bc_l, trans);
            t1.advance();
            const auto rhs_n = spade::b200::flux_div_rhs(fl);
            const auto bc_n  = spade::b200::exchange_bc(handle, pool);
            spade::time_integration::integrator_t t2(axis, alg, qd, rhs_n, bc_n, trans);
            t2.advance();
        }
    }
    spade::pde_algs::flux_div(q, r, spade::convective::cent_keep<6>(air), spade::algs::make_traits(spade::pde_algs::b200, spade::pde_algs::increment));
}

int main(int argc, char** argv)
{
    if (argc > 1000)     // never: the calls above need a GPU; they are here to be instantiated
    {
        std::vector<int> devices{0};
        spade::parallel::compute_env_t env(&argc, &argv, devices);
        env.exec([&](spade::parallel::pool_t& pool) { compile_only(pool); });
    }
    spade::fluid_state::ideal_gas_t<real_t> air(1.4, 287.15);
    spade::viscous_laws::constant_viscosity_t<real_t> vlaw(1.8e-5, 0.72);
    spade::subgrid_scale::wale_t eddy(air, real_t(0.55), real_t(0.1), real_t(0.9));
    spade::viscous_laws::sgs_visc_t slaw(vlaw, eddy);
    spade::convective::totani_lr tscheme(air);
    spade::convective::fweno_t<decltype(air)> fweno(air);
    spade::convective::rusanov_t rus(air);
    spade::convective::weno_t wrus(rus);
    spade::state_sensor::ducros_t<real_t> ducr(1e-2);
    spade::viscous::visc_lr vscheme(vlaw, air);
    spade::viscous::visc_lr sscheme(slaw, air);
    using spade::b200::flux_desc;
    using spade::omni::compose;
    show("totani_lr", flux_desc(tscheme));
    show("cent_keep<2>", flux_desc(spade::convective::cent_keep<2>(air)));
    show("cent_keep<4> + visc_lr", flux_desc(compose(spade::convective::cent_keep<4>(air), vscheme)));
    show("cent_keep<6> + visc_lr", flux_desc(compose(spade::convective::cent_keep<6>(air), vscheme)));
    show("cent_keep<8>", flux_desc(spade::convective::cent_keep<8>(air)));
    show("fweno_t", flux_desc(fweno));
    show("weno_t<rusanov_t>", flux_desc(wrus));
    show("fweno_t<disable_smooth>", flux_desc(spade::convective::fweno_t<decltype(air), spade::convective::disable_smooth>(air)));
    show("weno_t<rusanov_t, disable_smooth>", flux_desc(spade::convective::weno_t<decltype(rus), spade::convective::disable_smooth>(rus)));
    show("visc_lr", flux_desc(vscheme));
    show("totani_lr + visc_lr<sgs_visc_t<wale_t>>", flux_desc(compose(tscheme, sscheme)));
    {
        spade::convective::hybrid_scheme_t hyb(tscheme, fweno, ducr, spade::convective::full_flux);
        show("hybrid(totani, fweno, ducros, full) + visc", flux_desc(compose(hyb, vscheme)));
        show("hybrid(totani, fweno, ducros, full) + wale", flux_desc(compose(hyb, sscheme)));
    }
    {
        spade::convective::hybrid_scheme_t hyb(spade::convective::cent_keep<4>(air), fweno, ducr, spade::convective::diss_flux);
        show("hybrid(cent_keep<4>, fweno, ducros, diss)", flux_desc(compose(hyb, vscheme)));
    }
    {
        spade::convective::hybrid_scheme_t hyb(tscheme, wrus, ducr, spade::convective::full_flux);
        show("hybrid(totani, weno_t<rusanov_t>, ducros)", flux_desc(compose(hyb, vscheme)));
    }
    return 0;
}
