// Compile-only check of the shim's functor recognition (include/spade_b200_shim.hpp): every functor type of the implemented set
// must be accepted by spade::b200::flux_desc and fill the POD descriptor with the functor's own members. Built by
// integration/Makefile (no GPU needed to build; running it only prints the descriptors).
#include <cstdio>
#include "spade.h"
#include "spade_b200_shim.hpp"

using real_t = double;

static void show(const char* name, const spb_flux_desc& d)
{
    std::printf("%-44s conv=%d diss=%d blend=%d visc=%d sgs=%d gamma=%g R=%g mu=%g beta=%g prinv=%g eps=%g cw=%g delta=%g prt=%g\n", name, d.conv, d.diss,
                d.blend, d.visc, d.sgs, d.gamma, d.R, d.mu, d.beta, d.prandtl_inv, d.sensor_eps, d.sgs_cw, d.sgs_delta, d.sgs_prt);
}

int main()
{
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
