// LES closure: viscous::visc_lr over viscous_laws::sgs_visc_t(constant_viscosity_t, subgrid_scale::wale_t)
// (reference src/navier-stokes/viscous_laws.h:175-216, subgrid_scale.h:25-91): the eddy viscosity is evaluated from the
// face gradient that the viscous flux already needs, inside the RHS / fused-stage kernel (SGS = true instantiations of
// spb_flux_div_wide.cuh), on identity and on general coordinates.
#include "spb_flux_div_wide.cuh"

namespace spb
{
    template <bool CURV>
    static int sgs_dispatch(const spb_grid* g, const double* q, double* rhs, const spb_flux_desc* f, const FluxParams& P, int increment,
                            int64_t lb_begin, int64_t lb_end, cudaStream_t stream, double* q_out, const StageParams* stage, spb_exchange* exch)
    {
#define SPB_CASE(C, D) if (f->conv == C && f->diss == D) \
            return stage ? launch_fdiv<C, D, 1, true,  CURV, true>(g, q, rhs, P, 0, lb_begin, lb_end, stream, q_out, stage, exch) \
                         : launch_fdiv<C, D, 1, false, CURV, true>(g, q, rhs, P, increment, lb_begin, lb_end, stream)
        SPB_CASE(SPB_CONV_TOTANI,     SPB_DISS_NONE);
        SPB_CASE(SPB_CONV_NONE,       SPB_DISS_NONE);
        SPB_CASE(SPB_CONV_TOTANI,     SPB_DISS_FWENO);
        SPB_CASE(SPB_CONV_CENT_KEEP4, SPB_DISS_NONE);
        SPB_CASE(SPB_CONV_CENT_KEEP4, SPB_DISS_FWENO);
#undef SPB_CASE
        set_error("spb_flux_div: the WALE closure is implemented for visc_lr alone or with totani_lr / cent_keep<4> / their hybrids");
        return SPB_ERR_UNSUPPORTED;
    }

    int flux_div_sgs(const spb_grid* g, const double* q, double* rhs, const spb_flux_desc* f, const FluxParams& P, int increment,
                     int64_t lb_begin, int64_t lb_end, cudaStream_t stream, double* q_out, const StageParams* stage, spb_exchange* exch)
    {
        if (!f->visc) { set_error("spb_flux_div: an SGS model needs visc_lr"); return SPB_ERR_BAD_ARG; }
        if (!(f->sgs_prt > 0.0)) { set_error("spb_flux_div: wale_t needs a positive turbulent Prandtl number"); return SPB_ERR_BAD_ARG; }
        return g->metric_dev ? sgs_dispatch<true>(g, q, rhs, f, P, increment, lb_begin, lb_end, stream, q_out, stage, exch)
                             : sgs_dispatch<false>(g, q, rhs, f, P, increment, lb_begin, lb_end, stream, q_out, stage, exch);
    }
}
