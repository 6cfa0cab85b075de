// General (diagonal) coordinates: pde_algs::flux_div and the fused RK stage on stretched grids
// (reference src/pde-algs/flux-div/flux_div_basic.h:49-71 with coords::calc_jacobian / calc_normal_vector,
// src/core/coord_system.h:250-267,295-302). Every functor combination runs through the wide kernel template with
// CURV = true (spb_flux_div_wide.cuh); the metric comes from the per-block 1-D tables of spb_grid_set_metric.
#include "spb_flux_div_wide.cuh"

namespace spb
{
    int flux_div_curv(const spb_grid* g, const double* q, double* rhs, const spb_flux_desc* f, const FluxParams& P, int increment,
                      int64_t lb_begin, int64_t lb_end, cudaStream_t stream, double* q_out, const StageParams* stage, spb_exchange* exch)
    {
#define SPB_CASE(C, D, V) if (f->conv == C && f->diss == D && (f->visc != 0) == (V != 0)) \
            return stage ? launch_fdiv<C, D, V, true,  true>(g, q, rhs, P, 0, lb_begin, lb_end, stream, q_out, stage, exch) \
                         : launch_fdiv<C, D, V, false, true>(g, q, rhs, P, increment, lb_begin, lb_end, stream)
        SPB_CASE(SPB_CONV_TOTANI,     SPB_DISS_NONE,  1);
        SPB_CASE(SPB_CONV_TOTANI,     SPB_DISS_NONE,  0);
        SPB_CASE(SPB_CONV_NONE,       SPB_DISS_NONE,  1);
        SPB_CASE(SPB_CONV_TOTANI,     SPB_DISS_FWENO, 1);
        SPB_CASE(SPB_CONV_CENT_KEEP4, SPB_DISS_NONE,  1);
        SPB_CASE(SPB_CONV_CENT_KEEP4, SPB_DISS_FWENO, 1);
        SPB_CASE(SPB_CONV_CENT_KEEP4, SPB_DISS_NONE,  0);
        SPB_CASE(SPB_CONV_FWENO,      SPB_DISS_NONE,  0);
        SPB_CASE(SPB_CONV_TOTANI,     SPB_DISS_FWENO, 0);
        SPB_CASE(SPB_CONV_CENT_KEEP6, SPB_DISS_NONE,  1);
        SPB_CASE(SPB_CONV_CENT_KEEP8, SPB_DISS_NONE,  1);
#undef SPB_CASE
        set_error("spb_flux_div: this combination of flux functors is not in the implemented set");
        return SPB_ERR_UNSUPPORTED;
    }
}
