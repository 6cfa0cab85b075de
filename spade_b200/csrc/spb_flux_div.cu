// C entry points of the RHS: pde_algs::flux_div and the fused RK stage (reference src/pde-algs/flux-div/flux_div.h:23-41,
// src/time-integration/advance.h:254-275). Kernels: spb_flux_div_wide.cuh (wide stencils), spb_flux_div_narrow.cu
// (one-ghost-cell functors), spb_flux_div_curv.cu (general coordinates).
#include "spb_flux_div_wide.cuh"

namespace spb
{
    template <int CONV, int VISC>
    int launch_fdiv_narrow(const spb_grid* g, const double* q, double* rhs, const FluxParams& P, int increment,
                           int64_t lb_begin, int64_t lb_end, cudaStream_t stream, double* q_out, const StageParams* stage,
                           spb_exchange* exch);     // spb_flux_div_narrow.cu

}

extern "C"
{
    int spb_flux_div_blocks(const spb_grid* g, const double* q_dev, double* rhs_dev, const spb_flux_desc* f,
                            int increment, int64_t lb_begin, int64_t lb_end, void* stream)
    {
        using namespace spb;
        if (!g || !q_dev || !rhs_dev || !f) { set_error("spb_flux_div: null argument"); return SPB_ERR_BAD_ARG; }
        if (lb_begin < 0 || lb_end > g->nlb || lb_begin > lb_end) { set_error("spb_flux_div: bad block range"); return SPB_ERR_BAD_ARG; }
        const FluxParams P = make_params(f);
        cudaStream_t st = (cudaStream_t)stream;
        if (f->sgs == SPB_SGS_WALE) return flux_div_sgs(g, q_dev, rhs_dev, f, P, increment, lb_begin, lb_end, st, nullptr, nullptr, nullptr);
        if (f->sgs != SPB_SGS_NONE) { set_error("spb_flux_div: unknown SGS model"); return SPB_ERR_BAD_ARG; }
        if (g->metric_dev)
        {
            // general coordinates: the one-ghost-cell functors stay on the narrow kernel (its CURV instantiation), the rest and
            // non-uniform lattices run on the wide kernel
            int rc = SPB_ERR_UNSUPPORTED;
#define SPB_NARROW(C, V) if (f->conv == C && f->diss == SPB_DISS_NONE && (f->visc != 0) == (V != 0)) \
                rc = launch_fdiv_narrow<C, V>(g, q_dev, rhs_dev, P, increment, lb_begin, lb_end, st, nullptr, nullptr, nullptr)
            SPB_NARROW(SPB_CONV_TOTANI, 1);
            SPB_NARROW(SPB_CONV_TOTANI, 0);
            SPB_NARROW(SPB_CONV_NONE,   1);
#undef SPB_NARROW
            if (rc != SPB_ERR_UNSUPPORTED) return rc;
            return flux_div_curv(g, q_dev, rhs_dev, f, P, increment, lb_begin, lb_end, st, nullptr, nullptr, nullptr);
        }
#define SPB_CASE(C, D, V) if (f->conv == C && f->diss == D && (f->visc != 0) == (V != 0)) \
            return launch_fdiv<C, D, V>(g, q_dev, rhs_dev, P, increment, lb_begin, lb_end, st)
#define SPB_NARROW(C, V) if (f->conv == C && f->diss == SPB_DISS_NONE && (f->visc != 0) == (V != 0)) \
            return launch_fdiv_narrow<C, V>(g, q_dev, rhs_dev, P, increment, lb_begin, lb_end, st, nullptr, nullptr, nullptr)
        SPB_NARROW(SPB_CONV_TOTANI, 1);
        SPB_NARROW(SPB_CONV_TOTANI, 0);
        SPB_NARROW(SPB_CONV_NONE,   1);
#undef SPB_NARROW
        SPB_CASE(SPB_CONV_TOTANI,     SPB_DISS_FWENO, 1);
        SPB_CASE(SPB_CONV_CENT_KEEP4, SPB_DISS_NONE,  1);
        SPB_CASE(SPB_CONV_CENT_KEEP4, SPB_DISS_FWENO, 1);
        SPB_CASE(SPB_CONV_CENT_KEEP4, SPB_DISS_NONE,  0);
        SPB_CASE(SPB_CONV_FWENO,      SPB_DISS_NONE,  0);
        SPB_CASE(SPB_CONV_TOTANI,     SPB_DISS_FWENO, 0);
        SPB_CASE(SPB_CONV_CENT_KEEP6, SPB_DISS_NONE,  1);
        SPB_CASE(SPB_CONV_CENT_KEEP6, SPB_DISS_NONE,  0);
        SPB_CASE(SPB_CONV_CENT_KEEP8, SPB_DISS_NONE,  1);
        SPB_CASE(SPB_CONV_CENT_KEEP8, SPB_DISS_NONE,  0);
#undef SPB_CASE
        set_error("spb_flux_div: this combination of flux functors is not in the implemented set");
        return SPB_ERR_UNSUPPORTED;
    }

    // the functor sets the fused stage exists for (one place: both host sides ask here instead of keeping their own list)
    int spb_flux_div_rk_stage_supported(const spb_flux_desc* f)
    {
        if (!f) return 0;
        const bool hybrid_or_plain = f->diss == SPB_DISS_NONE || f->diss == SPB_DISS_FWENO;
        if (f->sgs == SPB_SGS_WALE)
            return f->visc && hybrid_or_plain && ((f->conv == SPB_CONV_NONE && f->diss == SPB_DISS_NONE) || f->conv == SPB_CONV_TOTANI || f->conv == SPB_CONV_CENT_KEEP4);
        if (f->sgs != SPB_SGS_NONE) return 0;
        const bool narrow = f->diss == SPB_DISS_NONE && (f->conv == SPB_CONV_TOTANI || (f->conv == SPB_CONV_NONE && f->visc));
        const bool wide = f->visc && ((f->conv == SPB_CONV_TOTANI && f->diss == SPB_DISS_FWENO)
                                      || (f->conv == SPB_CONV_CENT_KEEP4 && hybrid_or_plain)
                                      || ((f->conv == SPB_CONV_CENT_KEEP6 || f->conv == SPB_CONV_CENT_KEEP8) && f->diss == SPB_DISS_NONE));
        return (narrow || wide) ? 1 : 0;
    }

    int spb_flux_div_rk_stage(const spb_grid* g, const double* q_in, double* q_out, const spb_flux_desc* f,
                              const spb_stage_desc* sd, int64_t lb_begin, int64_t lb_end, void* stream)
    {
        return spb_flux_div_rk_stage_exchange(g, q_in, q_out, f, sd, nullptr, lb_begin, lb_end, stream);
    }

    int spb_flux_div_rk_stage_exchange(const spb_grid* g, const double* q_in, double* q_out, const spb_flux_desc* f,
                                       const spb_stage_desc* sd, spb_exchange* exch, int64_t lb_begin, int64_t lb_end, void* stream)
    {
        using namespace spb;
        if (!g || !q_in || !q_out || !f || !sd || q_in == q_out) { set_error("spb_flux_div_rk_stage: bad argument (q_out must differ from q_in)"); return SPB_ERR_BAD_ARG; }
        if (lb_begin < 0 || lb_end > g->nlb || lb_begin > lb_end) { set_error("spb_flux_div_rk_stage: bad block range"); return SPB_ERR_BAD_ARG; }
        if (sd->nin < 0 || sd->nin > 2 || (sd->nin > 0 && !sd->in[0]) || (sd->nin > 1 && !sd->in[1])) { set_error("spb_flux_div_rk_stage: bad inputs"); return SPB_ERR_BAD_ARG; }
        const FluxParams P = make_params(f);
        StageParams S{};
        S.nin = sd->nin; S.has_out = sd->out ? 1 : 0;
        for (int a = 0; a < 2; ++a) { S.in[a] = a < sd->nin ? sd->in[a] : nullptr; S.cq[a] = a < sd->nin ? sd->cq[a] : 0.0; S.co[a] = a < sd->nin ? sd->co[a] : 0.0; }
        S.cq_self = sd->cq_self; S.co_self = sd->co_self;
        S.gm1 = f->gamma - 1.0; S.inv_gm1 = 1.0/(f->gamma - 1.0); S.inv_R = 1.0/f->R;
        cudaStream_t st = (cudaStream_t)stream;
        if (f->sgs != SPB_SGS_NONE && f->sgs != SPB_SGS_WALE) { set_error("spb_flux_div_rk_stage: unknown SGS model"); return SPB_ERR_BAD_ARG; }
        if (g->metric_dev && f->sgs == SPB_SGS_NONE)
        {
            int rc = SPB_ERR_UNSUPPORTED;
#define SPB_NARROW(C, V) if (f->conv == C && f->diss == SPB_DISS_NONE && (f->visc != 0) == (V != 0)) \
                rc = launch_fdiv_narrow<C, V>(g, q_in, sd->out, P, 0, lb_begin, lb_end, st, q_out, &S, exch)
            SPB_NARROW(SPB_CONV_TOTANI, 1);
            SPB_NARROW(SPB_CONV_TOTANI, 0);
            SPB_NARROW(SPB_CONV_NONE,   1);
#undef SPB_NARROW
            if (rc != SPB_ERR_UNSUPPORTED) return rc;
        }
        // wide stencils, the WALE closure and general coordinates on the wide kernel: the same-rank ghost cells are stored by the
        // threads that own the source cells (no separate spb_exchange_local), for any number of exchange cells
        if (f->sgs == SPB_SGS_WALE) return flux_div_sgs(g, q_in, sd->out, f, P, 0, lb_begin, lb_end, st, q_out, &S, exch);
        if (g->metric_dev) return flux_div_curv(g, q_in, sd->out, f, P, 0, lb_begin, lb_end, st, q_out, &S, exch);
#define SPB_NARROW(C, V) if (f->conv == C && f->diss == SPB_DISS_NONE && (f->visc != 0) == (V != 0)) \
            return launch_fdiv_narrow<C, V>(g, q_in, sd->out, P, 0, lb_begin, lb_end, st, q_out, &S, exch)
        SPB_NARROW(SPB_CONV_TOTANI, 1);
        SPB_NARROW(SPB_CONV_TOTANI, 0);
        SPB_NARROW(SPB_CONV_NONE,   1);
#undef SPB_NARROW
        // wide stencils: the stage update rides on the rhs kernel; with a plan the owning threads also store the same-rank ghosts
#define SPB_WIDE(C, D) if (f->conv == C && f->diss == D && f->visc != 0) \
            return launch_fdiv<C, D, 1, true>(g, q_in, sd->out, P, 0, lb_begin, lb_end, st, q_out, &S, exch)
        SPB_WIDE(SPB_CONV_TOTANI,     SPB_DISS_FWENO);
        SPB_WIDE(SPB_CONV_CENT_KEEP4, SPB_DISS_NONE);
        SPB_WIDE(SPB_CONV_CENT_KEEP4, SPB_DISS_FWENO);
        SPB_WIDE(SPB_CONV_CENT_KEEP6, SPB_DISS_NONE);
        SPB_WIDE(SPB_CONV_CENT_KEEP8, SPB_DISS_NONE);
#undef SPB_WIDE
        set_error("spb_flux_div_rk_stage: the fused stage is implemented for totani_lr and/or visc_lr, and for the hybrid / cent_keep<4> schemes with visc_lr");
        return SPB_ERR_UNSUPPORTED;
    }

    int spb_flux_div_rk_stage_part(const spb_grid* g, const double* q_in, double* q_out, const spb_flux_desc* f,
                                   const spb_stage_desc* sd, spb_exchange* exch, int fuse_ghosts, int part, void* stream)
    {
        using namespace spb;
        if (!g) { set_error("spb_flux_div_rk_stage_part: null grid"); return SPB_ERR_BAD_ARG; }
        if (part == SPB_PART_ALL) return spb_flux_div_rk_stage_exchange(g, q_in, q_out, f, sd, fuse_ghosts ? exch : nullptr, 0, g->nlb, stream);
        if (!exch) { set_error("spb_flux_div_rk_stage_part: a part needs the exchange plan"); return SPB_ERR_BAD_ARG; }
        const int* list = nullptr; int64_t count = 0;
        int rc = exchange_block_list(exch, g->nlb, part, &list, &count); if (rc) return rc;
        if (count == 0) return 0;
        BlockList& bl = current_block_list();
        bl.dev = list; bl.count = count;
        rc = spb_flux_div_rk_stage_exchange(g, q_in, q_out, f, sd, fuse_ghosts ? exch : nullptr, 0, g->nlb, stream);
        bl.dev = nullptr; bl.count = 0;
        return rc;
    }

    int spb_flux_div(const spb_grid* g, const double* q_dev, double* rhs_dev, const spb_flux_desc* f, int increment, void* stream)
    {
        if (!g) { spb::set_error("spb_flux_div: null grid"); return SPB_ERR_BAD_ARG; }
        return spb_flux_div_blocks(g, q_dev, rhs_dev, f, increment, 0, g->nlb, stream);
    }
}
