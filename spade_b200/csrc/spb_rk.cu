// RK stage updates on interior cells, fused with the prim <-> cons conversion.
// Replaces detail::transform_advance_to and detail::opt_rk3_s0/s1/s2
// (reference src/time-integration/advance.h:57-102, 286-354) and fluid_state::convert_state
// (reference src/navier-stokes/fluid_state.h:103-135); the operation order of the reference is kept.
#include "spb_common.cuh"

namespace spb
{
    struct CellDims
    {
        int nx[3], ng[3], np[3];
        long long ncells;        // interior cells over all local blocks
    };

    __device__ __forceinline__ long long cell_offset(const CellDims& G, long long cell)
    {
        const int i = (int)(cell % G.nx[0]); long long t = cell / G.nx[0];
        const int j = (int)(t % G.nx[1]); t /= G.nx[1];
        const int k = (int)(t % G.nx[2]); const long long lb = t / G.nx[2];
        return 5ll*((i + G.ng[0]) + (long long)G.np[0]*((j + G.ng[1]) + (long long)G.np[1]*((k + G.ng[2]) + (long long)G.np[2]*lb)));
    }

    // reference fluid_state.h:103-116
    __device__ __forceinline__ void prim2cons(const double gamma, const double R, const double (&p)[5], double (&w)[5])
    {
        const double rho   = p[0]/(R*p[1]);
        const double rhoU2 = rho*(p[2]*p[2] + p[3]*p[3] + p[4]*p[4]);
        const double rhoE  = 0.5*rhoU2 + (p[0]/(gamma - 1.0));
        w[0] = rho; w[1] = rhoE; w[2] = rho*p[2]; w[3] = rho*p[3]; w[4] = rho*p[4];
    }
    // reference fluid_state.h:119-135
    __device__ __forceinline__ void cons2prim(const double gamma, const double R, const double (&w)[5], double (&p)[5])
    {
        const double rho = w[0];
        const double invrho = 1.0/rho;
        const double u = invrho*w[2], v = invrho*w[3], ww = invrho*w[4];
        const double rhoU2 = rho*(u*u + v*v + ww*ww);
        const double pr = (gamma - 1.0)*(w[1] - 0.5*rhoU2);
        p[0] = pr; p[1] = pr/(R*rho); p[2] = u; p[3] = v; p[4] = ww;
    }

    struct RkArgs { const double* k[4]; double coeff[4]; int nk; };

    __global__ void __launch_bounds__(256) rk_update_kernel(double* __restrict__ q, const RkArgs A, const double gamma, const double R, const CellDims G)
    {
        const long long stride = (long long)gridDim.x*blockDim.x;
        for (long long cell = (long long)blockIdx.x*blockDim.x + threadIdx.x; cell < G.ncells; cell += stride)
        {
            const long long o = cell_offset(G, cell);
            double p[5], w[5];
            #pragma unroll
            for (int v = 0; v < 5; ++v) p[v] = q[o + v];
            prim2cons(gamma, R, p, w);
            #pragma unroll
            for (int s = 0; s < 4; ++s)
            {
                if (s < A.nk && A.coeff[s] != 0.0)
                {
                    const double* __restrict__ ks = A.k[s];
                    #pragma unroll
                    for (int v = 0; v < 5; ++v) w[v] += A.coeff[s]*ks[o + v];
                }
            }
            cons2prim(gamma, R, w, p);
            #pragma unroll
            for (int v = 0; v < 5; ++v) q[o + v] = p[v];
        }
    }

    template <int STAGE>
    __global__ void __launch_bounds__(256) ssprk3_kernel(double* __restrict__ q, double* __restrict__ r0, const double* __restrict__ r1,
                                                         const double dt, const double gamma, const double R, const CellDims G)
    {
        const long long stride = (long long)gridDim.x*blockDim.x;
        for (long long cell = (long long)blockIdx.x*blockDim.x + threadIdx.x; cell < G.ncells; cell += stride)
        {
            const long long o = cell_offset(G, cell);
            double p[5], w[5];
            #pragma unroll
            for (int v = 0; v < 5; ++v) p[v] = q[o + v];
            prim2cons(gamma, R, p, w);
            #pragma unroll
            for (int v = 0; v < 5; ++v)
            {
                const double r0i = r0[o + v];
                if (STAGE == 0) { w[v] += dt*r0i; }
                if (STAGE == 1)
                {
                    const double r1i = r1[o + v];
                    const double nr0 = (1.0/6.0)*dt*(r0i + r1i);
                    r0[o + v] = nr0;
                    w[v] += (3.0/2.0)*nr0;
                    w[v] -= dt*r0i;
                }
                if (STAGE == 2)
                {
                    const double r1i = r1[o + v];
                    w[v] -= (1.0/2.0)*r0i;
                    w[v] += dt*(2.0/3.0)*r1i;
                }
            }
            cons2prim(gamma, R, w, p);
            #pragma unroll
            for (int v = 0; v < 5; ++v) q[o + v] = p[v];
        }
    }

    static CellDims make_dims(const spb_grid* g)
    {
        CellDims G;
        for (int d = 0; d < 3; ++d) { G.nx[d] = g->nx[d]; G.ng[d] = g->ng[d]; G.np[d] = g->np[d]; }
        G.ncells = (long long)g->nx[0]*g->nx[1]*g->nx[2]*g->nlb;
        return G;
    }

    static unsigned grid_for(const spb_grid* g, long long n, int threads, int per_sm)
    {
        long long blocks = (n + threads - 1)/threads;
        const long long cap = (long long)g->num_sms*per_sm;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        return (unsigned)blocks;
    }
}

namespace spb
{
    // generic integrate_advance (advance.h:149-153, 190-194, 216-221): the reference's three passes
    //   resid *= c (every element incl. exchange cells, grid_array.h:359-369);  sol +-= resid (interior, grid_array.h:289-321);
    //   resid *= 1.0/c
    // as one pass with the same per-element operations in the same order (no FMA contraction: the rounding of resid*c is
    // part of the reference's result, and resid comes back perturbed by the round trip exactly like there)
    __global__ void __launch_bounds__(256) axpy_roundtrip_kernel(double* __restrict__ sol, double* __restrict__ resid, const double c,
                                                                 const double inv_c, const int subtract, const CellDims G, const long long ncell_all)
    {
        const long long stride = (long long)gridDim.x*blockDim.x;
        for (long long cell = (long long)blockIdx.x*blockDim.x + threadIdx.x; cell < ncell_all; cell += stride)
        {
            const int ip = (int)(cell % G.np[0]); long long t = cell / G.np[0];
            const int jp = (int)(t % G.np[1]); t /= G.np[1];
            const int kp = (int)(t % G.np[2]);
            const bool interior = ip >= G.ng[0] && ip < G.ng[0] + G.nx[0] && jp >= G.ng[1] && jp < G.ng[1] + G.nx[1]
                               && kp >= G.ng[2] && kp < G.ng[2] + G.nx[2];
            #pragma unroll
            for (int v = 0; v < 5; ++v)
            {
                const double r = __dmul_rn(resid[5*cell + v], c);
                if (interior) sol[5*cell + v] = subtract ? __dsub_rn(sol[5*cell + v], r) : __dadd_rn(sol[5*cell + v], r);
                resid[5*cell + v] = __dmul_rn(r, inv_c);
            }
        }
    }
}

extern "C"
{
    // the planner of the fused path, shared by both host sides (see include/spade_b200.h)
    int spb_rk_fused_plan(int n, const double* diffs, spb_stage_plan* plan)
    {
        if (n < 1 || n > 8 || !diffs || !plan) { spb::set_error("spb_rk_fused_plan: bad argument"); return SPB_ERR_BAD_ARG; }
        auto D = [&](int i, int j) { return diffs[i*n + j]; };
        int nfinal = 0;
        for (int j = 0; j + 1 < n; ++j) if (D(n - 1, j) != 0.0) ++nfinal;
        const bool use_c = nfinal > 2;
        for (int i = 0; i < n; ++i)
        {
            spb_stage_plan& st = plan[i];
            st = spb_stage_plan{};
            st.out = -1;
            st.cq_self = D(i, i);
            if (i == n - 1 && use_c) { st.nin = 1; st.in[0] = n - 2; st.cq[0] = 1.0; continue; }      // register n-2 holds C
            for (int j = 0; j < i; ++j)
            {
                const bool prior = D(i, j) != 0.0;
                const bool extra = use_c && i == n - 2 && D(n - 1, j) != 0.0;
                if (!prior && !extra) continue;
                if (st.nin == 2) { spb::set_error("spb_rk_fused_plan: a stage needs more than two residual inputs"); return SPB_ERR_UNSUPPORTED; }
                st.in[st.nin] = j; st.cq[st.nin] = D(i, j); st.co[st.nin] = 0.0; ++st.nin;
            }
            if (use_c && i == n - 2)
            {
                st.out = n - 2; st.co_self = D(n - 1, i);
                for (int a = 0; a < st.nin; ++a) st.co[a] = D(n - 1, st.in[a]);
            }
            else
            {
                bool later = false;
                for (int m = i + 1; m < n; ++m) later = later || D(m, i) != 0.0;
                if (later) { st.out = i; st.co_self = 1.0; }
            }
        }
        return 0;
    }

    int spb_axpy_roundtrip(const spb_grid* g, double* sol_dev, double* resid_dev, double c, int subtract, void* stream)
    {
        using namespace spb;
        if (!g || !sol_dev || !resid_dev) { set_error("spb_axpy_roundtrip: null argument"); return SPB_ERR_BAD_ARG; }
        const CellDims G = make_dims(g);
        const long long nall = (long long)g->np[0]*g->np[1]*g->np[2]*g->nlb;
        if (nall == 0) return 0;
        axpy_roundtrip_kernel<<<grid_for(g, nall, 256, 8), 256, 0, (cudaStream_t)stream>>>(sol_dev, resid_dev, c, 1.0/c, subtract, G, nall);
        SPB_LAUNCH_CHECK();
        return 0;
    }

    int spb_rk_update(const spb_grid* g, double* q_dev, const double* const* k_dev, int nk, const double* coeff,
                      double gamma, double R, void* stream)
    {
        using namespace spb;
        if (!g || !q_dev || !k_dev || !coeff || nk < 0 || nk > 4) { set_error("spb_rk_update: bad argument"); return SPB_ERR_BAD_ARG; }
        RkArgs A; A.nk = nk;
        for (int s = 0; s < 4; ++s) { A.k[s] = s < nk ? k_dev[s] : nullptr; A.coeff[s] = s < nk ? coeff[s] : 0.0; }
        const CellDims G = make_dims(g);
        if (G.ncells == 0) return 0;
        rk_update_kernel<<<grid_for(g, G.ncells, 256, 8), 256, 0, (cudaStream_t)stream>>>(q_dev, A, gamma, R, G);
        SPB_LAUNCH_CHECK();
        return 0;
    }

    int spb_ssprk3_stage(const spb_grid* g, int stage, double* q_dev, double* r0_dev, const double* r1_dev,
                         double dt, double gamma, double R, void* stream)
    {
        using namespace spb;
        if (!g || !q_dev || !r0_dev || (stage > 0 && !r1_dev) || stage < 0 || stage > 2) { set_error("spb_ssprk3_stage: bad argument"); return SPB_ERR_BAD_ARG; }
        const CellDims G = make_dims(g);
        if (G.ncells == 0) return 0;
        const unsigned nb = grid_for(g, G.ncells, 256, 8);
        cudaStream_t st = (cudaStream_t)stream;
        if (stage == 0) ssprk3_kernel<0><<<nb, 256, 0, st>>>(q_dev, r0_dev, r1_dev, dt, gamma, R, G);
        if (stage == 1) ssprk3_kernel<1><<<nb, 256, 0, st>>>(q_dev, r0_dev, r1_dev, dt, gamma, R, G);
        if (stage == 2) ssprk3_kernel<2><<<nb, 256, 0, st>>>(q_dev, r0_dev, r1_dev, dt, gamma, R, G);
        SPB_LAUNCH_CHECK();
        return 0;
    }
}
