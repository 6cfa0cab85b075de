// Face-flux device functions of the RHS kernel: the closed set of SPADE flux functors, written
// for one face of compile-time direction D with every metric term of coords::identity folded away.
//
// Reference semantics (file:line in /root/reference/src):
//   totani_lr                 navier-stokes/convective.h:68-93
//   cent_keep<4>              navier-stokes/convective.h:118-184, core/finite_diff.h:26-34
//   fweno_t (enable_smooth)   navier-stokes/convective.h:355-496
//   hybrid_scheme_t           navier-stokes/hybrid_scheme.h:29-45
//   ducros_t                  navier-stokes/state_sensor.h:32-42
//   visc_lr + constant_viscosity_t   navier-stokes/viscous.h:39-80, viscous_laws.h:77-99
//   face value / face gradient       omni/infos/info_value.h:30-40, info_gradient.h:22-87
// The arithmetic is re-associated (each face once, unit normals folded, e = R T/(gamma-1) instead
// of p/(rho (gamma-1)), only the stress row the face needs) — results agree with the reference to
// round-off (gate: 1e-12 relative L2, tests/test_flux_div_gpu.py), not bit-for-bit.
#pragma once
#include "spb_common.cuh"

namespace spb
{
    struct FluxParams
    {
        double gamma, R, gm1, cv, inv_gm1;     // cv = R/(gamma-1)
        double mu, beta, two_mu, kappa;
        double eps;
        int    blend;                 // SPB_BLEND_*
        int    weno_linear;           // weno_smooth_indicator disable_smooth: the linear weights 1/3, 2/3, 2/3, 1/3 (convective.h:301-305,397-401)
        double sgs_c, sgs_cp_prt;     // wale_t: cw^2 delta^2; (gamma R/(gamma-1))/Pr_t
    };

    // Fused RK stage (spb_flux_div_rk_stage): with r = rhs(q_in) of a cell,
    //   q_out = prim(cons(q_in) + cq_self r + cq[0] in[0] + cq[1] in[1])      (advance.h:57-102, fluid_state.h:103-135)
    //   out   = co_self r + co[0] in[0] + co[1] in[1]                         (residual register for later stages)
    struct StageParams
    {
        int nin, has_out;
        const double* in[2];
        double cq_self, cq[2];
        double co_self, co[2];
        double gm1, inv_gm1, inv_R;
    };

    // 1/a from a MUFU.RCP64H seed (the upper 20 mantissa bits of a: relative error e0 <~ 1e-6) with ONE third-order step
    // x (1 + e + e^2), e = 1 - a x: the error becomes e0^3 ~ 1e-18, below the rounding of the last fma (round 1 / 2 used two
    // Newton steps, one fma more). No slow-path call (div.rn.f64 is ~3x the instructions and keeps a subroutine alive that costs
    // registers). Arguments here are densities, sound speeds and smoothness sums: positive and far from the denormal range.
    __device__ __forceinline__ double rcp_nr(const double a)
    {
        double x;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
        const double e = fma(-a, x, 1.0);
        return fma(x, fma(e, e, e), x);
    }

    // sqrt(a) for a > 0 from a MUFU.RSQ64H seed: one Newton step on y = 1/sqrt(a) (error ~1e-12), then the correction
    // s + (y/2)(a - s^2) of the root s = a y, which squares the error again and only needs y to 1e-12 (relative error of the
    // result ~1e-16, no slow-path call).
    __device__ __forceinline__ double sqrt_nr(const double a)
    {
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
        const double h = 0.5*a;
        const double t = fma(-h*y, y, 0.5);
        y = fma(y, t, y);
        const double s = a*y;
        const double r = fma(-s, s, a);
        return fma(0.5*y, r, s);
    }

    // q_v at offset (sD along D, sT1 along (D+1)%3, sT2 along (D+2)%3) from the face's right cell
    template <int D, class A>
    __device__ __forceinline__ double qrel(const A& a, int v, int sD, int sT1, int sT2)
    {
        constexpr int T1 = (D + 1) % 3, T2 = (D + 2) % 3;
        int o[3];
        o[D] = sD; o[T1] = sT1; o[T2] = sT2;
        return a(v, o[0], o[1], o[2]);
    }

    // General (diagonal) coordinates, reference src/core/coord_system.h:250-267,295-302: what one face of direction D needs.
    //   A      face-area factor: the D component of calc_normal_vector = (m0 m1 m2)/m_D = m_T1 m_T2 (the reference forms it
    //          per stencil cell as 1/(m_D * (1/(m0 m1 m2))); the tangential derivatives are the same for every cell of the
    //          stencil of a face, so the cells agree to round-off)
    //   gs[d]  gradient scale of direction d: inv_dx_d / m_d (normal: m_D at the face, tangential: m_t at the cell row)
    // Identity coordinates: A = 1, gs = inv_dx (the CURV = false instantiations never read A).
    struct FaceMetric
    {
        double A;
        double gs[3];
    };

    // ---- convective::totani_lr -------------------------------------------------------------
    template <int D>
    __device__ __forceinline__ void flux_totani(const FluxParams& P, const double (&qL)[5], const double (&qR)[5], double (&F)[5],
                                                const double rhoL, const double rhoR)
    {
        const double unL = qL[2+D], unR = qR[2+D];
        const double C = (rhoL + rhoR)*(unL + unR);                 // 4c
        double S = P.cv*(qL[1] + qR[1]);                            // e_L + e_R
        S = fma(qL[2], qR[2], S); S = fma(qL[3], qR[3], S); S = fma(qL[4], qR[4], S);
        const double C8 = 0.125*C;
        F[0] = 0.25*C;
        F[1] = fma(C8, S, 0.5*fma(unL, qR[0], unR*qL[0]));
        F[2] = C8*(qL[2] + qR[2]);
        F[3] = C8*(qL[3] + qR[3]);
        F[4] = C8*(qL[4] + qR[4]);
        F[2+D] = fma(0.5, qL[0] + qR[0], F[2+D]);
    }

    // ---- convective::cent_keep<4>: cells c0..c3 at half-offsets -3,-1,+1,+3 -------------------
    template <int D>
    __device__ __forceinline__ void flux_cent_keep4(const FluxParams& P, const double (&c0)[5], const double (&c1)[5],
                                                    const double (&c2)[5], const double (&c3)[5], double (&F)[5])
    {
        const double a1 = 2.0/3.0, a2 = -1.0/12.0;
        const double* q[4] = {c0, c1, c2, c3};
        double rho[4], eng[4];
        #pragma unroll
        for (int i = 0; i < 4; ++i) { rho[i] = q[i][0]*rcp_nr(P.R*q[i][1]); eng[i] = P.cv*q[i][1]; }
        double cc = 0.0, m[3] = {0.0, 0.0, 0.0}, g = 0.0, k = 0.0, ie = 0.0, pd = 0.0;
        auto pair = [&](const double a, const int i0, const int i1)
        {
            const double c_loc = a*0.25*(rho[i0] + rho[i1])*(q[i0][2+D] + q[i1][2+D]);
            cc += c_loc;
            #pragma unroll
            for (int d = 0; d < 3; ++d) m[d] = fma(c_loc, 0.5*(q[i0][2+d] + q[i1][2+d]), m[d]);
            g  = fma(a, 0.5*(q[i0][0] + q[i1][0]), g);
            k  = fma(0.5*c_loc, fma(q[i0][2], q[i1][2], fma(q[i0][3], q[i1][3], q[i0][4]*q[i1][4])), k);
            ie = fma(c_loc, 0.5*(eng[i0] + eng[i1]), ie);
            pd = fma(a, 0.5*fma(q[i0][2+D], q[i1][0], q[i1][2+D]*q[i0][0]), pd);
        };
        pair(a1, 1, 2);
        pair(a2, 1, 3);
        pair(a2, 0, 2);
        F[0] = 2.0*cc;
        F[1] = 2.0*(k + ie + pd);
        F[2] = 2.0*m[0]; F[3] = 2.0*m[1]; F[4] = 2.0*m[2];
        F[2+D] = fma(2.0, g, F[2+D]);
    }

    // ---- convective::cent_keep<ORDER>, ORDER = 6, 8 (convective.h:118-184): cells c[0 .. ORDER-1] at half-offsets
    // -(ORDER-1), ..., +(ORDER-1); coefficients core/finite_diff.h:26-34
    template <int IDX, int ORDER> struct cfd_coeff
    {
        static constexpr double fact(int n) { double f = 1.0; for (int i = 2; i <= n; ++i) f *= i; return f; }
        static constexpr int h = ORDER/2;
        static constexpr double value = (fact(h)/fact(h + IDX))*(fact(h)/fact(h - IDX))*(((1 + IDX) % 2 == 0) ? 1.0 : -1.0)/double(IDX);
    };
    template <int D, int ORDER>
    __device__ __forceinline__ void flux_cent_keep_wide(const FluxParams& P, const double (&q)[ORDER][5], double (&F)[5])
    {
        constexpr int HW = ORDER/2;
        double rho[ORDER], eng[ORDER];
        #pragma unroll
        for (int i = 0; i < ORDER; ++i) { rho[i] = q[i][0]*rcp_nr(P.R*q[i][1]); eng[i] = P.cv*q[i][1]; }
        double cc = 0.0, m[3] = {0.0, 0.0, 0.0}, g = 0.0, k = 0.0, ie = 0.0, pd = 0.0;
        auto pair = [&](const double a, const int i0, const int i1)
        {
            const double c_loc = a*0.25*(rho[i0] + rho[i1])*(q[i0][2+D] + q[i1][2+D]);
            cc += c_loc;
            #pragma unroll
            for (int d = 0; d < 3; ++d) m[d] = fma(c_loc, 0.5*(q[i0][2+d] + q[i1][2+d]), m[d]);
            g  = fma(a, 0.5*(q[i0][0] + q[i1][0]), g);
            k  = fma(0.5*c_loc, fma(q[i0][2], q[i1][2], fma(q[i0][3], q[i1][3], q[i0][4]*q[i1][4])), k);
            ie = fma(c_loc, 0.5*(eng[i0] + eng[i1]), ie);
            pd = fma(a, 0.5*fma(q[i0][2+D], q[i1][0], q[i1][2+D]*q[i0][0]), pd);
        };
        // the reference's order: ii = 1 .. HW, jj = 0 .. ii-1, i0 = HW-1-jj, i1 = i0 + ii
        const double coef[5] = {0.0, cfd_coeff<1, ORDER>::value, cfd_coeff<2, ORDER>::value, cfd_coeff<3, ORDER>::value,
                                ORDER >= 8 ? cfd_coeff<(ORDER >= 8 ? 4 : 1), ORDER>::value : 0.0};
        #pragma unroll
        for (int ii = 1; ii <= HW; ++ii)
            #pragma unroll
            for (int jj = 0; jj < ii; ++jj) pair(coef[ii], HW - 1 - jj, HW - 1 - jj + ii);
        F[0] = 2.0*cc;
        F[1] = 2.0*(k + ie + pd);
        F[2] = 2.0*m[0]; F[3] = 2.0*m[1]; F[4] = 2.0*m[2];
        F[2+D] = fma(2.0, g, F[2+D]);
    }

    // ---- convective::fweno_t ---------------------------------------------------------------------
    // Two 2-cell candidate reconstructions per side with the weights a1/(2a0 + a1) (convective.h:380-400). In the reference's
    // form r1 + w0 (r0 - r1) + r2 + w3 (r3 - r2) the candidates satisfy r0 - r1 = (b1 - b0)/2, r3 - r2 = (b3 - b2)/2 with the
    // differences b the smoothness indicators are built from, and r1 + r2 = f1 + f2: 34 fp64 instructions instead of 47.
    __device__ __forceinline__ double fweno_apply(const double (&f)[4], const double (&d)[4], const int linear)
    {
        const double f0u = f[0] + d[0];
        const double f1u = f[1] + d[1], f1d = f[1] - d[1];
        const double f2u = f[2] + d[2], f2d = f[2] - d[2];
        const double f3d = f[3] - d[3];
        const double b0 = f0u - f1u, b1 = f1u - f2u, b2 = f1d - f2d, b3 = f2d - f3d;
        const double eps = 1e-16;
        double a0 = fma(b0, b0, eps), a1 = fma(b1, b1, eps), a2 = fma(b2, b2, eps), a3 = fma(b3, b3, eps);
        a0 *= a0; a1 *= a1; a2 *= a2; a3 *= a3;
        double w0 = a1*rcp_nr(fma(2.0, a0, a1));
        double w3 = a2*rcp_nr(fma(2.0, a3, a2));
        if (linear) { w0 = 1.0/3.0; w3 = 1.0/3.0; }      // disable_smooth: the linear weights (two selects, no control flow)
        return fma(0.5, fma(w3, b3 - b2, w0*(b1 - b0)), f[1] + f[2]);
    }

    // The direction-independent per-cell part of fweno_t (convective.h:355-378): hr = rho/2 and hs = (rho/2)(|u| + c). One
    // reciprocal and two square roots per CELL: the wide kernel evaluates it once per cell and plane and hands it to the
    // faces (flux_fweno_pre) instead of once per stencil cell of every face (12 times per cell).
    __device__ __forceinline__ void weno_cell(const FluxParams& P, const double p, const double T, const double u, const double v,
                                              const double w, double& hr, double& hs)
    {
        const double a = P.R*T;
        const double u2 = fma(u, u, fma(v, v, w*w));
        hr = 0.5*(p*rcp_nr(a));
        hs = hr*(sqrt_nr(fmax(u2, 1e-300)) + sqrt_nr(a*P.gamma));      // |u| = 0 becomes 1e-150
    }

    // fweno_t on prepared cells: the same arithmetic as flux_fweno below with rho/2 and the half spectral radius given
    template <int D, bool CURV = false>
    __device__ __forceinline__ void flux_fweno_pre(const FluxParams& P, const double (&c0)[5], const double (&c1)[5],
                                                   const double (&c2)[5], const double (&c3)[5], const double (&hr)[4],
                                                   const double (&hsr)[4], double (&F)[5], const double A = 1.0)
    {
        const double* q[4] = {c0, c1, c2, c3};
        double a[4], ke[4], fm[4], fl[4], ds[4];
        #pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            a[i]  = P.R*q[i][1];
            ke[i] = 0.5*fma(q[i][2], q[i][2], fma(q[i][3], q[i][3], q[i][4]*q[i][4]));
            fm[i] = hr[i]*q[i][2+D];
            if (CURV) fm[i] *= A;
        }
        const double hA = CURV ? 0.5*A : 0.5;
        F[0] = fweno_apply(fm, hsr, P.weno_linear);
        #pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            const double engy = fma(a[i], P.inv_gm1, ke[i]);
            fl[i] = fm[i]*(engy + a[i]);
            ds[i] = hsr[i]*engy;
        }
        F[1] = fweno_apply(fl, ds, P.weno_linear);
        #pragma unroll
        for (int dr = 0; dr < 3; ++dr)
        {
            #pragma unroll
            for (int i = 0; i < 4; ++i)
            {
                fl[i] = fm[i]*q[i][2+dr];
                if (dr == D) fl[i] = fma(hA, q[i][0], fl[i]);
                ds[i] = hsr[i]*q[i][2+dr];
            }
            F[2+dr] = fweno_apply(fl, ds, P.weno_linear);
        }
    }

    // general coordinates: the metric scales the flux part (u.n, p n) but not the Rusanov dissipation, exactly like the
    // reference (convective.h:363-378 forms hlf_sig_rho without the metric)
    template <int D, bool CURV = false>
    __device__ __forceinline__ void flux_fweno(const FluxParams& P, const double (&c0)[5], const double (&c1)[5],
                                               const double (&c2)[5], const double (&c3)[5], double (&F)[5], const double A = 1.0)
    {
        const double* q[4] = {c0, c1, c2, c3};
        double rho[4], hsr[4], a[4], ke[4], fm[4], fl[4], ds[4];
        #pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            a[i]   = P.R*q[i][1];
            const double u2 = fma(q[i][2], q[i][2], fma(q[i][3], q[i][3], q[i][4]*q[i][4]));
            ke[i]  = 0.5*u2;
            rho[i] = q[i][0]*rcp_nr(a[i]);
            hsr[i] = 0.5*rho[i]*(sqrt_nr(fmax(u2, 1e-300)) + sqrt_nr(a[i]*P.gamma));      // |u| = 0 becomes 1e-150
            fm[i]  = 0.5*rho[i]*q[i][2+D];
            if (CURV) fm[i] *= A;
        }
        const double hA = CURV ? 0.5*A : 0.5;
        // continuity
        F[0] = fweno_apply(fm, hsr, P.weno_linear);
        // energy
        #pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            const double engy = fma(a[i], P.inv_gm1, ke[i]);
            fl[i] = fm[i]*(engy + a[i]);
            ds[i] = hsr[i]*engy;
        }
        F[1] = fweno_apply(fl, ds, P.weno_linear);
        // momentum
        #pragma unroll
        for (int dr = 0; dr < 3; ++dr)
        {
            #pragma unroll
            for (int i = 0; i < 4; ++i)
            {
                fl[i] = fm[i]*q[i][2+dr];
                if (dr == D) fl[i] = fma(hA, q[i][0], fl[i]);
                ds[i] = hsr[i]*q[i][2+dr];
            }
            F[2+dr] = fweno_apply(fl, ds, P.weno_linear);
        }
    }

    // ---- the composed functor on the lower face (direction D) of the cell the accessor is centred on
    // ---- subgrid_scale::wale_t::get_mu_t (subgrid_scale.h:45-90) from the face gradient g[dir][comp] and the face density
    __device__ __forceinline__ double wale_mu_t(const FluxParams& P, const double rho, const double (&g)[3][3])
    {
        // gij(i,j) = grad[j].u(i) = g[j][i];  gij2 = gij gij;  sd = sym(gij2) - tr(gij2)/3 I;  s = sym(gij)
        double g2[3][3];
        #pragma unroll
        for (int i = 0; i < 3; ++i)
            #pragma unroll
            for (int j = 0; j < 3; ++j) g2[i][j] = fma(g[0][i], g[j][0], fma(g[1][i], g[j][1], g[2][i]*g[j][2]));
        const double tr3 = 0.33333333333333333*(g2[0][0] + g2[1][1] + g2[2][2]);
        double ssd = 0.0, ss = 0.0;
        #pragma unroll
        for (int i = 0; i < 3; ++i)
            #pragma unroll
            for (int j = 0; j < 3; ++j)
            {
                double sd = 0.5*(g2[i][j] + g2[j][i]);
                if (i == j) sd -= tr3;
                ssd = fma(sd, sd, ssd);
                const double sij = 0.5*(g[j][i] + g[i][j]);
                ss = fma(sij, sij, ss);
            }
        const double sqrt0 = sqrt_nr(fmax(ssd, 1e-300));
        const double sqrt1 = sqrt_nr(sqrt0);
        const double sqrt2 = sqrt_nr(fmax(ss, 1e-300));
        return rho*P.sgs_c*ssd*sqrt0*rcp_nr(1e-8 + fma(ss*ss, sqrt2, ssd*sqrt1));
    }

    // PRE: hr4 / hs4 hold weno_cell of the four stencil cells (LL, L, R, RR) of the face
    template <int CONV, int DISS, int VISC, int D, bool CURV = false, bool SGS = false, bool PRE = false, class A>
    __device__ __forceinline__ void face_flux(const A& a, const FluxParams& P, const double (&invdx)[3], double (&F)[5],
                                              const double area = 1.0, const double* hr4 = nullptr, const double* hs4 = nullptr)
    {
        constexpr int T1 = (D + 1) % 3, T2 = (D + 2) % 3;
        constexpr bool WIDE = (CONV == SPB_CONV_CENT_KEEP4) || (CONV == SPB_CONV_FWENO) || (DISS != SPB_DISS_NONE);
        double qL[5], qR[5], qLL[5], qRR[5];
        #pragma unroll
        for (int v = 0; v < 5; ++v) { qL[v] = qrel<D>(a, v, -1, 0, 0); qR[v] = qrel<D>(a, v, 0, 0, 0); }
        if (WIDE)
        {
            #pragma unroll
            for (int v = 0; v < 5; ++v) { qLL[v] = qrel<D>(a, v, -2, 0, 0); qRR[v] = qrel<D>(a, v, 1, 0, 0); }
        }
        #pragma unroll
        for (int v = 0; v < 5; ++v) F[v] = 0.0;

        double hrv[4] = {0.0, 0.0, 0.0, 0.0}, hsv[4] = {0.0, 0.0, 0.0, 0.0};
        if (PRE)
        {
            #pragma unroll
            for (int i = 0; i < 4; ++i) { hrv[i] = hr4[i]; hsv[i] = hs4[i]; }
        }
        if (CONV == SPB_CONV_TOTANI)
        {
            // PRE: the prepared cells carry rho/2 = (p rcp(R T))/2, so 2 hr is the density bit for bit
            if (PRE) flux_totani<D>(P, qL, qR, F, 2.0*hrv[1], 2.0*hrv[2]);
            else     flux_totani<D>(P, qL, qR, F, qL[0]*rcp_nr(P.R*qL[1]), qR[0]*rcp_nr(P.R*qR[1]));
        }
        if (CONV == SPB_CONV_CENT_KEEP4) flux_cent_keep4<D>(P, qLL, qL, qR, qRR, F);
        if (CONV == SPB_CONV_FWENO)
        {
            if (PRE) flux_fweno_pre<D, CURV>(P, qLL, qL, qR, qRR, hrv, hsv, F, area);
            else     flux_fweno<D, CURV>(P, qLL, qL, qR, qRR, F, area);
        }
        if (CONV == SPB_CONV_CENT_KEEP6 || CONV == SPB_CONV_CENT_KEEP8)
        {
            constexpr int ORDER = (CONV == SPB_CONV_CENT_KEEP6) ? 6 : 8;
            double qw[ORDER][5];
            #pragma unroll
            for (int s = 0; s < ORDER; ++s)
                #pragma unroll
                for (int v = 0; v < 5; ++v) qw[s][v] = qrel<D>(a, v, s - ORDER/2, 0, 0);
            flux_cent_keep_wide<D, ORDER>(P, qw, F);
        }
        if (CURV && (CONV == SPB_CONV_TOTANI || CONV == SPB_CONV_CENT_KEEP4 || CONV == SPB_CONV_CENT_KEEP6 || CONV == SPB_CONV_CENT_KEEP8))
        {
            // both are linear in the metric vector (convective.h:76-91, 128-182)
            #pragma unroll
            for (int v = 0; v < 5; ++v) F[v] *= area;
        }

        if (VISC || DISS)
        {
            // face gradient g[dir][comp] of the velocity; gT = normal derivative of T
            double g[3][3];
            #pragma unroll
            for (int c = 0; c < 3; ++c) g[D][c] = (qR[2+c] - qL[2+c])*invdx[D];
            const double gT = (qR[1] - qL[1])*invdx[D];
            const double c1 = 0.25*invdx[T1], c2 = 0.25*invdx[T2];
            #pragma unroll
            for (int c = 0; c < 3; ++c)
            {
                const bool need1 = DISS || SGS || (c == T1) || (c == D);
                const bool need2 = DISS || SGS || (c == T2) || (c == D);
                g[T1][c] = 0.0; g[T2][c] = 0.0;
                if (need1)
                    g[T1][c] = c1*((qrel<D>(a, 2+c, -1,  1, 0) - qrel<D>(a, 2+c, -1, -1, 0))
                                 + (qrel<D>(a, 2+c,  0,  1, 0) - qrel<D>(a, 2+c,  0, -1, 0)));
                if (need2)
                    g[T2][c] = c2*((qrel<D>(a, 2+c, -1, 0,  1) - qrel<D>(a, 2+c, -1, 0, -1))
                                 + (qrel<D>(a, 2+c,  0, 0,  1) - qrel<D>(a, 2+c,  0, 0, -1)));
            }
            const double div = g[0][0] + g[1][1] + g[2][2];
            if (DISS)
            {
                const double th2 = div*div;
                const double w0 = g[1][2] - g[2][1];
                const double w1 = g[2][0] - g[0][2];
                const double w2 = g[0][1] - g[1][0];
                const double vort = fma(w0, w0, fma(w1, w1, w2*w2));
                const double alpha = th2*rcp_nr(th2 + vort + P.eps);
                double F1[5];
                if (PRE) flux_fweno_pre<D, CURV>(P, qLL, qL, qR, qRR, hrv, hsv, F1, area);
                else     flux_fweno<D, CURV>(P, qLL, qL, qR, qRR, F1, area);
                const double coeff0 = (P.blend == SPB_BLEND_FULL_FLUX) ? (1.0 - alpha) : 1.0;
                #pragma unroll
                for (int v = 0; v < 5; ++v) F[v] = fma(alpha, F1[v], coeff0*F[v]);
            }
            if (VISC)
            {
                double mu = P.mu, two_mu = P.two_mu, beta = P.beta, kappa = P.kappa;
                if (SGS)
                {
                    // viscous_laws::sgs_visc_t::get_all (viscous_laws.h:185-196) with the face value of the density
                    const double rho_f = 0.5*(qL[0] + qR[0])*rcp_nr(P.R*(0.5*(qL[1] + qR[1])));
                    const double mu_t = wale_mu_t(P, rho_f, g);
                    mu += mu_t; two_mu = 2.0*mu;
                    beta = fma(-0.66666666667, mu_t, beta);
                    kappa = fma(P.sgs_cp_prt, mu_t, kappa);
                }
                double tDD = fma(two_mu, g[D][D], beta*div);
                double tD1 = mu*(g[D][T1] + g[T1][D]);
                double tD2 = mu*(g[D][T2] + g[T2][D]);
                const double ufD = 0.5*(qL[2+D]  + qR[2+D]);
                const double uf1 = 0.5*(qL[2+T1] + qR[2+T1]);
                const double uf2 = 0.5*(qL[2+T2] + qR[2+T2]);
                double h = fma(ufD, tDD, fma(uf1, tD1, fma(uf2, tD2, kappa*gT)));
                if (CURV) { h *= area; tDD *= area; tD1 *= area; tD2 *= area; }      // -(n . tau), n = area e_D (viscous.h:70-74)
                F[1]    -= h;
                F[2+D]  -= tDD;
                F[2+T1] -= tD1;
                F[2+T2] -= tD2;
            }
        }
    }

    template <int CONV, int DISS> struct stencil_halo
    {
        static constexpr int value = (CONV == SPB_CONV_CENT_KEEP8) ? 4 : (CONV == SPB_CONV_CENT_KEEP6) ? 3
            : ((CONV == SPB_CONV_CENT_KEEP4) || (CONV == SPB_CONV_FWENO) || (DISS != SPB_DISS_NONE)) ? 2 : 1;
    };
}
