// Fused RHS kernel for the one-ghost-cell ("narrow") functor set: totani_lr and/or visc_lr.
// Replaces pde_algs::flux_div (reference src/pde-algs/flux-div/flux_div_basic.h:17-77) composed with
// convective::totani_lr (src/navier-stokes/convective.h:68-93) and viscous::visc_lr
// (src/navier-stokes/viscous.h:39-80; face value/gradient src/omni/infos/info_value.h:30-40,
// info_gradient.h:22-87) for coords::identity.
//
// Design (B200, sm_100a):
//  * One CTA = TI x TJ column of one block, marching through k. 8 compute warps (one cell per
//    thread) + 1 edge warp that owns the halo ring (cells i = -1, j = -1) and the faces on the
//    upper tile edge, so no compute warp does more than three faces per plane.
//  * Planes of q (reference AoS order) arrive in a 4-slot shared-memory ring through TMA
//    (cp.async.bulk.tensor.4d + mbarrier), two planes ahead of the compute.
//  * A thread keeps its own column (k-1, k, k+1) in registers. Everything a neighbour needs from a
//    cell is published once per plane in SoA shared memory: density and the scaled central
//    differences D_t = 0.25/dx_t (q(c+e_t) - q(c-e_t)), because the reference's tangential face
//    gradient is exactly D_t(L) + D_t(R). Every face flux is evaluated once; x/y fluxes are handed
//    to the neighbour through shared memory, the z flux stays in registers.
//  * The finished rhs plane is staged in shared memory and written with one TMA tensor store
//    (hardware clips ragged tiles to the interior), so global stores are fully coalesced.
#include <cstdlib>
#include "spb_common.cuh"
#include "spb_tma.cuh"
#include "spb_flux.cuh"

// Timing-only experiment switches (`make -C spade_b200/csrc exp EXP=N` builds libspade_b200_exp<N>.so with -DSPB_EXP=N, `tools/gpu_visit.sh exp` times them; the results of
// those builds are WRONG by construction, they bound what a restructuring could gain): bit 0 drops barrier (2), bit 1
// drops the flux hand-off through shared memory, bit 2 drops the published differences.
#ifndef SPB_EXP
#define SPB_EXP 0
#endif
#define SPB_BAR2() do { if (!(SPB_EXP & 1)) __syncthreads(); } while (0)

namespace spb
{
    namespace nrw
    {
        // Tile shapes: 32 x 8 (one warp per row) for blocks wider than 16 cells, 16 x 16 (two rows per warp) for blocks of
        // up to 16 cells along i, which would leave half of every 32-wide row idle. Everything below is written on TI, TJ.
        constexpr int NCOMPUTE = 256;                   // 8 compute warps, one cell column per thread
        constexpr int NCW = NCOMPUTE/32;
        constexpr int NTHREADS = NCOMPUTE + 64;         // + edge warp + ghost warp (fused stage with the same-rank ghost exchange)
        constexpr int NTHREADS_NOGHOST = NCOMPUTE + 32; // + edge warp: 288 threads leave 112 registers per thread at 2 CTAs per SM
        template <int N> struct IC { static constexpr int v = N; };
        constexpr int NP = 3;                           // ring slots: planes k, k+1 resident, k+2 in flight
        constexpr int NPUB = 7;                         // rho, cX, Dy.u, Dz.u, cY, Dz.v, Dx.v
        template <int TI_, int TJ_> struct Lay
        {
            static_assert((TI_*TJ_) % 32 == 0 && TI_ <= 32 && 2*TJ_ <= 32, "whole compute warps, rows within a warp, 2 TJ edge lanes");
            static constexpr int NCOMP = TI_*TJ_;                  // compute threads: one cell column each (256, or 512 for the 32 x 16 tile)
            static constexpr int TI = TI_, TJ = TJ_;
            static constexpr int TIp = TI + 2 + 2;                 // halo + 16-byte TMA start alignment slack
            static constexpr int TJp = TJ + 2;
            static constexpr int PLANE_DOUBLES = TIp*TJp*5;
            static constexpr int PLANE_BYTES = PLANE_DOUBLES*8;
            static constexpr int PLANE_STRIDE = ((PLANE_BYTES + 127)/128*128)/8;
            static constexpr int PW = TI + 2;                      // published arrays: [NPUB][TJ+2][PW]
            static constexpr int PSZ = (TJ + 2)*PW;
            static constexpr int FX_DOUBLES = TJ*(TI + 1)*5;
            static constexpr int FY_DOUBLES = (TJ + 1)*TI*5;
            static constexpr int STAGE_DOUBLES = TJ*TI*5;          // 10 240 B, a multiple of 128
            static constexpr int OFF_P = NP*PLANE_STRIDE;
            static constexpr int OFF_STAGE_K = (OFF_P + NPUB*PSZ + 15)/16*16;        // 128-byte aligned for the TMA stores
            static constexpr int OFF_STAGE_Q = OFF_STAGE_K + STAGE_DOUBLES;
            static constexpr int OFF_FX = OFF_STAGE_Q + STAGE_DOUBLES;
            static constexpr int OFF_FY = OFF_FX + FX_DOUBLES;
            static constexpr int OFF_BAR = OFF_FY + FY_DOUBLES;
            // fused ghost exchange: the same-rank neighbour table of this block (27 ints)
            static constexpr int OFF_NBR = (OFF_BAR + NP + 1 + 15)/16*16;
            static constexpr int SMEM_BYTES = (OFF_NBR + 16)*8 + 128;
        };

        enum { P_RHO = 0, P_CX /* Dy.v + Dz.w */, P_DYU, P_DZU, P_CY /* Dz.w + Dx.u */, P_DZV, P_DXV };

        struct Dims
        {
            int nx[3], ng[3], np[3];
            int tiles_i, tiles_j;
            long long block_stride;
            long long lb0;
            int increment;
            int tma_store;
            int ghost;                      // fused stage: also store the finished q planes into the same-rank neighbours' ghost cells
            double idx[3], cdx[3];          // uniform lattice: 1/dx and 0.25/dx
            double lev_inv[3][16];          // non-uniform lattice (AMR): the distinct 1/dx of each direction (spb_grid::lev_inv)
            double lev_cinv[3][16];         // ... and 0.25/dx
            const int* lev;                 // ... and the packed level indices of every block (spb_grid::lev_dev)
            int lm;                         // general coordinates: row length of the metric tables (spb_grid::metric_lm)
            const int* blist;               // local block of CTA group t (a scattered block set in one launch), or null: lb0 + t
        };

        // One tensor map per neighbour direction e = (ex+1) + 3*(ey+1) + 9*(ez+1): a view of q_out whose extents are exactly
        // the destination ghost box of that direction (get_transaction.h:54-86), so a whole staged tile can be stored at a
        // shifted coordinate and the hardware clips everything that is not a ghost cell of that box.
        struct GhostMaps { CUtensorMap m[27]; };

        using Stage = spb::StageParams;

        // rho = p/(R T) (reference convective.h:70, fluid_state.h:105) with the hardware reciprocal seed (MUFU.RCP64H, >= 20 bits)
        // and one third-order step (spb_flux.cuh: rcp_nr) -> relative error ~1e-16, no slow-path branch.
        __device__ __forceinline__ double density(const double R, const double p, const double T)
        {
            return p*rcp_nr(R*T);
        }

        template <int CONV, int VISC, int D>
        __device__ __forceinline__ void face(const FluxParams& P, const double (&qL)[5], const double (&qR)[5],
                                             const double rhoL, const double rhoR, const double ac, const double b,
                                             const double d, const double invdxD, double (&F)[5])
        {
            constexpr int T1 = (D + 1) % 3, T2 = (D + 2) % 3;
            const double s0 = qL[2] + qR[2], s1 = qL[3] + qR[3], s2 = qL[4] + qR[4];
            const double s[3] = {s0, s1, s2};
            if (CONV == SPB_CONV_TOTANI)
            {
                // reference convective.h:68-93 with n = e_D
                const double C  = (rhoL + rhoR)*s[D];                                   // 4c
                double S = P.cv*(qL[1] + qR[1]);                                        // e_L + e_R
                S = fma(qL[2], qR[2], S); S = fma(qL[3], qR[3], S); S = fma(qL[4], qR[4], S);
                const double C8 = 0.125*C;
                F[0] = 0.25*C;
                F[1] = fma(C8, S, 0.5*fma(qL[2+D], qR[0], qR[2+D]*qL[0]));
                F[2] = C8*s0; F[3] = C8*s1; F[4] = C8*s2;
                F[2+D] = fma(0.5, qL[0] + qR[0], F[2+D]);
            }
            else
            {
                #pragma unroll
                for (int v = 0; v < 5; ++v) F[v] = 0.0;
            }
            if (VISC)
            {
                // reference viscous.h:39-80 with n = e_D: only the stress row D is needed
                const double gD  = (qR[2+D] - qL[2+D])*invdxD;
                const double div = gD + ac;
                const double tDD = fma(P.two_mu, gD, P.beta*div);
                const double tD1 = P.mu*fma(qR[2+T1] - qL[2+T1], invdxD, b);
                const double tD2 = P.mu*fma(qR[2+T2] - qL[2+T2], invdxD, d);
                const double hh  = fma(s[D], tDD, fma(s[T1], tD1, s[T2]*tD2));         // 2 u_f . tau_D
                F[1] = fma(-0.5, hh, F[1]);
                F[1] = fma(-(P.kappa*invdxD), qR[1] - qL[1], F[1]);                               // kappa dT/dx_D
                F[2+D]  -= tDD;
                F[2+T1] -= tD1;
                F[2+T2] -= tD2;
            }
        }

        __device__ __forceinline__ double fast_rcp(const double a) { return rcp_nr(a); }

        __device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc)
        {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
        }
        __device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
        __device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

        // Inverse spacings: for a uniform lattice (all blocks the same dx) they come from the kernel-parameter constant
        // bank and cost no registers. An AMR grid has a handful of distinct spacings per direction (its refinement levels):
        // they sit in the constant bank too, and the block's level indices are made warp-uniform FOR THE COMPILER by a
        // warp reduction (REDUX writes a uniform register), so the six values live in uniform registers instead of twelve
        // vector registers per thread (the per-block table of rounds 1 and 2 cost the kernel 9 % through spills).
        template <bool UNIF> struct Spacing
        {
            double i0, i1, i2, c0, c1, c2;
            __device__ __forceinline__ Spacing(const Dims& G, const double* __restrict__ tab, long long lb)
            {
                if (UNIF) { i0 = G.idx[0]; i1 = G.idx[1]; i2 = G.idx[2]; c0 = G.cdx[0]; c1 = G.cdx[1]; c2 = G.cdx[2]; }
                else
                {
                    const unsigned pk = __reduce_max_sync(0xffffffffu, (unsigned)__ldg(G.lev + lb));
                    i0 = G.lev_inv[0][pk & 15u]; i1 = G.lev_inv[1][(pk >> 8) & 15u]; i2 = G.lev_inv[2][(pk >> 16) & 15u];
                    c0 = G.lev_cinv[0][pk & 15u]; c1 = G.lev_cinv[1][(pk >> 8) & 15u]; c2 = G.lev_cinv[2][(pk >> 16) & 15u];
                }
            }
        };

        // CURV: general (diagonal) coordinates (spb_grid_set_metric; reference core/coord_system.h:250-267,295-302 with
        // flux_div_basic.h:49-71). Everything a face needs is separable: the tangential differences a cell publishes carry
        // 1/m_t of the cell's own row / column / plane (which is the face centre's tangential coordinate), the normal
        // difference of a face carries 1/m_n at the face, the finished flux is scaled by the area factor m_t1 m_t2 and the
        // divergence of a cell by J = 1/(m0 m1 m2). Row 0 of a table = m (as info::metric sees it), row 1 = 1/m at the cell
        // centres, row 2 = 1/m at the faces.
        template <int CONV, int VISC, bool UNIF, bool FUSED, int TI, int TJ, bool CURV = false, bool GHOSTW = false>
        __global__ void __launch_bounds__(TI*TJ + (GHOSTW ? 64 : 32), (TI*TJ > 256) ? 1 : 2)
        flux_div_narrow_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_rhs,
                               const __grid_constant__ CUtensorMap tmap_qout, const __grid_constant__ CUtensorMap tmap_in0,
                               const __grid_constant__ CUtensorMap tmap_in1, double* __restrict__ rhs,
                               const __grid_constant__ FluxParams P, const __grid_constant__ Dims G,
                               const __grid_constant__ Stage S, const double* __restrict__ inv_dx_tab,
                               const __grid_constant__ GhostMaps GM, const int* __restrict__ nbr_tab,
                               double* __restrict__ qout_raw, const double* __restrict__ met)
        {
            using L = Lay<TI, TJ>;
            constexpr int NCOMPUTE = L::NCOMP, NCW = NCOMPUTE/32;       // shadow the 256-thread defaults of the namespace
            constexpr int TIp = L::TIp, PLANE_BYTES = L::PLANE_BYTES, PLANE_STRIDE = L::PLANE_STRIDE, PW = L::PW, PSZ = L::PSZ;
            constexpr int STAGE_DOUBLES = L::STAGE_DOUBLES, OFF_P = L::OFF_P, OFF_STAGE_K = L::OFF_STAGE_K, OFF_STAGE_Q = L::OFF_STAGE_Q;
            constexpr int OFF_FX = L::OFF_FX, OFF_FY = L::OFF_FY, OFF_BAR = L::OFF_BAR, OFF_NBR = L::OFF_NBR;
            extern __shared__ __align__(128) double smem_raw[];
            double*   ring    = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u)/8u;
            double*   pub     = ring + OFF_P;
            double*   stage_k = ring + OFF_STAGE_K;     // rhs / residual-register plane on its way out (fused: also input 1 on its way in)
            double*   stage_q = ring + OFF_STAGE_Q;     // fused: q_out plane on its way out (also input 0 on its way in)
            double*   Fx      = ring + OFF_FX;
            double*   Fy      = ring + OFF_FY;
            uint64_t* bars    = (uint64_t*)(ring + OFF_BAR);
            int*      nbr_s   = (int*)(ring + OFF_NBR); // fused ghost exchange: destination block per neighbour direction

            const int tid = threadIdx.x;
            const int lane = tid & 31, warp = tid >> 5;
            const bool is_edge = (warp == NCW), is_ghost = GHOSTW && (warp == NCW + 1);

            int t = blockIdx.x;
            const int ti = t % G.tiles_i; t /= G.tiles_i;
            const int tj = t % G.tiles_j; t /= G.tiles_j;
            const long long lb = G.blist ? (long long)G.blist[t] : G.lb0 + t;
            const int i0 = ti*TI, j0 = tj*TJ;
            const int nz = G.nx[2];
            const int ni_t = min(TI, G.nx[0] - i0);
            const int nj_t = min(TJ, G.nx[1] - j0);
            const Spacing<UNIF> H(G, inv_dx_tab, lb);
            // metric rows of this block: MT(d, row, idx); indices are clamped so that idle lanes of ragged tiles stay in range
            const double* mt = CURV ? met + lb*9*(long long)G.lm : nullptr;
            auto MT = [&](const int d, const int row, const int idx) { return __ldg(mt + (d*3 + row)*G.lm + min(idx, G.np[d])); };

            // TMA coordinates (fused (v,i) dimension first); the box starts one cell early if the halo start is odd
            const int ash = (i0 + G.ng[0] - 1) & 1;
            const int c0 = 5*(i0 + G.ng[0] - 1 - ash);
            const int c1 = j0 + G.ng[1] - 1;
            const int c2base = G.ng[2] - 1;               // plane p <-> k = p - 1
            const int nplanes = nz + 2;

            if (tid == NCOMPUTE)
            {
                prefetch_tmap(&tmap_q);
                if (G.tma_store) prefetch_tmap(&tmap_rhs);
                if (FUSED) { prefetch_tmap(&tmap_qout); if (S.nin > 0) prefetch_tmap(&tmap_in0); if (S.nin > 1) prefetch_tmap(&tmap_in1); }
                #pragma unroll
                for (int s = 0; s < NP + 1; ++s) mbar_init(&bars[s], 1);
                fence_mbar_init();
            }
            __syncthreads();
            if (tid == NCOMPUTE)
            {
                #pragma unroll
                for (int p = 0; p < NP; ++p)
                    if (p < nplanes)
                    {
                        mbar_arrive_expect_tx(&bars[p], PLANE_BYTES);
                        tma_load_4d(ring + p*PLANE_STRIDE, &tmap_q, &bars[p], c0, c1, c2base + p, (int)lb);
                    }
            }

            // raw cell (ci, cj) of a staged plane; ci, cj are tile-local and may be -1 .. TI / TJ
            auto cell_off = [&](int ci, int cj) { return ((cj + 1)*TIp + (ci + 1 + ash))*5; };
            auto pidx = [&](int ci, int cj) { return (cj + 1)*PW + (ci + 1); };

            mbar_wait(&bars[0], 0);
            mbar_wait(&bars[1], 0);

            // Ring: plane p lives in slot p % 3. At step k (plane index pk = k + 1) planes pk and pk+1 are resident and
            // pk+2 is in flight; plane pk is last read before barrier (2) of step k, after which its slot is re-armed
            // with plane pk+3 (one full step ahead of its first use). Plane 0 (k = -1) is consumed by the prologue.
            if (!is_edge && !is_ghost)
            {
                // ======================= compute warps: one cell column per thread =======================
                const int il = tid % TI, jl = tid / TI;
                const bool active = (il < ni_t) && (jl < nj_t);
                const int co = cell_off(il, jl);
                const int po = pidx(il, jl);
                const int so = (jl*TI + il)*5;                  // this thread's slot in the staging tiles
                // loop-carried state: the own column lives in three register sets (cells k-1, k, k+1) whose roles rotate with
                // the ring period, so the k loop is unrolled by 3 and no register is ever moved: per cell the state q, its
                // density, the in-plane differences the next z-face needs (dc = Dx.u + Dy.v, dxw, dyw) and the plane's 1/m_z
                struct Col { double q[5]; double rho, dc, dxw, dyw, jk; };
                Col A, B, C3;
                #pragma unroll
                for (int v = 0; v < 5; ++v) { A.q[v] = ring[co + v]; B.q[v] = ring[PLANE_STRIDE + co + v]; C3.q[v] = B.q[v]; }
                A.rho = density(P.R, A.q[0], A.q[1]);
                B.rho = density(P.R, B.q[0], B.q[1]);
                if (!active) { A.rho = 1.0; B.rho = 1.0; }
                C3.rho = 1.0;
                A.jk = B.jk = C3.jk = 1.0;
                // scales of this column: tangential differences (cx, cy), normal differences of the lower x / y face (gx, gy),
                // area factors and the in-plane part of the Jacobian
                const int ipc = i0 + il + G.ng[0], jpc = j0 + jl + G.ng[1];
                double cx = H.c0, cy = H.c1, gx = H.i0, gy = H.i1, ar0 = 1.0, ar1 = 1.0, jij = 1.0;
                if (CURV)
                {
                    const double r0 = MT(0, 1, ipc), r1 = MT(1, 1, jpc);
                    cx = H.c0*r0; cy = H.c1*r1; jij = r0*r1;
                    gx = H.i0*MT(0, 2, ipc); gy = H.i1*MT(1, 2, jpc);
                    ar0 = MT(0, 0, ipc); ar1 = MT(1, 0, jpc);
                }
                A.dc = A.dxw = A.dyw = 0.0; B.dc = B.dxw = B.dyw = 0.0; C3.dc = C3.dxw = C3.dyw = 0.0;
                if (VISC)
                {
                    const double* pl = ring;
                    A.dc  = cx*(pl[co + 5 + 2] - pl[co - 5 + 2]) + cy*(pl[co + 5*TIp + 3] - pl[co - 5*TIp + 3]);
                    A.dxw = cx*(pl[co + 5 + 4] - pl[co - 5 + 4]);
                    A.dyw = cy*(pl[co + 5*TIp + 4] - pl[co - 5*TIp + 4]);
                }
                double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
                const long long cell0 = lb*G.block_stride
                    + 5ll*((i0 + il + G.ng[0]) + (long long)G.np[0]*((j0 + jl + G.ng[1]) + (long long)G.np[1]*G.ng[2]));
                const long long kstride = 5ll*G.np[0]*G.np[1];
                __syncthreads();                                                // (0) plane 0 consumed

                // one k step; SK / SP = ring slots of planes k and k+1 (compile-time: every shared-memory address of the step is
                // an immediate offset), m / c / p = the register sets of cells k-1, k, k+1, par = mbarrier parity of plane k+1
                auto step = [&](auto sk_tag, auto sp_tag, const int k, Col& m, Col& c, Col& p, const uint32_t par)
                {
                    constexpr int SK = decltype(sk_tag)::v, SP = decltype(sp_tag)::v;
                    const int pk = k + 1;                       // plane index of k
                    const double* plk = ring + SK*PLANE_STRIDE + co;
                    const double* plp = ring + SP*PLANE_STRIDE + co;

                    // in-plane neighbours of cell k (plane k has been resident since the previous step): u, v, w of the x- and
                    // y-neighbours. The lower ones are kept: they are 3 of the 5 values the lower x / y face needs after (1).
                    double xl[3] = {0.0, 0.0, 0.0}, yl[3] = {0.0, 0.0, 0.0};
                    double dxu = 0.0, dxv = 0.0, dxw = 0.0, dyu = 0.0, dyv = 0.0, dyw = 0.0;
                    if (VISC)
                    {
                        #pragma unroll
                        for (int v = 0; v < 3; ++v) { xl[v] = plk[-5 + 2 + v]; yl[v] = plk[-5*TIp + 2 + v]; }
                        dxu = cx*(plk[5 + 2] - xl[0]);
                        dxv = cx*(plk[5 + 3] - xl[1]);
                        dxw = cx*(plk[5 + 4] - xl[2]);
                        dyu = cy*(plk[5*TIp + 2] - yl[0]);
                        dyv = cy*(plk[5*TIp + 3] - yl[1]);
                        dyw = cy*(plk[5*TIp + 4] - yl[2]);
                    }
                    const double cZ = dxu + dyv;
                    c.dc = cZ; c.dxw = dxw; c.dyw = dyw;
                    // plane k: z scales (the same for the whole plane)
                    double cz = H.c2, gz = H.i2, ar2 = 1.0;
                    if (CURV)
                    {
                        const int kp = k + G.ng[2];
                        c.jk = MT(2, 1, kp); cz = H.c2*c.jk; gz = H.i2*MT(2, 2, kp); ar2 = MT(2, 0, kp);
                    }
                    // z-face k-1/2: registers only, overlaps the wait for plane k+1. acc carries the divergence of a cell:
                    // lower-face fluxes are added as they are computed, the neighbours' (upper-face) fluxes are subtracted
                    // after barrier (2).
                    {
                        double Fz[5];
                        face<CONV, VISC, 2>(P, m.q, c.q, m.rho, c.rho, m.dc + cZ, m.dxw + dxw, m.dyw + dyw, gz, Fz);
                        if (CURV)
                        {
                            const double az = ar0*ar1;
                            #pragma unroll
                            for (int v = 0; v < 5; ++v) Fz[v] *= az;
                        }
                        if (k >= 1)
                        {
                            double r[5];
                            #pragma unroll
                            for (int v = 0; v < 5; ++v) r[v] = fma(-Fz[v], H.i2, acc[v]);      // rhs of cell k-1
                            if (CURV)
                            {
                                const double jac = jij*m.jk;                                    // J of cell k-1
                                #pragma unroll
                                for (int v = 0; v < 5; ++v) r[v] *= jac;
                            }
                            if (FUSED)
                            {
                                double w[5], o[5];
                                #pragma unroll
                                for (int v = 0; v < 5; ++v) { w[v] = S.cq_self*r[v]; o[v] = S.co_self*r[v]; }
                                if (S.nin > 0)
                                {
                                    mbar_wait(&bars[NP], (k - 1) & 1);          // input tiles of cell k-1 have landed (TMA)
                                    #pragma unroll
                                    for (int v = 0; v < 5; ++v)
                                    {
                                        const double a0 = stage_q[so + v];
                                        w[v] = fma(S.cq[0], a0, w[v]); o[v] = fma(S.co[0], a0, o[v]);
                                    }
                                    if (S.nin > 1)
                                    {
                                        #pragma unroll
                                        for (int v = 0; v < 5; ++v)
                                        {
                                            const double a1 = stage_k[so + v];
                                            w[v] = fma(S.cq[1], a1, w[v]); o[v] = fma(S.co[1], a1, o[v]);
                                        }
                                    }
                                }
                                if (S.has_out)
                                {
                                    #pragma unroll
                                    for (int v = 0; v < 5; ++v) stage_k[so + v] = o[v];
                                }
                                // prim -> cons (fluid_state.h:103-116), add the increment, cons -> prim (fluid_state.h:119-135)
                                const double u2 = fma(m.q[2], m.q[2], fma(m.q[3], m.q[3], m.q[4]*m.q[4]));
                                const double rho  = m.rho + w[0];
                                const double rhoE = fma(0.5*m.rho, u2, m.q[0]*S.inv_gm1) + w[1];
                                const double mx = fma(m.rho, m.q[2], w[2]), my = fma(m.rho, m.q[3], w[3]), mz = fma(m.rho, m.q[4], w[4]);
                                const double ir = fast_rcp(active ? rho : 1.0);
                                const double un = ir*mx, vn = ir*my, wn = ir*mz;
                                const double pn = S.gm1*fma(-0.5*rho, fma(un, un, fma(vn, vn, wn*wn)), rhoE);
                                const double Tn = pn*ir*S.inv_R;
                                stage_q[so + 0] = pn;
                                stage_q[so + 1] = Tn;
                                stage_q[so + 2] = un; stage_q[so + 3] = vn; stage_q[so + 4] = wn;
                            }
                            else if (G.tma_store)
                            {
                                #pragma unroll
                                for (int v = 0; v < 5; ++v) stage_k[so + v] = r[v];
                            }
                            else if (active)
                            {
                                double* o = rhs + cell0 + (long long)(k - 1)*kstride;
                                #pragma unroll
                                for (int v = 0; v < 5; ++v)
                                {
                                    double rr = r[v];
                                    if (G.increment) rr += o[v];
                                    o[v] = rr;
                                }
                            }
                        }
                        #pragma unroll
                        for (int v = 0; v < 5; ++v) acc[v] = Fz[v]*H.i2;
                    }
                    if (pk + 1 < nplanes)
                    {
                        mbar_wait(&bars[SP], par);
                        #pragma unroll
                        for (int v = 0; v < 5; ++v) p.q[v] = plp[v];
                    }
                    p.rho = density(P.R, p.q[0], p.q[1]);               // stale at k = nz, never used
                    if (!active) p.rho = 1.0;
                    double cX = 0.0, cY = 0.0, dzu = 0.0, dzv = 0.0;
                    if (VISC && k < nz)
                    {
                        const double dzw = cz*(p.q[4] - m.q[4]);
                        dzu = cz*(p.q[2] - m.q[2]);
                        dzv = cz*(p.q[3] - m.q[3]);
                        cX = dyv + dzw; cY = dzw + dxu;
                    }
                    if (k < nz && !(SPB_EXP & 4))
                    {
                        pub[P_RHO*PSZ + po] = c.rho;
                        if (VISC)
                        {
                            pub[P_CX*PSZ + po] = cX; pub[P_DYU*PSZ + po] = dyu; pub[P_DZU*PSZ + po] = dzu;
                            pub[P_CY*PSZ + po] = cY; pub[P_DZV*PSZ + po] = dzv; pub[P_DXV*PSZ + po] = dxv;
                        }
                    }
                    if (G.tma_store) fence_proxy_async();
                    __syncthreads();                                            // (1) published data + staged planes visible
                    if (k < nz)
                    {
                        {
                            double qL[5], F[5];
                            qL[0] = plk[-5]; qL[1] = plk[-5 + 1];
                            #pragma unroll
                            for (int v = 0; v < 3; ++v) qL[2 + v] = VISC ? xl[v] : plk[-5 + 2 + v];
                            const int pl_ = po - 1;
                            const double rhoL = (SPB_EXP & 4) ? m.rho : pub[P_RHO*PSZ + pl_];
                            double ac = 0.0, b = 0.0, d = 0.0;
                            if (VISC && !(SPB_EXP & 4)) { ac = pub[P_CX*PSZ + pl_] + cX; b = pub[P_DYU*PSZ + pl_] + dyu; d = pub[P_DZU*PSZ + pl_] + dzu; }
                            face<CONV, VISC, 0>(P, qL, c.q, rhoL, c.rho, ac, b, d, gx, F);
                            if (CURV)
                            {
                                const double ax = ar1*ar2;
                                #pragma unroll
                                for (int v = 0; v < 5; ++v) F[v] *= ax;
                            }
                            if (active && !(SPB_EXP & 2))
                            {
                                #pragma unroll
                                for (int v = 0; v < 5; ++v) Fx[(jl*(TI + 1) + il)*5 + v] = F[v];
                            }
                            #pragma unroll
                            for (int v = 0; v < 5; ++v) acc[v] = fma(F[v], H.i0, acc[v]);
                        }
                        {
                            double qL[5], F[5];
                            qL[0] = plk[-5*TIp]; qL[1] = plk[-5*TIp + 1];
                            #pragma unroll
                            for (int v = 0; v < 3; ++v) qL[2 + v] = VISC ? yl[v] : plk[-5*TIp + 2 + v];
                            const int pl_ = po - PW;
                            const double rhoL = (SPB_EXP & 4) ? m.rho : pub[P_RHO*PSZ + pl_];
                            double ac = 0.0, b = 0.0, d = 0.0;
                            if (VISC && !(SPB_EXP & 4)) { ac = pub[P_CY*PSZ + pl_] + cY; b = pub[P_DZV*PSZ + pl_] + dzv; d = pub[P_DXV*PSZ + pl_] + dxv; }
                            face<CONV, VISC, 1>(P, qL, c.q, rhoL, c.rho, ac, b, d, gy, F);
                            if (CURV)
                            {
                                const double ay = ar2*ar0;
                                #pragma unroll
                                for (int v = 0; v < 5; ++v) F[v] *= ay;
                            }
                            if (active && !(SPB_EXP & 2))
                            {
                                #pragma unroll
                                for (int v = 0; v < 5; ++v) Fy[(jl*TI + il)*5 + v] = F[v];
                            }
                            #pragma unroll
                            for (int v = 0; v < 5; ++v) acc[v] = fma(F[v], H.i1, acc[v]);
                        }
                    }
                    SPB_BAR2();                                                 // (2) fluxes visible; pub, plane k, staging tiles free
                    if (k < nz && !(SPB_EXP & 2))
                    {
                        #pragma unroll
                        for (int v = 0; v < 5; ++v)
                        {
                            acc[v] = fma(-Fx[(jl*(TI + 1) + il + 1)*5 + v], H.i0, acc[v]);
                            acc[v] = fma(-Fy[((jl + 1)*TI + il)*5 + v], H.i1, acc[v]);
                        }
                    }
                };
                // plane p lives in slot p % 3 and cell k in plane k + 1: k = 3 kb + u has planes (k, k+1) in slots ((u+1)%3, (u+2)%3);
                // plane k + 2 = 3 (kb + (u+2)/3) + (u+2)%3 completes phase kb + (u+2)/3 of its slot's mbarrier
                for (int k = 0, kb = 0; k <= nz; k += 3, ++kb)
                {
                    step(IC<1>{}, IC<2>{}, k, A, B, C3, (uint32_t)(kb & 1));
                    if (k + 1 > nz) break;
                    step(IC<2>{}, IC<0>{}, k + 1, B, C3, A, (uint32_t)((kb + 1) & 1));
                    if (k + 2 > nz) break;
                    step(IC<0>{}, IC<1>{}, k + 2, C3, A, B, (uint32_t)((kb + 1) & 1));
                }
            }
            else if (is_edge)
            {
                // ======================= edge warp: halo ring + upper-edge faces + TMA traffic =======================
                // row job (lanes 0..31): cell (lane, -1) is published, cell (lane, nj_t) is the R cell of the upper y-face
                // col job (lanes 0..15): cell (-1, lane) is published, cell (ni_t, lane-8) is the R cell of the upper x-face
                const bool row_on = lane < ni_t;
                const int  ccj = lane % TJ;
                const bool col_lo = lane < TJ, col_on = (lane < 2*TJ) && (ccj < nj_t);
                const int  cci = col_lo ? -1 : ni_t;
                const int rl = min(lane, TI - 1);                // 16-wide tiles: the upper half of the warp has no row cell (its loads stay inside the plane)
                const int co_r0 = cell_off(rl, -1), co_r1 = cell_off(rl, nj_t), co_c = cell_off(cci, ccj);
                // scales of the edge cells: the row cells sit in column i0 + rl, the column cells in row j0 + ccj
                double ecx = H.c0, ecy = H.c1, egy = H.i1, egx = H.i0, ear0 = 1.0, ear1 = 1.0;
                if (CURV)
                {
                    ecx = H.c0*MT(0, 1, i0 + rl + G.ng[0]);
                    ecy = H.c1*MT(1, 1, j0 + ccj + G.ng[1]);
                    egy = H.i1*MT(1, 2, j0 + nj_t + G.ng[1]);          // upper y-face of the tile
                    egx = H.i0*MT(0, 2, i0 + ni_t + G.ng[0]);          // upper x-face of the tile
                    ear0 = MT(0, 0, i0 + rl + G.ng[0]);
                    ear1 = MT(1, 0, j0 + ccj + G.ng[1]);
                }
                // z-neighbours (k-1, k, k+1) of the edge cells live in three rotating register sets like the compute warps'
                // columns: v,w of the two row cells, u,w of the column cell
                struct ECol { double r0[2], r1[2], cc[2]; };
                ECol EA, EB, EC;
                {
                    const double* pa = ring; const double* pb = ring + PLANE_STRIDE;
                    EA.r0[0] = pa[co_r0 + 3]; EA.r0[1] = pa[co_r0 + 4]; EB.r0[0] = pb[co_r0 + 3]; EB.r0[1] = pb[co_r0 + 4];
                    EA.r1[0] = pa[co_r1 + 3]; EA.r1[1] = pa[co_r1 + 4]; EB.r1[0] = pb[co_r1 + 3]; EB.r1[1] = pb[co_r1 + 4];
                    EA.cc[0] = pa[co_c + 2];  EA.cc[1] = pa[co_c + 4];  EB.cc[0] = pb[co_c + 2];  EB.cc[1] = pb[co_c + 4];
                    EC = EB;
                }
                __syncthreads();                                                // (0) plane 0 consumed
                if (lane == 0 && NP < nplanes)
                {
                    mbar_arrive_expect_tx(&bars[0], PLANE_BYTES);
                    tma_load_4d(ring, &tmap_q, &bars[0], c0, c1, c2base + NP, (int)lb);
                }
                auto estep = [&](auto sk_tag, auto sp_tag, const int k, ECol& m, ECol& c, ECol& p, const uint32_t par)
                {
                    constexpr int SK = decltype(sk_tag)::v, SP = decltype(sp_tag)::v;
                    const int pk = k + 1;
                    const double* plk = ring + SK*PLANE_STRIDE;
                    const double* plp = ring + SP*PLANE_STRIDE;
                    p.r0[0] = p.r0[1] = p.r1[0] = p.r1[1] = p.cc[0] = p.cc[1] = 0.0;
                    if (pk + 1 < nplanes)
                    {
                        mbar_wait(&bars[SP], par);
                        p.r0[0] = plp[co_r0 + 3]; p.r0[1] = plp[co_r0 + 4];
                        p.r1[0] = plp[co_r1 + 3]; p.r1[1] = plp[co_r1 + 4];
                        p.cc[0] = plp[co_c + 2];  p.cc[1] = plp[co_c + 4];
                    }
                    // ---- before (1): publish the lower halo, and prepare what the upper faces need from their R cells
                    double uCY = 0.0, uDzv = 0.0, uDxv = 0.0, xCX = 0.0, xDyu = 0.0, xDzu = 0.0;
                    double uRho = 1.0, xRho = 1.0;
                    double ecz = H.c2, ear2 = 1.0;
                    if (CURV) { ecz = H.c2*MT(2, 1, k + G.ng[2]); ear2 = MT(2, 0, k + G.ng[2]); }
                    if (k < nz)
                    {
                        if (row_on)
                        {
                            const int po = pidx(lane, -1);
                            pub[P_RHO*PSZ + po] = density(P.R, plk[co_r0], plk[co_r0 + 1]);
                            uRho = density(P.R, plk[co_r1], plk[co_r1 + 1]);
                            if (VISC)
                            {
                                pub[P_CY*PSZ + po]  = ecz*(p.r0[1] - m.r0[1]) + ecx*(plk[co_r0 + 5 + 2] - plk[co_r0 - 5 + 2]);
                                pub[P_DZV*PSZ + po] = ecz*(p.r0[0] - m.r0[0]);
                                pub[P_DXV*PSZ + po] = ecx*(plk[co_r0 + 5 + 3] - plk[co_r0 - 5 + 3]);
                                uCY  = ecz*(p.r1[1] - m.r1[1]) + ecx*(plk[co_r1 + 5 + 2] - plk[co_r1 - 5 + 2]);
                                uDzv = ecz*(p.r1[0] - m.r1[0]);
                                uDxv = ecx*(plk[co_r1 + 5 + 3] - plk[co_r1 - 5 + 3]);
                            }
                        }
                        if (col_on)
                        {
                            xRho = density(P.R, plk[co_c], plk[co_c + 1]);
                            if (VISC)
                            {
                                xCX  = ecy*(plk[co_c + 5*TIp + 3] - plk[co_c - 5*TIp + 3]) + ecz*(p.cc[1] - m.cc[1]);
                                xDyu = ecy*(plk[co_c + 5*TIp + 2] - plk[co_c - 5*TIp + 2]);
                                xDzu = ecz*(p.cc[0] - m.cc[0]);
                            }
                            if (col_lo)
                            {
                                const int po = pidx(-1, ccj);
                                pub[P_RHO*PSZ + po] = xRho;
                                if (VISC) { pub[P_CX*PSZ + po] = xCX; pub[P_DYU*PSZ + po] = xDyu; pub[P_DZU*PSZ + po] = xDzu; }
                            }
                        }
                    }
                    __syncthreads();                                            // (1)
                    if (lane == 0 && G.tma_store && k >= 1)
                    {
                        // `increment` trait: the same staged plane leaves as a tensor reduction (rhs += tile, one fp64 add per
                        // element at the L2: bit-identical to a load-add-store, and the SMs never read rhs)
                        if (!FUSED && G.increment) tma_reduce_add_4d(&tmap_rhs, stage_k, 5*i0, j0, k - 1, (int)lb);
                        else if (!FUSED || S.has_out) tma_store_4d(&tmap_rhs, stage_k, 5*i0, j0, k - 1, (int)lb);
                        if (FUSED) tma_store_4d(&tmap_qout, stage_q, 5*i0, j0, k - 1, (int)lb);
                        tma_store_commit();
                    }
                    if (k < nz)
                    {
                        if (row_on)                                              // upper y-face (lane, nj_t)
                        {
                            double qL[5], qR[5], F[5];
                            #pragma unroll
                            for (int v = 0; v < 5; ++v) { qL[v] = plk[co_r1 - 5*TIp + v]; qR[v] = plk[co_r1 + v]; }
                            const int pl_ = pidx(lane, nj_t - 1);
                            const double rhoL = pub[P_RHO*PSZ + pl_];
                            double ac = 0.0, b = 0.0, d = 0.0;
                            if (VISC) { ac = pub[P_CY*PSZ + pl_] + uCY; b = pub[P_DZV*PSZ + pl_] + uDzv; d = pub[P_DXV*PSZ + pl_] + uDxv; }
                            face<CONV, VISC, 1>(P, qL, qR, rhoL, uRho, ac, b, d, egy, F);
                            if (CURV)
                            {
                                const double ay = ear2*ear0;
                                #pragma unroll
                                for (int v = 0; v < 5; ++v) F[v] *= ay;
                            }
                            #pragma unroll
                            for (int v = 0; v < 5; ++v) Fy[(nj_t*TI + lane)*5 + v] = F[v];
                        }
                        if (col_on && !col_lo)                                   // upper x-face (ni_t, ccj)
                        {
                            double qL[5], qR[5], F[5];
                            #pragma unroll
                            for (int v = 0; v < 5; ++v) { qL[v] = plk[co_c - 5 + v]; qR[v] = plk[co_c + v]; }
                            const int pl_ = pidx(ni_t - 1, ccj);
                            const double rhoL = pub[P_RHO*PSZ + pl_];
                            double ac = 0.0, b = 0.0, d = 0.0;
                            if (VISC) { ac = pub[P_CX*PSZ + pl_] + xCX; b = pub[P_DYU*PSZ + pl_] + xDyu; d = pub[P_DZU*PSZ + pl_] + xDzu; }
                            face<CONV, VISC, 0>(P, qL, qR, rhoL, xRho, ac, b, d, egx, F);
                            if (CURV)
                            {
                                const double ax = ear1*ear2;
                                #pragma unroll
                                for (int v = 0; v < 5; ++v) F[v] *= ax;
                            }
                            #pragma unroll
                            for (int v = 0; v < 5; ++v) Fx[(ccj*(TI + 1) + ni_t)*5 + v] = F[v];
                        }
                    }
                    if (lane == 0 && G.tma_store) tma_store_wait_read<0>();     // staging tiles are rewritten after (2)
                    SPB_BAR2();                                                 // (2)
                    if (FUSED && lane == 0 && S.nin > 0 && k < nz)
                    {
                        // input tiles of cell plane k (its rhs completes in the next step) land in the staging tiles, which
                        // the stores of plane k-1 have finished reading (wait_read above)
                        mbar_arrive_expect_tx(&bars[NP], (uint32_t)(STAGE_DOUBLES*8*S.nin));
                        tma_load_4d(stage_q, &tmap_in0, &bars[NP], 5*i0, j0, k, (int)lb);
                        if (S.nin > 1) tma_load_4d(stage_k, &tmap_in1, &bars[NP], 5*i0, j0, k, (int)lb);
                        if (k + 1 < nz)                                          // next plane's tiles: into L2 now
                        {
                            tma_prefetch_4d(&tmap_in0, 5*i0, j0, k + 1, (int)lb);
                            if (S.nin > 1) tma_prefetch_4d(&tmap_in1, 5*i0, j0, k + 1, (int)lb);
                        }
                    }
                    if (lane == 0)
                    {
                        // plane pk has been consumed by every thread: re-arm its slot with plane pk + NP
                        const int pnew = pk + NP;
                        if (pnew < nplanes)
                        {
                            mbar_arrive_expect_tx(&bars[SK], PLANE_BYTES);
                            tma_load_4d(ring + SK*PLANE_STRIDE, &tmap_q, &bars[SK], c0, c1, c2base + pnew, (int)lb);
                            if (pnew + 1 < nplanes) tma_prefetch_4d(&tmap_q, c0, c1, c2base + pnew + 1, (int)lb);   // and the one after into L2
                        }
                    }
                };
                for (int k = 0, kb = 0; k <= nz; k += 3, ++kb)
                {
                    estep(IC<1>{}, IC<2>{}, k, EA, EB, EC, (uint32_t)(kb & 1));
                    if (k + 1 > nz) break;
                    estep(IC<2>{}, IC<0>{}, k + 1, EB, EC, EA, (uint32_t)((kb + 1) & 1));
                    if (k + 2 > nz) break;
                    estep(IC<0>{}, IC<1>{}, k + 2, EC, EA, EB, (uint32_t)((kb + 1) & 1));
                }
                if (lane == 0 && G.tma_store) tma_store_wait<0>();
            }
            else if (GHOSTW)
            {
                // ======================= ghost warp: same-rank ghost exchange of the finished q_out plane =======================
                // A cell of the source box of direction e = (ex,ey,ez) goes to cell (i - ex n0, j - ey n1, k - ez n2) of the
                // neighbour block (get_transaction.h:54-86). Directions with ex = 0 are whole rows of the staged tile: TMA stores
                // through the clipped views GM.m[e] (lane e). Directions with ex != 0 are 2-cell pieces of a row: read from the
                // staged tile and stored with plain 8-byte stores, 10 consecutive doubles per row (TMA stores cannot start at a
                // negative coordinate, tools/probes/tma_store_probe.cu, and 80-byte boxes cost a TMA operation each).
                const bool on = FUSED && G.ghost;
                const int xhi0 = G.nx[0] - G.ng[0] - i0, yhi0 = G.nx[1] - G.ng[1] - j0;      // tile-local start of the +1 source cells
                if (on && lane < 27) nbr_s[lane] = nbr_tab[27*lb + lane];
                int gdst = -1, gc1 = 0, gzs = 0, gez = 0;
                const double* gsrc = stage_q;
                if (on && lane < 27 && lane != 13 && (lane % 3) == 1)
                {
                    const int ey = (lane/3) % 3 - 1;
                    gez = lane/9 - 1;
                    const bool need_y = (ey == 0) || (ey < 0 ? (j0 < G.ng[1]) : (yhi0 >= 0 && yhi0 < TJ));
                    if (need_y) gdst = nbr_tab[27*lb + lane];
                    gc1 = ey > 0 ? 0 : j0;                                   // ey > 0: source advanced to the first y-high row
                    gsrc = ey > 0 ? stage_q + yhi0*(5*TI) : stage_q;
                    gzs = gez > 0 ? nz - G.ng[2] : 0;
                    prefetch_tmap(&GM.m[lane]);
                }
                // x pieces: piece p = (side, row r) is 10 consecutive doubles = 5 double2. With TJ = 8 there are 16 pieces and two
                // lanes share one (half 0 moves double2 0, 2, 4, half 1 moves 1, 3); with TJ = 16 every lane owns a whole piece.
                // Everything that does not depend on the plane is computed here.
                constexpr int NPIECE = 2*TJ, HALVES = 32/NPIECE;
                const int xp = lane % NPIECE, xh = lane / NPIECE;
                const int xside = xp / TJ, xr = xp % TJ;
                const bool x_on = on && (xr < nj_t) && (xside ? (xhi0 >= 0 && xhi0 < TI) : (i0 == 0));
                const int xsm = xr*(5*TI) + (xside ? 5*xhi0 : 0) + 2*xh;                       // first double2 of this lane in stage_q
                const int xj = j0 + xr;
                const int xip = (xside ? xhi0 + i0 - G.nx[0] : i0 + G.nx[0]) + G.ng[0];       // padded i of the first destination cell
                const bool y_lo = x_on && (xj < G.ng[1]), y_hi = x_on && (xj >= G.nx[1] - G.ng[1]);
                const long long xrow = 5ll*G.np[0], xplane = 5ll*G.np[0]*G.np[1];
                __syncthreads();                                                // (0)
                for (int k = 0; k <= nz; ++k)
                {
                    __syncthreads();                                            // (1) stage_q holds plane k - 1 of q_out
                    if (on && k >= 1)
                    {
                        const int kk = k - 1;
                        if (gdst >= 0)
                        {
                            const bool need_z = (gez == 0) || (gez < 0 ? (kk < G.ng[2]) : (kk >= nz - G.ng[2]));
                            if (need_z)
                            {
                                tma_store_4d(&GM.m[lane], gsrc, 5*i0, gc1, kk - gzs, gdst);
                                tma_store_commit();
                            }
                        }
                        if (x_on)
                        {
                            // this lane's double2 of the piece: indices xh, xh + HALVES, ... < 5
                            constexpr int NM = (5 + HALVES - 1)/HALVES;
                            double2 a[NM];
                            #pragma unroll
                            for (int m = 0; m < NM; ++m)
                                if (xh + m*HALVES < 5) a[m] = *reinterpret_cast<const double2*>(stage_q + xsm + 2*m*HALVES);
                            #pragma unroll
                            for (int ez = -1; ez <= 1; ++ez)
                            {
                                if (!(ez == 0 || (ez < 0 ? (kk < G.ng[2]) : (kk >= nz - G.ng[2])))) continue;     // uniform
                                #pragma unroll
                                for (int ey = -1; ey <= 1; ++ey)
                                {
                                    if (!(ey == 0 || (ey < 0 ? y_lo : y_hi))) continue;
                                    const int dst = nbr_s[2*xside + 3*(ey + 1) + 9*(ez + 1)];
                                    if (dst < 0) continue;
                                    double* o = qout_raw + dst*G.block_stride + 5ll*xip + xrow*(xj - ey*G.nx[1] + G.ng[1])
                                                + xplane*(kk - ez*nz + G.ng[2]) + 2*xh;
                                    #pragma unroll
                                    for (int m = 0; m < NM; ++m)
                                        if (xh + m*HALVES < 5) *reinterpret_cast<double2*>(o + 2*m*HALVES) = a[m];
                                }
                            }
                        }
                    }
                    if (gdst >= 0) tma_store_wait_read<0>();                    // stage_q is refilled after (2)
                    SPB_BAR2();                                                 // (2)
                }
                if (gdst >= 0) tma_store_wait<0>();
            }
        }

        static int make_map(CUtensorMap* m, const void* base, const cuuint64_t dims[4], const cuuint64_t strides[3],
                            const cuuint32_t box[4], CUtensorMapL2promotion prom, const char* what)
        {
            encode_tiled_fn enc = get_encode_tiled();
            if (!enc) { set_error("spb_flux_div: cuTensorMapEncodeTiled not available from the driver"); return SPB_ERR_DRIVER; }
            const cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult cr = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)base, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) { set_error(std::string("spb_flux_div: cuTensorMapEncodeTiled(") + what + ") failed with CUresult " + std::to_string((int)cr)); return SPB_ERR_DRIVER; }
            return 0;
        }
    }

    // stage == nullptr: plain flux_div; otherwise the fused RK stage (q_out, stage description)
    template <int CONV, int VISC, int TI, int TJ>
    static int launch_fdiv_narrow_tile(const spb_grid* g, const double* q, double* rhs, const FluxParams& P, int increment,
                                       int64_t lb_begin, int64_t lb_end, cudaStream_t stream, double* q_out, const nrw::Stage* stage,
                                       spb_exchange* exch)
    {
        using namespace nrw;
        using L = Lay<TI, TJ>;
        constexpr int TIp = L::TIp, TJp = L::TJp, SMEM_BYTES = L::SMEM_BYTES;
        for (int d = 0; d < 3; ++d)
            if (g->ng[d] < 1) { set_error("spb_flux_div: scheme needs 1 exchange cell"); return SPB_ERR_BAD_ARG; }
        if ((5*g->np[0]) % 2 != 0) { set_error("spb_flux_div: n0 + 2*g0 must be even (16-byte TMA row pitch)"); return SPB_ERR_UNSUPPORTED; }

        CUtensorMap tq, tr, tqo, ti0, ti1;
        const cuuint64_t strides[3] = {(cuuint64_t)40*g->np[0], (cuuint64_t)40*g->np[0]*g->np[1], (cuuint64_t)8*g->block_stride};
        {
            const cuuint64_t dims[4] = {(cuuint64_t)5*g->np[0], (cuuint64_t)g->np[1], (cuuint64_t)g->np[2], (cuuint64_t)g->nlb};
            const cuuint32_t box[4]  = {(cuuint32_t)(5*TIp), (cuuint32_t)TJp, 1, 1};
            int rc = make_map(&tq, q, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "q"); if (rc) return rc;
        }
        // interior-only views for the tensor stores: the hardware clips ragged tiles to the interior
        const long long org = 5ll*(g->ng[0] + (long long)g->np[0]*(g->ng[1] + (long long)g->np[1]*g->ng[2]));
        const cuuint64_t idims[4] = {(cuuint64_t)5*g->nx[0], (cuuint64_t)g->nx[1], (cuuint64_t)g->nx[2], (cuuint64_t)g->nlb};
        const cuuint32_t ibox[4]  = {(cuuint32_t)(5*TI), (cuuint32_t)TJ, 1, 1};
        const bool aligned = (org % 2 == 0);
        int tma_store = (aligned && rhs && (((uintptr_t)rhs) % 16 == 0)) ? 1 : 0;
        if (tma_store && make_map(&tr, rhs + org, idims, strides, ibox, CU_TENSOR_MAP_L2_PROMOTION_NONE, "rhs")) tma_store = 0;
        if (!tma_store) tr = tq;
        tqo = tq;
        if (stage)
        {
            if (!aligned || !q_out || (((uintptr_t)q_out) % 16 != 0) || (stage->has_out && !tma_store))
            { set_error("spb_flux_div_rk_stage: arrays must be 16-byte aligned with an even interior origin"); return SPB_ERR_UNSUPPORTED; }
            int rc = make_map(&tqo, q_out + org, idims, strides, ibox, CU_TENSOR_MAP_L2_PROMOTION_NONE, "q_out"); if (rc) return rc;
            tma_store = 1;
        }
        ti0 = tq; ti1 = tq;
        if (stage)
            for (int a = 0; a < stage->nin; ++a)
            {
                if (((uintptr_t)stage->in[a]) % 16 != 0) { set_error("spb_flux_div_rk_stage: input arrays must be 16-byte aligned"); return SPB_ERR_UNSUPPORTED; }
                int rc = make_map(a == 0 ? &ti0 : &ti1, stage->in[a] + org, idims, strides, ibox, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "in"); if (rc) return rc;
            }

        // fused same-rank ghost exchange: q_out's ghost boxes as 26 clipped views (reference boxes: get_transaction.h:54-86)
        static thread_local GhostMaps GM;
        const int* nbr_tab = nullptr;
        if (stage && exch)
        {
            int rc = exchange_fuse_table(exch, g->nx, g->ng, g->nlb, &nbr_tab); if (rc) return rc;
            if (g->ng[0] != 2) { set_error("fused exchange: implemented for 2 exchange cells along i (use spb_exchange_local)"); return SPB_ERR_UNSUPPORTED; }
            for (int d = 0; d < 2; ++d)
            {
                const int T = d == 0 ? TI : TJ;
                if (g->ng[d] > g->nx[d] || (g->nx[d] - g->ng[d])/T != (g->nx[d] - 1)/T)
                { set_error("fused exchange: the high source box straddles two tiles (use spb_exchange_local)"); return SPB_ERR_UNSUPPORTED; }
            }
            const long long cstride[3] = {5ll, 5ll*g->np[0], 5ll*g->np[0]*g->np[1]};
            for (int e = 0; e < 27; ++e)
            {
                const int ed[3] = {e % 3 - 1, (e/3) % 3 - 1, e/9 - 1};
                if (e == 13 || ed[0] != 0) { GM.m[e] = tqo; continue; }       // ex != 0: plain stores by the ghost warp
                long long shift = 0;
                cuuint64_t gd[4];
                for (int d = 0; d < 3; ++d)
                {
                    shift += cstride[d]*(ed[d] == 0 ? 0 : (ed[d] < 0 ? g->nx[d] : -g->ng[d]));
                    gd[d] = (cuuint64_t)(ed[d] == 0 ? g->nx[d] : g->ng[d]);
                }
                gd[0] *= 5; gd[3] = (cuuint64_t)g->nlb;
                // box: whole tile rows; TJ rows, or only the ng[1] source rows for ey != 0
                const cuuint32_t gbox[4] = {(cuuint32_t)(5*TI), (cuuint32_t)(ed[1] == 0 ? TJ : g->ng[1]), 1, 1};
                rc = make_map(&GM.m[e], q_out + org + shift, gd, strides, gbox, CU_TENSOR_MAP_L2_PROMOTION_NONE, "q_out ghost box"); if (rc) return rc;
            }
        }

        Dims G;
        for (int d = 0; d < 3; ++d) { G.nx[d] = g->nx[d]; G.ng[d] = g->ng[d]; G.np[d] = g->np[d]; }
        G.ghost = nbr_tab ? 1 : 0;
        G.tiles_i = (g->nx[0] + TI - 1)/TI;
        G.tiles_j = (g->nx[1] + TJ - 1)/TJ;
        G.block_stride = g->block_stride;
        G.lb0 = lb_begin;
        G.increment = increment;
        G.lm = g->metric_lm;
        G.tma_store = tma_store;
        const BlockList& bl = current_block_list();
        G.blist = bl.dev;
        if (bl.dev) { lb_begin = 0; lb_end = g->nlb; }                  // the spacing check below then covers every block
        const int64_t nblk = (bl.dev ? bl.count : lb_end - lb_begin)*G.tiles_i*G.tiles_j;
        if (nblk <= 0) return 0;
        // A lattice counts as uniform when the per-block spacings agree up to what the rounding of the block bounds explains:
        // box.size/num_cell formed block by block (cartesian_grid.h:134-135) wobbles with the block origin by 2 ulp of the
        // COORDINATE — 10 and 20 ulp of the spacing on ranks 1 and 2 of a weak-scaled box whose z runs to 2 pi N. A fixed 8-ulp
        // test (rounds 1 and 2) sent exactly those ranks to the per-block-table variant, which is 6 % slower, and every other
        // rank waited for them (profiles/r02_pair_diagnosis.log). The common spacing changes the rhs by < 1e-13 relative.
        bool uniform = true;
        for (int64_t b = lb_begin; b < lb_end && uniform; ++b)
            for (int d = 0; d < 3; ++d)
            {
                const double a = g->inv_dx_host[3*b + d], r = g->inv_dx_host[3*lb_begin + d];
                const double tol = fmin(fmax(8.0*2.220446049250313e-16, g->spacing_round_tol[d]), 1e-13);
                uniform = uniform && (fabs(a - r) <= tol*fabs(r));
            }
        for (int d = 0; d < 3; ++d) { G.idx[d] = g->inv_dx_host[3*lb_begin + d]; G.cdx[d] = 0.25*G.idx[d]; }
        G.lev = g->lev_dev;
        for (int d = 0; d < 3; ++d)
            for (int l = 0; l < 16; ++l) { G.lev_inv[d][l] = g->lev_inv[d][l]; G.lev_cinv[d][l] = 0.25*g->lev_inv[d][l]; }
        if (!uniform && (g->lev_n[0] < 0 || g->lev_n[1] < 0 || g->lev_n[2] < 0))
        { set_error("spb_flux_div: more than 16 distinct block spacings along one direction"); return SPB_ERR_UNSUPPORTED; }
        Stage S{};
        if (stage) S = *stage;
        constexpr int NTHREADS = L::NCOMP + 64, NTHREADS_NOGHOST = L::NCOMP + 32;
        auto go = [&](auto kern, const int nthreads = L::NCOMP + 32) -> int
        {
            SPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            SPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            kern<<<(unsigned)nblk, nthreads, SMEM_BYTES, stream>>>(tq, tr, tqo, ti0, ti1, rhs, P, G, S, g->inv_dx_dev, GM, nbr_tab, q_out, g->metric_dev);
            SPB_LAUNCH_CHECK();
            return 0;
        };
        if (g->metric_dev)
        {
            // general coordinates: the computational lattice must be uniform (per-block dxi with a metric is the wide kernel's job)
            if (!uniform) { set_error("spb_flux_div: general coordinates on a non-uniform block lattice run on the wide kernel"); return SPB_ERR_UNSUPPORTED; }
            if (stage && G.ghost) return go(flux_div_narrow_kernel<CONV, VISC, true, true, TI, TJ, true, true>, NTHREADS);
            return stage ? go(flux_div_narrow_kernel<CONV, VISC, true, true, TI, TJ, true>) : go(flux_div_narrow_kernel<CONV, VISC, true, false, TI, TJ, true>);
        }
        // the fused stage with the same-rank ghost exchange carries a tenth warp (the ghost warp)
        if (stage && G.ghost) return uniform ? go(flux_div_narrow_kernel<CONV, VISC, true, true, TI, TJ, false, true>, NTHREADS) : go(flux_div_narrow_kernel<CONV, VISC, false, true, TI, TJ, false, true>, NTHREADS);
        if (stage) return uniform ? go(flux_div_narrow_kernel<CONV, VISC, true, true, TI, TJ>) : go(flux_div_narrow_kernel<CONV, VISC, false, true, TI, TJ>);
        return uniform ? go(flux_div_narrow_kernel<CONV, VISC, true, false, TI, TJ>) : go(flux_div_narrow_kernel<CONV, VISC, false, false, TI, TJ>);
    }

    template <int CONV, int VISC>
    int launch_fdiv_narrow(const spb_grid* g, const double* q, double* rhs, const FluxParams& P, int increment,
                           int64_t lb_begin, int64_t lb_end, cudaStream_t stream, double* q_out, const nrw::Stage* stage,
                           spb_exchange* exch)
    {
        // blocks of up to 16 cells along i (16^3 blocks: BASELINE configs 1 and 5) would fill half of a 32-wide tile row
        if (g->nx[0] <= 16) return launch_fdiv_narrow_tile<CONV, VISC, 16, 16>(g, q, rhs, P, increment, lb_begin, lb_end, stream, q_out, stage, exch);
        // A 32 x 16 tile (16 compute warps, one CTA per SM: less halo traffic, one edge warp per 16 rows) was measured in round 2 and
        // is slower: flux_div 0.393 vs 0.372 ms at 256^3, sustained 512^3 step 21.0 vs 20.1 ms (profiles/r02_tile_32x16.log) — with a
        // single CTA per SM nothing fills the barrier bubbles. The kernel is written on Lay<TI, TJ>::NCOMP, so the variant is one line.
        return launch_fdiv_narrow_tile<CONV, VISC, 32, 8>(g, q, rhs, P, increment, lb_begin, lb_end, stream, q_out, stage, exch);
    }

    template int launch_fdiv_narrow<SPB_CONV_TOTANI, 1>(const spb_grid*, const double*, double*, const FluxParams&, int, int64_t, int64_t, cudaStream_t, double*, const nrw::Stage*, spb_exchange*);
    template int launch_fdiv_narrow<SPB_CONV_TOTANI, 0>(const spb_grid*, const double*, double*, const FluxParams&, int, int64_t, int64_t, cudaStream_t, double*, const nrw::Stage*, spb_exchange*);
    template int launch_fdiv_narrow<SPB_CONV_NONE, 1>(const spb_grid*, const double*, double*, const FluxParams&, int, int64_t, int64_t, cudaStream_t, double*, const nrw::Stage*, spb_exchange*);
}
