// (template; instantiated by spb_flux_div.cu for coords::identity and by spb_flux_div_curv.cu for general coordinates)
// Fused RHS kernel: replaces pde_algs::flux_div (reference src/pde-algs/flux-div/flux_div_basic.h:17-77
// and its shared-memory variants flux_div_fldbc.h / flux_div_ldbal.h) on sm_100a.
//
// One CTA owns a TI x TJ column of cells of one block and marches through k. Planes of the
// primitive array (reference AoS order, 5 doubles per cell) are staged into a shared-memory ring by
// TMA (cp.async.bulk.tensor.4d over the tensor [5*(n0+2g), n1+2g, n2+2g, nlb]) and signalled on
// mbarriers; the next plane is in flight while the current one is computed. Every face flux is
// evaluated once: a thread computes the lower x-, y- and z-face of its cell, x/y fluxes are
// shared through shared memory, the z flux is carried in registers to the next plane.
#pragma once
#include "spb_common.cuh"
#include "spb_tma.cuh"
#include "spb_flux.cuh"
#include <type_traits>

namespace spb
{
    constexpr int TI = 32;

    struct FdivDims
    {
        int nx[3], ng[3], np[3];
        int tiles_i, tiles_j;
        long long block_stride;
        long long lb0;
        int increment;
        int lm;                     // general coordinates: length of one metric table row (spb_grid::metric_lm)
        const int* blist;           // local block of CTA group t (a scattered block set in one launch), or null: lb0 + t
        // the distinct 1/dx of each direction (refinement levels, spb_grid::lev_inv) and the packed level indices of every block:
        // read through a warp-uniform index (REDUX) they live in uniform registers, not in six vector registers per thread
        double lev_inv[3][16];
        const int* lev;
    };

    template <int H> struct FdivSmem
    {
        // +2: fp64 TMA boxes must start on a 16-byte boundary (measured on B200: an odd first coordinate raises
        // an illegal-instruction fault), so the box starts one cell early when the halo start is an odd cell
        static constexpr int TJ = (H >= 4) ? 4 : 8;                      // tile rows: 9 ring planes of an 8-row tile with 4 halo cells exceed 227 KB
        static constexpr int NT = TI*TJ;                                  // threads per CTA
        static constexpr int TIp = TI + 2*H + 2, TJp = TJ + 2*H;
        // Step k reads planes k-H .. k+AH-H: the z-face of a 2H-cell stencil reaches H-1 planes up (AH = 2H-1), the tangential
        // differences of the x/y faces one plane up (AH = H+1). One more plane is in flight: NP = AH + 2 ring slots.
        static constexpr int AH = (H + 1 > 2*H - 1) ? H + 1 : 2*H - 1;
        static constexpr int NP = AH + 2;
        static constexpr int PLANE_DOUBLES = TIp*TJp*5;
        static constexpr int PLANE_BYTES = PLANE_DOUBLES*8;
        static constexpr int PLANE_STRIDE_BYTES = (PLANE_BYTES + 127)/128*128;
        static constexpr int PLANE_STRIDE = PLANE_STRIDE_BYTES/8;
        static constexpr int FX_DOUBLES = TJ*(TI + 1)*5;
        static constexpr int FY_DOUBLES = (TJ + 1)*TI*5;
        static constexpr int BYTES = NP*PLANE_STRIDE_BYTES + (FX_DOUBLES + FY_DOUBLES)*8 + NP*8 + 128;
        // WENO kernels (PRE): the x-fluxes travel by warp shuffle (only the tile's upper-edge column goes through shared memory) and
        // the freed space holds the per-cell WENO preparation of ONE plane, two values per cell of the tile + 2 halo cells
        static constexpr int PW = TI + 4, PH = TJ + 4, PUB_CELLS = PW*PH;
        static constexpr int FXE_DOUBLES = TJ*5;
        static constexpr int BYTES_PRE = NP*PLANE_STRIDE_BYTES + (FXE_DOUBLES + FY_DOUBLES + 2*PUB_CELLS)*8 + NP*8 + 128;
        // PRE kernels on general coordinates: the three metric rows of the tile's x and y range (cells and faces 0 .. TI / TJ)
        static constexpr int METS_ZCAP = 65;                           // staged z entries per row (blocks of up to 64 cells along k)
        static constexpr int METS_DOUBLES = 3*(TI + 1) + 3*(TJ + 1) + 3*METS_ZCAP;
    };

    // accessor of the staged planes, centred on tile-local cell (il, jl) of the current plane
    template <int H> struct TileAcc
    {
        const double* ring;
        int cell;          // ((jl+H)*TIp + (il+H))*5
        int pl[FdivSmem<H>::AH + 1];      // plane offsets (doubles) for dk = -H .. AH-H
        __device__ __forceinline__ double operator()(int v, int di, int dj, int dk) const
        {
            return ring[pl[dk + H] + cell + (dj*FdivSmem<H>::TIp + di)*5 + v];
        }
    };

    // FUSED: the RK stage update of spb_flux_div_rk_stage rides on the rhs of each finished cell (see spb_flux.cuh:
    // StageParams); rhs is then the residual register written (or null) and q_out the new state.
    // CURV: general (diagonal) coordinates, reference src/core/coord_system.h:250-267,295-302 with flux_div_basic.h:49-71:
    //   rhs(c) = J(c) sum_d (F_lower - F_upper)/dxi_d,  F = flux with the metric vector area*e_d,  J = 1/(m0 m1 m2).
    // `met` holds, per block and direction, three rows of length G.lm (spb_grid_set_metric): row 0 = m as info::metric
    // evaluates it at the cell centres, row 1 = 1/m at the computational cell centres (Jacobian, tangential gradient
    // transform), row 2 = 1/m at the faces (normal gradient transform); index = padded cell / face index.
    template <int CONV, int DISS, int VISC, bool FUSED, bool CURV = false, bool SGS = false>
    __global__ void __launch_bounds__(FdivSmem<stencil_halo<CONV, DISS>::value>::NT, 2)
    flux_div_kernel(const __grid_constant__ CUtensorMap tmap_q, double* __restrict__ rhs, const FluxParams P,
                    const FdivDims G, const double* __restrict__ inv_dx_tab, double* __restrict__ q_out, const StageParams ST,
                    const double* __restrict__ met, const int* __restrict__ nbr)
    {
        constexpr int H = stencil_halo<CONV, DISS>::value;
        using S = FdivSmem<H>;
        constexpr int TJ = S::TJ, AH = S::AH;
        extern __shared__ __align__(128) double smem_raw[];
        // pointer arithmetic on the __shared__ symbol (no integer round trip) keeps the address space known
        // to the compiler: LDS/STS instead of generic LD/ST
        double*   ring = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u)/8u;
        // PRE: per-cell WENO preparation published once per plane (see FdivSmem), x-fluxes by shuffle
        constexpr bool PRE = (CONV == SPB_CONV_FWENO) || (DISS == SPB_DISS_FWENO);
        double*   Fx   = ring + S::NP*S::PLANE_STRIDE;                           // PRE: only the upper-edge column [TJ][5]
        double*   Fy   = Fx + (PRE ? S::FXE_DOUBLES : S::FX_DOUBLES);
        double*   pubh = Fy + S::FY_DOUBLES;                                     // PRE: rho/2 of the published plane
        double*   pubs = pubh + S::PUB_CELLS;                                    // PRE: (rho/2)(|u| + c)
        // PRE + CURV: the metric rows of the tile (x, y range of the tile, z range of the block) live in shared memory. A thread's
        // own x / y entries are loop invariants: read through __ldg the compiler keeps them in registers for the whole k loop
        // (12 registers, paid in spills); shared-memory loads are not moved across the barriers of a step, so they are simply
        // re-read where a face needs them. The z entries change every step: as global loads they sat at the top of the step in
        // front of the z-face (long-scoreboard stalls 1.31 per issue against 0.47 on identity coordinates).
        constexpr bool METS = PRE && CURV;
        double*   mets = pubs + S::PUB_CELLS;
        uint64_t* bars = (uint64_t*)(PRE ? pubs + S::PUB_CELLS + (METS ? S::METS_DOUBLES : 0) : Fy + S::FY_DOUBLES);

        const int tid = threadIdx.x;
        const int il = tid & 31, jl = tid >> 5;

        int t = blockIdx.x;
        const int ti = t % G.tiles_i; t /= G.tiles_i;
        const int tj = t % G.tiles_j; t /= G.tiles_j;
        const long long lb = G.blist ? (long long)G.blist[t] : G.lb0 + t;
        const int i0 = ti*TI, j0 = tj*TJ;
        const int nz = G.nx[2];
        const int ni_t = min(TI, G.nx[0] - i0);     // interior cells of this tile along i / j
        const int nj_t = min(TJ, G.nx[1] - j0);
        const bool active = (il < ni_t) && (jl < nj_t);

        const unsigned lpk = __reduce_max_sync(0xffffffffu, (unsigned)__ldg(G.lev + lb));
        const double invdx[3] = {G.lev_inv[0][lpk & 15u], G.lev_inv[1][(lpk >> 8) & 15u], G.lev_inv[2][(lpk >> 16) & 15u]};
        // metric rows of this block: M(d, row, idx)
        const double* mt = CURV ? met + lb*9*(long long)G.lm : nullptr;
        auto M = [&](const int d, const int row, const int idx) { return __ldg(mt + (d*3 + row)*G.lm + idx); };
        // the same entry addressed by padded index; METS: x / y entries come from the staged rows (tile-local index)
        auto MP = [&](const int d, const int row, const int idx)
        {
            if (METS && d == 0) return mets[row*(TI + 1) + (idx - i0 - G.ng[0])];
            if (METS && d == 1) return mets[3*(TI + 1) + row*(TJ + 1) + (idx - j0 - G.ng[1])];
            if (METS && d == 2 && idx - G.ng[2] < S::METS_ZCAP) return mets[3*(TI + 1) + 3*(TJ + 1) + row*S::METS_ZCAP + (idx - G.ng[2])];
            return M(d, row, idx);
        };
        // gradient scales and area factor of the lower face of direction D of the padded cell (ip, jp, kp)
        auto face_metric = [&](auto Dc, const int ip, const int jp, const int kp, double (&gs)[3], double& area)
        {
            constexpr int D = decltype(Dc)::value, T1 = (D + 1) % 3, T2 = (D + 2) % 3;
            const int idx[3] = {ip, jp, kp};
            area   = MP(T1, 0, idx[T1])*MP(T2, 0, idx[T2]);
            gs[D]  = invdx[D]*MP(D, 2, idx[D]);
            gs[T1] = invdx[T1]*MP(T1, 1, idx[T1]);
            gs[T2] = invdx[T2]*MP(T2, 1, idx[T2]);
        };
        constexpr std::integral_constant<int, 0> DX{};
        constexpr std::integral_constant<int, 1> DY{};
        constexpr std::integral_constant<int, 2> DZ{};
        const int ipc = i0 + il + G.ng[0], jpc = j0 + jl + G.ng[1];       // padded indices of this thread's cell column

        // TMA coordinates of the tile (fused (v,i) dimension first)
        const int ash = (i0 + G.ng[0] - H) & 1;          // alignment shift (cells)
        const int c0 = 5*(i0 + G.ng[0] - H - ash);
        const int c1 = j0 + G.ng[1] - H;
        const int c2base = G.ng[2] - H;               // plane p -> k = p - H -> coordinate p + ng - H
        const int nplanes = nz + 2*H;

        if (tid == 0)
        {
            prefetch_tmap(&tmap_q);
            #pragma unroll
            for (int s = 0; s < S::NP; ++s) mbar_init(&bars[s], 1);
            fence_mbar_init();
        }
        if (METS)
        {
            // entry (row, loc) of direction d: padded index origin + loc, clamped to the table (ragged tiles, short blocks)
            for (int e0 = tid; e0 < S::METS_DOUBLES; e0 += blockDim.x)
            {
                const int d = e0 < 3*(TI + 1) ? 0 : (e0 < 3*(TI + 1) + 3*(TJ + 1) ? 1 : 2);
                const int e = e0 - (d == 0 ? 0 : (d == 1 ? 3*(TI + 1) : 3*(TI + 1) + 3*(TJ + 1)));
                const int len = d == 0 ? TI + 1 : (d == 1 ? TJ + 1 : S::METS_ZCAP);
                const int row = e / len, loc = e - row*len;
                const int org = d == 0 ? i0 + G.ng[0] : (d == 1 ? j0 + G.ng[1] : G.ng[2]);
                mets[e0] = M(d, row, min(org + loc, G.lm - 1));
            }
        }
        __syncthreads();
        if (tid == 0)
        {
            #pragma unroll
            for (int p = 0; p < S::NP; ++p)
            {
                if (p < nplanes)
                {
                    mbar_arrive_expect_tx(&bars[p], S::PLANE_BYTES);
                    tma_load_4d(ring + p*S::PLANE_STRIDE, &tmap_q, &bars[p], c0, c1, c2base + p, (int)lb);
                }
            }
        }

        TileAcc<H> acc;
        acc.ring = ring;
        acc.cell = ((jl + H)*S::TIp + (il + H + ash))*5;

        // planes p = 0 .. AH-1 are needed by step 0 besides plane AH
        uint32_t parity_bits = 0;                      // bit s = parity to wait for on slot s
        #pragma unroll
        for (int p = 0; p < AH; ++p) { mbar_wait(&bars[p], 0); }
        parity_bits = (1u << AH) - 1;                  // slots 0..AH-1 consumed once
        int slot_next = AH;                            // slot of plane k+AH at step k
        // plane offsets for dk = -H .. AH-H at step k=0: plane p = k + H + dk
        #pragma unroll
        for (int d = 0; d < AH; ++d) acc.pl[d] = d*S::PLANE_STRIDE;

        // PRE: weno_cell of the own column for the cells k-2 .. k+1 (the stencil of the z-face k-1/2) rolls through registers; the
        // plane the x / y faces of a step read is published in shared memory: the interior cells from these registers, the
        // 2-cell halo ring (176 cells) evaluated by the first 176 threads
        double zh[4] = {0.0, 0.0, 0.0, 0.0}, zs[4] = {0.0, 0.0, 0.0, 0.0};
        const int pown = (jl + 2)*S::PW + (il + 2);
        // The halo cells belong to the threads of warps 2 .. 7 (warps 0 and 1 carry the faces on the tile's upper edge, a fourth
        // face pass per step: everything else that can be taken off them is), and a halo cell of plane k+1 is evaluated BEFORE the
        // first barrier of step k, while its warp would otherwise wait for warps 0 and 1; only the two stores follow the barrier.
        int phalo = -1, chalo = 0;                   // published index and ring offset of this thread's halo cell
        constexpr int HALO_T0 = (S::NT - 64 >= S::PUB_CELLS - TI*TJ) ? 64 : 0;
        if (PRE && tid >= HALO_T0 && tid - HALO_T0 < S::PUB_CELLS - TI*TJ)
        {
            const int ht = tid - HALO_T0;
            int pi, pj;
            if (ht < 4*S::PW) { pj = ht / S::PW; pi = ht - pj*S::PW; if (pj >= 2) pj += TJ; }
            else { const int r = ht - 4*S::PW; pj = 2 + r/4; const int c = r & 3; pi = (c < 2) ? c : TI + c; }
            phalo = pj*S::PW + pi;
            chalo = (pj*S::TIp + (pi + ash))*5;      // H = 2: tile-local cell (pi - 2, pj - 2) sits at box column pi - 2 + H + ash
        }
        double hh = 0.0, hsh = 0.0;                  // weno_cell of this thread's halo cell, on its way to the published plane
        auto halo_eval = [&](const int dk)
        {
            if (phalo >= 0)
            {
                const double* c = ring + acc.pl[dk + H] + chalo;
                weno_cell(P, c[0], c[1], c[2], c[3], c[4], hh, hsh);
            }
        };
        auto publish = [&](const double hr_own, const double hs_own)
        {
            pubh[pown] = hr_own; pubs[pown] = hs_own;
            if (phalo >= 0) { pubh[phalo] = hh; pubs[phalo] = hsh; }
        };
        if (PRE)
        {
            #pragma unroll
            for (int d = 0; d < 3; ++d)
                weno_cell(P, acc(0, 0, 0, d - 2), acc(1, 0, 0, d - 2), acc(2, 0, 0, d - 2), acc(3, 0, 0, d - 2), acc(4, 0, 0, d - 2), zh[d], zs[d]);
            halo_eval(0);
            publish(zh[2], zs[2]);
            __syncthreads();
        }
        double rprev[5] = {0.0, 0.0, 0.0, 0.0, 0.0};   // partial rhs of cell k-1 (x, y and lower-z parts)
        const long long col0 = lb*G.block_stride
            + 5ll*((i0 + il + G.ng[0]) + (long long)G.np[0]*((j0 + jl + G.ng[1]) + (long long)G.np[1]*G.ng[2]));
        double* rhs_col = rhs + col0;
        const long long kstride = 5ll*G.np[0]*G.np[1];

        for (int k = 0; k <= nz; ++k)
        {
            // plane k + AH
            if (k + AH < nplanes)
            {
                mbar_wait(&bars[slot_next], (parity_bits >> slot_next) & 1u);
                parity_bits ^= (1u << slot_next);
            }
            acc.pl[AH] = slot_next*S::PLANE_STRIDE;

            if (FUSED && active && k < nz)
            {
                // the stage inputs of cell k are read at the top of the next step (40 bytes per thread, AoS): start them towards L1 now
                #pragma unroll
                for (int a = 0; a < 2; ++a)
                    if (a < ST.nin)
                    {
                        const double* pin = ST.in[a] + col0 + (long long)k*kstride;
                        asm volatile("prefetch.global.L1 [%0];" :: "l"(pin));
                        asm volatile("prefetch.global.L1 [%0];" :: "l"(pin + 4));
                    }
            }
            double Fz[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            double jac_prev = 1.0;                   // Jacobian of cell k-1
            if (PRE && k + AH < nplanes)             // the cell entering the z-stencil (plane k+1)
                weno_cell(P, acc(0, 0, 0, 1), acc(1, 0, 0, 1), acc(2, 0, 0, 1), acc(3, 0, 0, 1), acc(4, 0, 0, 1), zh[3], zs[3]);
            if (active)
            {
                if (CURV)
                {
                    double gs[3], area;
                    face_metric(DZ, ipc, jpc, k + G.ng[2], gs, area);
                    face_flux<CONV, DISS, VISC, 2, true, SGS, PRE>(acc, P, gs, Fz, area, zh, zs);
                    if (k >= 1) jac_prev = MP(0, 1, ipc)*MP(1, 1, jpc)*MP(2, 1, k - 1 + G.ng[2]);
                }
                else face_flux<CONV, DISS, VISC, 2, false, SGS, PRE>(acc, P, invdx, Fz, 1.0, zh, zs);
            }

            if (k >= 1 && active)
            {
                double* o = rhs_col + (long long)(k - 1)*kstride;
                if (!FUSED)
                {
                    #pragma unroll
                    for (int v = 0; v < 5; ++v)
                    {
                        double r = fma(-Fz[v], invdx[2], rprev[v]);
                        if (CURV) r *= jac_prev;
                        if (G.increment) r += o[v];
                        o[v] = r;
                    }
                }
                else
                {
                    const long long c = col0 + (long long)(k - 1)*kstride;
                    double w[5], ov[5];
                    #pragma unroll
                    for (int v = 0; v < 5; ++v)
                    {
                        double r = fma(-Fz[v], invdx[2], rprev[v]);
                        if (CURV) r *= jac_prev;
                        w[v] = ST.cq_self*r; ov[v] = ST.co_self*r;
                    }
                    #pragma unroll
                    for (int a = 0; a < 2; ++a)
                        if (a < ST.nin)
                        {
                            #pragma unroll
                            for (int v = 0; v < 5; ++v)
                            {
                                const double x = ST.in[a][c + v];
                                w[v] = fma(ST.cq[a], x, w[v]); ov[v] = fma(ST.co[a], x, ov[v]);
                            }
                        }
                    if (ST.has_out)
                    {
                        #pragma unroll
                        for (int v = 0; v < 5; ++v) o[v] = ov[v];
                    }
                    // prim -> cons (fluid_state.h:103-116), add the increment, cons -> prim (fluid_state.h:119-135)
                    double qc[5];
                    #pragma unroll
                    for (int v = 0; v < 5; ++v) qc[v] = acc(v, 0, 0, -1);
                    const double rho0 = qc[0]*rcp_nr(P.R*qc[1]);
                    const double u2 = fma(qc[2], qc[2], fma(qc[3], qc[3], qc[4]*qc[4]));
                    const double rho  = rho0 + w[0];
                    const double rhoE = fma(0.5*rho0, u2, qc[0]*ST.inv_gm1) + w[1];
                    const double mx = fma(rho0, qc[2], w[2]), my = fma(rho0, qc[3], w[3]), mz = fma(rho0, qc[4], w[4]);
                    const double ir = rcp_nr(rho);
                    const double un = ir*mx, vn = ir*my, wn = ir*mz;
                    const double pn = ST.gm1*fma(-0.5*rho, fma(un, un, fma(vn, vn, wn*wn)), rhoE);
                    const double Tn = pn*ir*ST.inv_R;
                    double* qo = q_out + c;
                    qo[0] = pn; qo[1] = Tn; qo[2] = un; qo[3] = vn; qo[4] = wn;
                    if (nbr)
                    {
                        // Fused same-rank ghost exchange (make_exchange.h:166-203): a cell of the source box of direction e goes to
                        // cell (i - ex n0, j - ey n1, k - ez n2) of neighbour e (get_transaction.h:54-86). Only the cells of the
                        // block's outer shell take part; nbr[27 lb + e] is the destination block or -1 (spb_exchange.cu).
                        const int ci = i0 + il, cj = j0 + jl, ck = k - 1;
                        const bool xl = ci < G.ng[0], xh = ci >= G.nx[0] - G.ng[0];
                        const bool yl = cj < G.ng[1], yh = cj >= G.nx[1] - G.ng[1];
                        const bool zl = ck < G.ng[2], zh = ck >= G.nx[2] - G.ng[2];
                        if (xl || xh || yl || yh || zl || zh)
                        {
                            const int* nb = nbr + 27*lb;
                            // the directions a cell takes part in form a contiguous range per axis: {-1 if low} + {0} + {+1 if high}
                            for (int ez = zl ? -1 : 0; ez <= (zh ? 1 : 0); ++ez)
                            {
                                for (int ey = yl ? -1 : 0; ey <= (yh ? 1 : 0); ++ey)
                                {
                                    for (int ex = xl ? -1 : 0; ex <= (xh ? 1 : 0); ++ex)
                                    {
                                        if (ex == 0 && ey == 0 && ez == 0) continue;
                                        const int dst = nb[(ex + 1) + 3*(ey + 1) + 9*(ez + 1)];
                                        if (dst < 0) continue;
                                        double* o = q_out + dst*G.block_stride
                                            + 5ll*((ci - ex*G.nx[0] + G.ng[0]) + (long long)G.np[0]*((cj - ey*G.nx[1] + G.ng[1])
                                            + (long long)G.np[1]*(ck - ez*G.nx[2] + G.ng[2])));
                                        o[0] = pn; o[1] = Tn; o[2] = un; o[3] = vn; o[4] = wn;
                                    }
                                }
                            }
                        }
                    }
                }
            }

            // PRE: the z part of the divergence of cell k, then the x part as soon as the x pass is done (F_x(own) - F_x(il+1) by warp
            // shuffle for all but the tile's last column), so that only five accumulators stay live through the y and edge passes
            double rz[5];
            #pragma unroll
            for (int v = 0; v < 5; ++v) rz[v] = Fz[v]*invdx[2];
            if (k < nz)
            {
                double Fxo[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
                if (active)
                {
                    double F[5], gs[3], area;
                    double h4[4], s4[4];
                    if (PRE)
                    {
                        #pragma unroll
                        for (int i = 0; i < 4; ++i) { h4[i] = pubh[pown - 2 + i]; s4[i] = pubs[pown - 2 + i]; }
                    }
                    if (CURV) { face_metric(DX, ipc, jpc, k + G.ng[2], gs, area); face_flux<CONV, DISS, VISC, 0, true, SGS, PRE>(acc, P, gs, F, area, h4, s4); }
                    else face_flux<CONV, DISS, VISC, 0, false, SGS, PRE>(acc, P, invdx, F, 1.0, h4, s4);
                    if (PRE)
                    {
                        #pragma unroll
                        for (int v = 0; v < 5; ++v) Fxo[v] = F[v];
                    }
                    else
                    {
                        #pragma unroll
                        for (int v = 0; v < 5; ++v) Fx[(jl*(TI + 1) + il)*5 + v] = F[v];
                    }
                }
                if (PRE)
                {
                    // the x-flux of the right neighbour comes by shuffle (all 32 lanes take part; a warp is one tile row); the last
                    // column subtracts the edge flux after the barrier
                    #pragma unroll
                    for (int v = 0; v < 5; ++v)
                    {
                        const double up = __shfl_down_sync(0xffffffffu, Fxo[v], 1);
                        rz[v] = fma((il == ni_t - 1) ? Fxo[v] : Fxo[v] - up, invdx[0], rz[v]);
                    }
                }
                if (active)
                {
                    double F[5], gs[3], area;
                    double h4[4], s4[4];
                    if (PRE)
                    {
                        #pragma unroll
                        for (int i = 0; i < 4; ++i) { h4[i] = pubh[pown + (i - 2)*S::PW]; s4[i] = pubs[pown + (i - 2)*S::PW]; }
                    }
                    if (CURV) { face_metric(DY, ipc, jpc, k + G.ng[2], gs, area); face_flux<CONV, DISS, VISC, 1, true, SGS, PRE>(acc, P, gs, F, area, h4, s4); }
                    else face_flux<CONV, DISS, VISC, 1, false, SGS, PRE>(acc, P, invdx, F, 1.0, h4, s4);
                    #pragma unroll
                    for (int v = 0; v < 5; ++v) Fy[(jl*TI + il)*5 + v] = F[v];
                }
                // faces on the upper edge of the tile
                if (tid < nj_t)                      // warp 0: x-face i = ni_t of row tid
                {
                    TileAcc<H> e = acc;
                    e.cell = ((tid + H)*S::TIp + (ni_t + H + ash))*5;
                    double F[5], gs[3], area;
                    double h4[4], s4[4];
                    if (PRE)
                    {
                        const int pe = (tid + 2)*S::PW + (ni_t + 2);
                        #pragma unroll
                        for (int i = 0; i < 4; ++i) { h4[i] = pubh[pe - 2 + i]; s4[i] = pubs[pe - 2 + i]; }
                    }
                    if (CURV) { face_metric(DX, i0 + ni_t + G.ng[0], j0 + tid + G.ng[1], k + G.ng[2], gs, area); face_flux<CONV, DISS, VISC, 0, true, SGS, PRE>(e, P, gs, F, area, h4, s4); }
                    else face_flux<CONV, DISS, VISC, 0, false, SGS, PRE>(e, P, invdx, F, 1.0, h4, s4);
                    #pragma unroll
                    for (int v = 0; v < 5; ++v) Fx[PRE ? tid*5 + v : (tid*(TI + 1) + ni_t)*5 + v] = F[v];
                }
                if (tid >= 32 && tid < 32 + ni_t)    // warp 1: y-face j = nj_t of column tid-32
                {
                    TileAcc<H> e = acc;
                    e.cell = ((nj_t + H)*S::TIp + (tid - 32 + H + ash))*5;
                    double F[5], gs[3], area;
                    double h4[4], s4[4];
                    if (PRE)
                    {
                        const int pe = (nj_t + 2)*S::PW + (tid - 32 + 2);
                        #pragma unroll
                        for (int i = 0; i < 4; ++i) { h4[i] = pubh[pe + (i - 2)*S::PW]; s4[i] = pubs[pe + (i - 2)*S::PW]; }
                    }
                    if (CURV) { face_metric(DY, i0 + tid - 32 + G.ng[0], j0 + nj_t + G.ng[1], k + G.ng[2], gs, area); face_flux<CONV, DISS, VISC, 1, true, SGS, PRE>(e, P, gs, F, area, h4, s4); }
                    else face_flux<CONV, DISS, VISC, 1, false, SGS, PRE>(e, P, invdx, F, 1.0, h4, s4);
                    #pragma unroll
                    for (int v = 0; v < 5; ++v) Fy[(nj_t*TI + (tid - 32))*5 + v] = F[v];
                }
            }
            // the halo cells of the plane published next (k+1), evaluated by warps 2 .. 7 while warps 0 and 1 finish the edge faces
            if (PRE && k + 1 < nz) halo_eval(1);
            __syncthreads();
            if (k < nz && active)
            {
                #pragma unroll
                for (int v = 0; v < 5; ++v)
                {
                    const double dFy = Fy[(jl*TI + il)*5 + v] - Fy[((jl + 1)*TI + il)*5 + v];
                    double r = fma(dFy, invdx[1], rz[v]);
                    if (PRE) { if (il == ni_t - 1) r = fma(-Fx[jl*5 + v], invdx[0], r); }
                    else     r = fma(Fx[(jl*(TI + 1) + il)*5 + v] - Fx[(jl*(TI + 1) + il + 1)*5 + v], invdx[0], r);
                    rprev[v] = r;
                }
            }
            // PRE: the x / y faces of this step have read the published plane (barrier above): publish plane k+1 for the next step
            if (PRE && k + 1 < nz) publish(zh[3], zs[3]);
            __syncthreads();
            if (PRE)
            {
                #pragma unroll
                for (int d = 0; d < 3; ++d) { zh[d] = zh[d + 1]; zs[d] = zs[d + 1]; }
            }
            // slot of plane p = k (dk = -H) is free now: refill it with plane p + NP
            if (tid == 0)
            {
                const int pnew = k + S::NP;
                if (pnew < nplanes)
                {
                    const int s = k % S::NP;
                    mbar_arrive_expect_tx(&bars[s], S::PLANE_BYTES);
                    tma_load_4d(ring + s*S::PLANE_STRIDE, &tmap_q, &bars[s], c0, c1, c2base + pnew, (int)lb);
                }
            }
            #pragma unroll
            for (int d = 0; d < AH; ++d) acc.pl[d] = acc.pl[d + 1];
            slot_next = (slot_next + 1 == S::NP) ? 0 : slot_next + 1;
        }
    }

    template <int CONV, int DISS, int VISC, bool FUSED = false, bool CURV = false, bool SGS = false>
    static int launch_fdiv(const spb_grid* g, const double* q, double* rhs, const FluxParams& P, int increment,
                           int64_t lb_begin, int64_t lb_end, cudaStream_t stream, double* q_out = nullptr, const StageParams* stage = nullptr,
                           spb_exchange* exch = nullptr)
    {
        constexpr int H = stencil_halo<CONV, DISS>::value;
        using S = FdivSmem<H>;
        for (int d = 0; d < 3; ++d)
            if (g->ng[d] < H) { set_error("spb_flux_div: scheme needs " + std::to_string(H) + " exchange cells"); return SPB_ERR_BAD_ARG; }
        if ((5*g->np[0]) % 2 != 0) { set_error("spb_flux_div: n0 + 2*g0 must be even (16-byte TMA row pitch)"); return SPB_ERR_UNSUPPORTED; }
        encode_tiled_fn enc = get_encode_tiled();
        if (!enc) { set_error("spb_flux_div: cuTensorMapEncodeTiled not available from the driver"); return SPB_ERR_DRIVER; }

        CUtensorMap tq;
        const cuuint64_t dims[4]    = {(cuuint64_t)5*g->np[0], (cuuint64_t)g->np[1], (cuuint64_t)g->np[2], (cuuint64_t)g->nlb};
        const cuuint64_t strides[3] = {(cuuint64_t)40*g->np[0], (cuuint64_t)40*g->np[0]*g->np[1], (cuuint64_t)8*g->block_stride};
        const cuuint32_t box[4]     = {(cuuint32_t)(5*S::TIp), (cuuint32_t)S::TJp, 1, 1};
        const cuuint32_t estr[4]    = {1, 1, 1, 1};
        CUresult cr = enc(&tq, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)q, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { set_error("spb_flux_div: cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)cr)); return SPB_ERR_DRIVER; }

        FdivDims G;
        for (int d = 0; d < 3; ++d) { G.nx[d] = g->nx[d]; G.ng[d] = g->ng[d]; G.np[d] = g->np[d]; }
        G.tiles_i = (g->nx[0] + TI - 1)/TI;
        G.tiles_j = (g->nx[1] + S::TJ - 1)/S::TJ;
        G.block_stride = g->block_stride;
        G.lb0 = lb_begin;
        G.increment = increment;
        G.lm = g->metric_lm;
        if (g->lev_n[0] < 0 || g->lev_n[1] < 0 || g->lev_n[2] < 0)
        { set_error("spb_flux_div: more than 16 distinct block spacings along one direction"); return SPB_ERR_UNSUPPORTED; }
        G.lev = g->lev_dev;
        for (int d = 0; d < 3; ++d)
            for (int l = 0; l < 16; ++l) G.lev_inv[d][l] = g->lev_inv[d][l];
        // fused same-rank ghost exchange: the neighbour table of the plan (refused for plans with non-canonical transactions)
        const int* nbr_tab = nullptr;
        if (FUSED && exch) { int rc = exchange_fuse_table(exch, g->nx, g->ng, g->nlb, &nbr_tab); if (rc) return rc; }
        if (CURV && !g->metric_dev) { set_error("spb_flux_div: general-coordinate kernel without a metric (spb_grid_set_metric)"); return SPB_ERR_BAD_ARG; }
        const BlockList& bl = current_block_list();
        G.blist = bl.dev;
        const int64_t nblk = (bl.dev ? bl.count : lb_end - lb_begin)*G.tiles_i*G.tiles_j;
        if (nblk <= 0) return 0;
        auto kern = flux_div_kernel<CONV, DISS, VISC, FUSED, CURV, SGS>;
        StageParams SP{};
        if (stage) SP = *stage;
        constexpr bool PRE = (CONV == SPB_CONV_FWENO) || (DISS == SPB_DISS_FWENO);
        constexpr int SMEM = PRE ? S::BYTES_PRE + (CURV ? S::METS_DOUBLES*8 : 0) : S::BYTES;
        SPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        kern<<<(unsigned)nblk, S::NT, SMEM, stream>>>(tq, rhs, P, G, g->inv_dx_dev, q_out, SP, g->metric_dev, nbr_tab);
        SPB_LAUNCH_CHECK();
        return 0;
    }

    static inline FluxParams make_params(const spb_flux_desc* f)
    {
        FluxParams P;
        P.gamma = f->gamma; P.R = f->R; P.gm1 = f->gamma - 1.0; P.cv = f->R/(f->gamma - 1.0); P.inv_gm1 = 1.0/(f->gamma - 1.0);
        P.mu = f->mu; P.beta = f->beta; P.two_mu = 2.0*f->mu;
        // reference viscous.h:64-68: cond = (gamma R/(gamma-1)) * (mu * prandtl_inv)
        P.kappa = (f->gamma*f->R/(f->gamma - 1.0))*(f->mu*f->prandtl_inv);
        P.eps = f->sensor_eps; P.blend = f->blend; P.weno_linear = f->weno_linear ? 1 : 0;
        // wale_t (subgrid_scale.h:86-89): mu_t = rho cw cw delta delta (...); sgs_visc_t: alpha += mu_t/Pr_t (viscous_laws.h:192-195)
        P.sgs_c = f->sgs_cw*f->sgs_cw*f->sgs_delta*f->sgs_delta;
        P.sgs_cp_prt = f->sgs ? (f->gamma*f->R/(f->gamma - 1.0))/f->sgs_prt : 0.0;
        return P;
    }

    // LES closure visc_lr<sgs_visc_t<constant_viscosity_t, wale_t>> (spb_flux_div_sgs.cu): identity and general coordinates
    int flux_div_sgs(const spb_grid* g, const double* q, double* rhs, const spb_flux_desc* f, const FluxParams& P, int increment,
                     int64_t lb_begin, int64_t lb_end, cudaStream_t stream, double* q_out, const StageParams* stage, spb_exchange* exch);
    // general coordinates (spb_flux_div_curv.cu): every functor combination through the wide kernel with CURV = true
    int flux_div_curv(const spb_grid* g, const double* q, double* rhs, const spb_flux_desc* f, const FluxParams& P, int increment,
                      int64_t lb_begin, int64_t lb_end, cudaStream_t stream, double* q_out, const StageParams* stage, spb_exchange* exch);
}
