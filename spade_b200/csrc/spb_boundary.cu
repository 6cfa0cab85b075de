// Domain-boundary ghost fill and cell source terms: the other half of a channel run's callbacks.
// Replaces algs::boundary_fill (reference src/grid/boundary_fill.h:32-133) and pde_algs::source_term
// (reference src/pde-algs/source_term.h:25-51) for the closed kernel sets of spb_bc_desc / spb_source_desc.
// Pure streaming kernels; the arithmetic uses explicit round-to-nearest multiplies and adds (no FMA
// contraction) so that the ghost values are bit-identical to the reference's CPU path.
#include "spb_common.cuh"
#include <cstring>

namespace spb
{
    struct BcDims { int nx[3], ng[3], np[3]; long long block_stride; };

    template <int KIND>
    __global__ void __launch_bounds__(256) boundary_fill_kernel(double* __restrict__ q, const BcDims G, const int64_t* __restrict__ blocks,
                                                                const long long cells_per_block, const long long ntot,
                                                                const int idir, const int pm, const spb_bc_desc bc)
    {
        const long long id = (long long)blockIdx.x*blockDim.x + threadIdx.x;
        if (id >= ntot) return;
        const long long b = id / cells_per_block;
        long long r = id - b*cells_per_block;
        // extents of the ghost slab: ng along idir, the whole padded range along the other two (boundary_fill.h:49-59)
        int ext[3], lo[3];
        #pragma unroll
        for (int d = 0; d < 3; ++d) { ext[d] = G.np[d]; lo[d] = -G.ng[d]; }
        ext[idir] = G.ng[idir];
        if (pm) lo[idir] = G.nx[idir];
        int idx[3];
        idx[0] = lo[0] + (int)(r % ext[0]); r /= ext[0];
        idx[1] = lo[1] + (int)(r % ext[1]); r /= ext[1];
        idx[2] = lo[2] + (int)r;
        const long long lb = blocks[b];
        auto off = [&](int i, int j, int k)
        { return lb*G.block_stride + 5ll*((i + G.ng[0]) + (long long)G.np[0]*((j + G.ng[1]) + (long long)G.np[1]*(k + G.ng[2]))); };
        double* fill = q + off(idx[0], idx[1], idx[2]);
        double ghost[5];
        if (KIND == SPB_BC_EXTRAP)
        {
            // boundary_fill.h:66-100: Lagrange extrapolation through the order+1 interior cells next to the face
            #pragma unroll
            for (int v = 0; v < 5; ++v) ghost[v] = 0.0;
            const int i0 = pm*(G.nx[idir] - bc.order - 1);
            for (int jj = 0; jj <= bc.order; ++jj)
            {
                double lj = 1.0;
                for (int ii = 0; ii <= bc.order; ++ii)
                    if (ii != jj)
                    {
                        const int x = idx[idir], xj = i0 + jj, xi = i0 + ii;
                        lj = __ddiv_rn(__dmul_rn(lj, (double)(x - xi)), (double)(xj - xi));
                    }
                int src[3] = {idx[0], idx[1], idx[2]};
                src[idir] = i0 + jj;
                const double* s = q + off(src[0], src[1], src[2]);
                #pragma unroll
                for (int v = 0; v < 5; ++v) ghost[v] = __dadd_rn(ghost[v], __dmul_rn(lj, s[v]));
            }
        }
        else
        {
            // boundary_fill.h:104-129: the mirror image cell through the boundary face, then kern(domain_val, idir)
            int img[3] = {idx[0], idx[1], idx[2]};
            img[idir] = pm ? 2*G.nx[idir] - (idx[idir] + 1) : -1 - idx[idir];
            const double* s = q + off(img[0], img[1], img[2]);
            #pragma unroll
            for (int v = 0; v < 5; ++v) ghost[v] = __dadd_rn(__dmul_rn(bc.a[v], s[v]), bc.b[v]);
            if (bc.use_normal) ghost[2 + idir] = __dadd_rn(__dmul_rn(bc.a_normal, s[2 + idir]), bc.b[2 + idir]);
        }
        #pragma unroll
        for (int v = 0; v < 5; ++v) fill[v] = ghost[v];
    }

    __global__ void __launch_bounds__(256) source_term_kernel(const double* __restrict__ q, double* __restrict__ rhs, const BcDims G,
                                                              const long long ncells, const spb_source_desc sd,
                                                              const double* __restrict__ met, const int lm)
    {
        const long long stride = (long long)gridDim.x*blockDim.x;
        for (long long cell = (long long)blockIdx.x*blockDim.x + threadIdx.x; cell < ncells; cell += stride)
        {
            const int i = (int)(cell % G.nx[0]); long long t = cell / G.nx[0];
            const int j = (int)(t % G.nx[1]); t /= G.nx[1];
            const int k = (int)(t % G.nx[2]); const long long lb = t / G.nx[2];
            const long long o = lb*G.block_stride + 5ll*((i + G.ng[0]) + (long long)G.np[0]*((j + G.ng[1]) + (long long)G.np[1]*(k + G.ng[2])));
            double S[5];
            if (sd.kind == SPB_SRC_BODY_FORCE)
            {
                // S = (0, f.u, fx, fy, fz); source_term.h:47-48: rhs += S/jac with jac = 1
                S[0] = 0.0;
                S[1] = __dadd_rn(__dadd_rn(__dmul_rn(sd.f[0], q[o + 2]), __dmul_rn(sd.f[1], q[o + 3])), __dmul_rn(sd.f[2], q[o + 4]));
                S[2] = sd.f[0]; S[3] = sd.f[1]; S[4] = sd.f[2];
            }
            else
            {
                #pragma unroll
                for (int v = 0; v < 5; ++v) S[v] = sd.f[v];
            }
            if (met)
            {
                // general coordinates: rhs += S/jac, jac = 1/(m0 m1 m2) (source_term.h:38-46); row 1 of the tables holds 1/m
                const double* mt = met + lb*9*(long long)lm;
                const double jac = mt[lm + i + G.ng[0]]*mt[4*lm + j + G.ng[1]]*mt[7*lm + k + G.ng[2]];
                #pragma unroll
                for (int v = 0; v < 5; ++v) S[v] = __ddiv_rn(S[v], jac);
            }
            #pragma unroll
            for (int v = 0; v < 5; ++v) rhs[o + v] = __dadd_rn(rhs[o + v], S[v]);
        }
    }

    static BcDims make_dims(const spb_grid* g)
    {
        BcDims G;
        for (int d = 0; d < 3; ++d) { G.nx[d] = g->nx[d]; G.ng[d] = g->ng[d]; G.np[d] = g->np[d]; }
        G.block_stride = g->block_stride;
        return G;
    }
}

extern "C"
{
    int spb_boundary_fill(const spb_grid* g, double* q_dev, int idir, int pm, const int64_t* blocks_host, int64_t nblocks,
                          const spb_bc_desc* bc, void* stream)
    {
        using namespace spb;
        if (!g || !q_dev || !bc || idir < 0 || idir > 2 || pm < 0 || pm > 1 || nblocks < 0 || (nblocks > 0 && !blocks_host))
        { set_error("spb_boundary_fill: bad argument"); return SPB_ERR_BAD_ARG; }
        if (bc->kind != SPB_BC_MIRROR && bc->kind != SPB_BC_EXTRAP) { set_error("spb_boundary_fill: unknown kernel kind"); return SPB_ERR_BAD_ARG; }
        if (bc->kind == SPB_BC_EXTRAP && (bc->order < 0 || bc->order + 1 > g->nx[idir])) { set_error("spb_boundary_fill: extrapolation order does not fit the block"); return SPB_ERR_BAD_ARG; }
        if (nblocks == 0 || g->ng[idir] == 0) return 0;
        for (int64_t b = 0; b < nblocks; ++b)
            if (blocks_host[b] < 0 || blocks_host[b] >= g->nlb) { set_error("spb_boundary_fill: block id out of range"); return SPB_ERR_BAD_ARG; }
        // the block list of each of the six boundaries is kept on the device (re-uploaded only when it changes)
        const int ib = 2*idir + pm;
        std::vector<int64_t>& cache = g->bnd_blocks_host[ib];
        if (cache.size() != (size_t)nblocks || std::memcmp(cache.data(), blocks_host, sizeof(int64_t)*nblocks) != 0)
        {
            if (g->bnd_blocks_dev[ib]) { SPB_CUDA(cudaFree(g->bnd_blocks_dev[ib])); g->bnd_blocks_dev[ib] = nullptr; }
            SPB_CUDA(cudaMalloc((void**)&g->bnd_blocks_dev[ib], sizeof(int64_t)*nblocks));
            SPB_CUDA(cudaMemcpy(g->bnd_blocks_dev[ib], blocks_host, sizeof(int64_t)*nblocks, cudaMemcpyHostToDevice));
            cache.assign(blocks_host, blocks_host + nblocks);
        }
        const BcDims G = make_dims(g);
        long long cpb = g->ng[idir];
        for (int d = 0; d < 3; ++d) if (d != idir) cpb *= g->np[d];
        const long long ntot = cpb*nblocks;
        const unsigned nb = (unsigned)((ntot + 255)/256);
        cudaStream_t st = (cudaStream_t)stream;
        if (bc->kind == SPB_BC_EXTRAP) boundary_fill_kernel<SPB_BC_EXTRAP><<<nb, 256, 0, st>>>(q_dev, G, g->bnd_blocks_dev[ib], cpb, ntot, idir, pm, *bc);
        else                           boundary_fill_kernel<SPB_BC_MIRROR><<<nb, 256, 0, st>>>(q_dev, G, g->bnd_blocks_dev[ib], cpb, ntot, idir, pm, *bc);
        SPB_LAUNCH_CHECK();
        return 0;
    }

    int spb_source_term(const spb_grid* g, const double* q_dev, double* rhs_dev, const spb_source_desc* src, void* stream)
    {
        using namespace spb;
        if (!g || !q_dev || !rhs_dev || !src) { set_error("spb_source_term: null argument"); return SPB_ERR_BAD_ARG; }
        if (src->kind != SPB_SRC_BODY_FORCE && src->kind != SPB_SRC_CONSTANT) { set_error("spb_source_term: unknown kernel kind"); return SPB_ERR_BAD_ARG; }
        const long long ncells = (long long)g->nx[0]*g->nx[1]*g->nx[2]*g->nlb;
        if (ncells == 0) return 0;
        long long nb = (ncells + 255)/256;
        const long long cap = (long long)g->num_sms*16;
        if (nb > cap) nb = cap;
        source_term_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(q_dev, rhs_dev, make_dims(g), ncells, *src, g->metric_dev, g->metric_lm);
        SPB_LAUNCH_CHECK();
        return 0;
    }
}
