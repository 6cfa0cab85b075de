// Reductions over interior cells: replaces algs::transform_reduce + destructive_reduce
// (reference src/algs/transform_reduce.h:53-191, src/algs/destructive_reduce.h:11-74).
// One launch: per-thread strided accumulation, warp-shuffle tree, one partial per CTA, and the
// last CTA to finish folds the partials in a fixed order (deterministic for sums too).
#include "spb_common.cuh"

namespace spb
{
    struct RedDims { int nx[3], ng[3], np[3]; long long ncells; };

    // max keeps a NaN once it has seen one (fmax would drop it and a diverged field would report a finite CFL speed)
    template <int OP> __device__ __forceinline__ double red_op(double a, double b) { return OP == SPB_RED_MAX ? ((a < b || b != b) ? b : a) : a + b; }
    template <int OP> __device__ __forceinline__ double red_identity() { return OP == SPB_RED_MAX ? -1.7976931348623157e308 : 0.0; }

    template <int FN> __device__ __forceinline__ double red_fn(const double* __restrict__ q, int ivar, double gamma, double R)
    {
        if (FN == SPB_FN_WAVESPEED) return sqrt(gamma*R*q[1]) + sqrt(q[2]*q[2] + q[3]*q[3] + q[4]*q[4]);
        if (FN == SPB_FN_VAR)       return q[ivar];
        if (FN == SPB_FN_ABSVAR)    return fabs(q[ivar]);
        const double rho = q[0]/(R*q[1]);
        return 0.5*rho*(q[2]*q[2] + q[3]*q[3] + q[4]*q[4]);
    }

    template <int OP, int FN>
    __global__ void __launch_bounds__(256) reduce_kernel(const double* __restrict__ q, const RedDims G, int ivar, double gamma, double R,
                                                         double* __restrict__ partials, unsigned int* __restrict__ counter, double* __restrict__ result)
    {
        // four cells per thread and trip with independent accumulators: 20 loads in flight instead of 5 (the fold order is fixed by the
        // launch shape, so sums stay deterministic)
        double a4[4] = {red_identity<OP>(), red_identity<OP>(), red_identity<OP>(), red_identity<OP>()};
        const long long stride = (long long)gridDim.x*blockDim.x;
        auto offset = [&](const long long cell)
        {
            const int i = (int)(cell % G.nx[0]); long long t = cell / G.nx[0];
            const int j = (int)(t % G.nx[1]); t /= G.nx[1];
            const int k = (int)(t % G.nx[2]); const long long lb = t / G.nx[2];
            return 5ll*((i + G.ng[0]) + (long long)G.np[0]*((j + G.ng[1]) + (long long)G.np[1]*((k + G.ng[2]) + (long long)G.np[2]*lb)));
        };
        long long cell = (long long)blockIdx.x*blockDim.x + threadIdx.x;
        for (; cell + 3*stride < G.ncells; cell += 4*stride)
        {
            #pragma unroll
            for (int u = 0; u < 4; ++u) a4[u] = red_op<OP>(a4[u], red_fn<FN>(q + offset(cell + u*stride), ivar, gamma, R));
        }
        for (; cell < G.ncells; cell += stride) a4[0] = red_op<OP>(a4[0], red_fn<FN>(q + offset(cell), ivar, gamma, R));
        double acc = red_op<OP>(red_op<OP>(a4[0], a4[1]), red_op<OP>(a4[2], a4[3]));
        #pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc = red_op<OP>(acc, __shfl_down_sync(0xffffffffu, acc, s));
        __shared__ double warp_part[8];
        __shared__ bool is_last;
        if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0)
        {
            double b = warp_part[0];
            for (int w = 1; w < 8; ++w) b = red_op<OP>(b, warp_part[w]);
            partials[blockIdx.x] = b;
            __threadfence();
            const unsigned done = atomicAdd(counter, 1u);
            is_last = (done == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last && threadIdx.x < 32)
        {
            __threadfence();
            double r = red_identity<OP>();
            // fixed order: lane l folds partials l, l+32, ... then a shuffle tree
            for (unsigned b = threadIdx.x; b < gridDim.x; b += 32) r = red_op<OP>(r, ((volatile double*)partials)[b]);
            #pragma unroll
            for (int s = 16; s > 0; s >>= 1) r = red_op<OP>(r, __shfl_down_sync(0xffffffffu, r, s));
            if (threadIdx.x == 0) { *result = r; *counter = 0u; }
        }
    }
}

extern "C" int spb_reduce(const spb_grid* g, const double* q_dev, int op, int fn, int ivar, double gamma, double R,
                          double* out_host, void* stream)
{
    using namespace spb;
    if (!g || !q_dev || !out_host || ivar < 0 || ivar > 4) { set_error("spb_reduce: bad argument"); return SPB_ERR_BAD_ARG; }
    RedDims G;
    for (int d = 0; d < 3; ++d) { G.nx[d] = g->nx[d]; G.ng[d] = g->ng[d]; G.np[d] = g->np[d]; }
    G.ncells = (long long)g->nx[0]*g->nx[1]*g->nx[2]*g->nlb;
    long long nb = (G.ncells + 255)/256;
    const long long cap = (long long)g->num_sms*8;
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (g->red_cap < nb + 2)                         // partials[nb] | result | counter, kept in the grid handle
    {
        if (g->red_scratch) { SPB_CUDA(cudaFree(g->red_scratch)); g->red_scratch = nullptr; g->red_cap = 0; }
        SPB_CUDA(cudaMalloc((void**)&g->red_scratch, sizeof(double)*(nb + 2)));
        g->red_cap = nb + 2;
    }
    double* scratch = g->red_scratch;
    double* result = scratch + nb;
    unsigned int* counter = (unsigned int*)(scratch + nb + 1);
    SPB_CUDA(cudaMemsetAsync(counter, 0, sizeof(double), st));
#define SPB_RCASE(O, F) if (op == O && fn == F) reduce_kernel<O, F><<<(unsigned)nb, 256, 0, st>>>(q_dev, G, ivar, gamma, R, scratch, counter, result)
    SPB_RCASE(SPB_RED_MAX, SPB_FN_WAVESPEED); else SPB_RCASE(SPB_RED_MAX, SPB_FN_VAR); else SPB_RCASE(SPB_RED_MAX, SPB_FN_ABSVAR);
    else SPB_RCASE(SPB_RED_MAX, SPB_FN_KINETIC); else SPB_RCASE(SPB_RED_SUM, SPB_FN_WAVESPEED); else SPB_RCASE(SPB_RED_SUM, SPB_FN_VAR);
    else SPB_RCASE(SPB_RED_SUM, SPB_FN_ABSVAR); else SPB_RCASE(SPB_RED_SUM, SPB_FN_KINETIC);
    else { set_error("spb_reduce: unknown op/fn"); return SPB_ERR_BAD_ARG; }
#undef SPB_RCASE
    SPB_LAUNCH_CHECK();
    SPB_CUDA(cudaMemcpyAsync(out_host, result, sizeof(double), cudaMemcpyDeviceToHost, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    return 0;
}
