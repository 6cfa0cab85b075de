#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../spade_b200/csrc/spb_tma.cuh"
using namespace spb;
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_1d(void* smem_dst, const void* g, uint64_t* bar, int bytes)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(smem_u32(smem_dst)), "l"(g), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int MODE>
__global__ void k(const __grid_constant__ CUtensorMap tm, const CUtensorMap* tmg, const float* src, float* out, int nel)
{
    extern __shared__ __align__(128) unsigned char sm[];
    float* buf = (float*)sm;
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, nel*4);
        if (MODE == 0) tma_load_2d(buf, &tm, &bar, 0, 0);
        if (MODE == 1) tma_load_2d(buf, tmg, &bar, 0, 0);
        if (MODE == 2) bulk_1d(buf, src, &bar, nel*4);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < nel; i += blockDim.x) out[i] = buf[i];
}
int main()
{
    int n0 = 256, n1 = 64; size_t N = (size_t)n0*n1;
    std::vector<float> h(N); for (size_t i = 0; i < N; ++i) h[i] = (float)i;
    float *d, *o; cudaMalloc(&d, N*4); cudaMalloc(&o, 1<<20); cudaMemcpy(d, h.data(), N*4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(ge), (int)q, p);
    enc_fn enc = (enc_fn)p;
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)n0, (cuuint64_t)n1};
    cuuint64_t str[1] = {(cuuint64_t)n0*4};
    cuuint32_t box[2] = {64, 8};
    cuuint32_t es[2] = {1,1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    CUtensorMap* tmg; cudaMalloc(&tmg, sizeof(CUtensorMap)); cudaMemcpy(tmg, &tm, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    int nel = 64*8;
    for (int mode = 2; mode >= 0; --mode) {
        if (mode == 0) k<0><<<1, 128, nel*4 + 128>>>(tm, tmg, d, o, nel);
        if (mode == 1) k<1><<<1, 128, nel*4 + 128>>>(tm, tmg, d, o, nel);
        if (mode == 2) k<2><<<1, 128, nel*4 + 128>>>(tm, tmg, d, o, nel);
        cudaError_t e = cudaDeviceSynchronize();
        printf("mode %d run -> %s\n", mode, cudaGetErrorString(e));
        if (e != cudaSuccess) break;
        std::vector<float> ho(nel); cudaMemcpy(ho.data(), o, nel*4, cudaMemcpyDeviceToHost);
        printf("  out[1]=%f out[64]=%f\n", ho[1], ho[64]);
    }
    return 0;
}
