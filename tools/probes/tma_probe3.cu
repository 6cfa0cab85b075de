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
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tm, unsigned char* out, int bytes, int c0, int c1, int c2, int c3)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, bytes);
        if (RANK == 2) tma_load_2d(sm, &tm, &bar, c0, c1);
        if (RANK == 4) tma_load_4d(sm, &tm, &bar, c0, c1, c2, c3);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}
int main()
{
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    enc_fn enc = (enc_fn)p;
    unsigned char *d, *o; cudaMalloc(&d, 64<<20); cudaMalloc(&o, 1<<20); cudaMemset(d, 1, 64<<20);
    struct T { int f64; int rank; int prom; int n0; int bw; int c0; int c1; };
    T tests[] = {{1,4,1,256,64,0,1},{1,4,1,256,64,2,0},{1,4,1,256,64,1,0},{1,4,1,180,64,0,0},{1,4,1,180,170,0,0},{1,4,1,180,170,4,1},{1,4,1,180,170,5,1}};
    for (auto t: tests) {
        int es = t.f64 ? 8 : 4;
        CUtensorMap tm;
        cuuint64_t dims[4] = {(cuuint64_t)t.n0, 64, 8, 2};
        cuuint64_t str[3] = {(cuuint64_t)t.n0*es, (cuuint64_t)t.n0*64*es, (cuuint64_t)t.n0*64*8*es};
        cuuint32_t box[4] = {(cuuint32_t)t.bw, 8, 1, 1};
        cuuint32_t est[4] = {1,1,1,1};
        CUresult r = enc(&tm, t.f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, t.rank, d, dims, str, box, est,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, t.prom ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        int bytes = t.bw*8*es;
        if (t.rank == 2) k<2><<<1, 128, bytes + 128>>>(tm, o, bytes, t.c0, t.c1, 0, 0); else k<4><<<1, 128, bytes + 128>>>(tm, o, bytes, t.c0, t.c1, 3, 1);
        cudaError_t e = cudaDeviceSynchronize();
        printf("n0=%d bw=%d c0=%d c1=%d f64=%d rank=%d prom=%d encode=%d run=%s\n", t.n0, t.bw, t.c0, t.c1, t.f64, t.rank, t.prom, (int)r, cudaGetErrorString(e));
        if (e != cudaSuccess) { printf("(context dead, stopping)\n"); break; }
    }
    return 0;
}
