// probe: which TMA tensor-map shapes work for fp64 on this GPU
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../spade_b200/csrc/spb_tma.cuh"
using namespace spb;
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k4(const __grid_constant__ CUtensorMap tm, double* out, int nel, int c0, int c1, int c2, int c3)
{
    extern __shared__ __align__(128) unsigned char sm[];
    double* buf = (double*)sm;
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_arrive_expect_tx(&bar, nel*8); tma_load_4d(buf, &tm, &bar, c0, c1, c2, c3); }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < nel; i += blockDim.x) out[i] = buf[i];
}
int main(int argc, char** argv)
{
    int bw = argc > 1 ? atoi(argv[1]) : 170;   // inner box (elements)
    int bh = argc > 2 ? atoi(argv[2]) : 10;
    int n0 = 180, n1 = 36, n2 = 36, n3 = 2;
    size_t N = (size_t)n0*n1*n2*n3;
    std::vector<double> h(N); for (size_t i = 0; i < N; ++i) h[i] = (double)i;
    double *d, *o; cudaMalloc(&d, N*8); cudaMalloc(&o, 1<<20); cudaMemcpy(d, h.data(), N*8, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    enc_fn enc = (enc_fn)p;
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)n0, (cuuint64_t)n1, (cuuint64_t)n2, (cuuint64_t)n3};
    cuuint64_t str[3] = {(cuuint64_t)n0*8, (cuuint64_t)n0*n1*8, (cuuint64_t)n0*n1*n2*8};
    cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
    cuuint32_t es[4] = {1,1,1,1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode bw=%d bh=%d -> %d\n", bw, bh, (int)r);
    int nel = bw*bh;
    cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    k4<<<1, 128, nel*8 + 128>>>(tm, o, nel, 5, 1, 3, 1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run -> %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<double> ho(nel); cudaMemcpy(ho.data(), o, nel*8, cudaMemcpyDeviceToHost);
        double exp0 = 5 + 180.0*(1 + 36.0*(3 + 36.0*1));
        printf("first %f expect %f ; elem[bw] %f expect %f\n", ho[0], exp0, ho[bw], exp0 + 180);
    }
    return 0;
}
