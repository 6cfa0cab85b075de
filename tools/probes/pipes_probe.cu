// Micro-probe of the SM pipes the RHS kernel leans on (B200, sm_100a): DFMA issue rate, LDS.64 / LDS.128
// wavefront cost, 64-bit shuffle cost, and LDS + SHFL + DFMA issued together.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipes_probe tools/probes/pipes_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) probe(double* out, int stride)
{
    __shared__ double sm[256*6 + 64];
    const int t = threadIdx.x;
    for (int i = t; i < 256*6 + 64; i += 256) sm[i] = i*1e-3;
    __syncthreads();
    double a0 = t*1e-3, a1 = 1.0 + a0, a2 = 2.0 + a0, a3 = 3.0 + a0, a4 = a0 - 1.0, a5 = a0 - 2.0, a6 = a0*0.5, a7 = a0*0.25;
    const double m = 1.0000001, c = 1e-9;
    const double* p = sm + (t*stride) % (256*5);
    #pragma unroll 1
    for (int it = 0; it < ITERS; ++it)
    {
        if (MODE == 0 || MODE == 4 || MODE == 5)         // 8 independent DFMA
        {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
        if (MODE == 1 || MODE == 4 || MODE == 6)         // 4 LDS.64 (stride in doubles chosen by the host)
        {
            const volatile double* q = p;
            a0 += q[0]; a1 += q[1]; a2 += q[2]; a3 += q[3];
        }
        if (MODE == 2)                                   // 2 LDS.128
        {
            const volatile double2* q = (const volatile double2*)(sm + 2*t);
            double2 x = {q[0].x, q[0].y}, y = {q[256].x, q[256].y};
            a0 += x.x; a1 += x.y; a2 += y.x; a3 += y.y;
        }
        if (MODE == 3 || MODE == 5 || MODE == 6)         // 4 64-bit shuffles (8 SHFL.32)
        {
            a0 += __shfl_up_sync(0xffffffffu, a4, 1); a1 += __shfl_up_sync(0xffffffffu, a5, 1);
            a2 += __shfl_up_sync(0xffffffffu, a6, 1); a3 += __shfl_up_sync(0xffffffffu, a7, 1);
        }
        if (MODE == 7 && (t & 31) == 0)                  // 4 LDS.64 issued by ONE lane of each warp (halo hand-off of a boundary lane)
        {
            const volatile double* q = p;
            a0 += q[0]; a1 += q[1]; a2 += q[2]; a3 += q[3];
        }
        if (MODE == 8)                                   // 4 STS.64
        {
            volatile double* q = sm + (t*stride) % (256*5);
            q[0] = a0; q[1] = a1; q[2] = a2; q[3] = a3;
            a0 += 1e-9;
        }
    }
    out[blockIdx.x*256 + t] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int MODE> void run(const char* name, int stride, double per_iter_units, const char* unit)
{
    double* out; cudaMalloc(&out, 148*8*256*8);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int blocks = sms*8;
    probe<MODE><<<blocks, 256>>>(out, stride);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 256>>>(out, stride);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // warp-instructions of interest per SM per cycle (nominal max clock)
    const double warps = blocks*8.0/sms;
    const double cyc = ms*1e-3*clk*1e3;
    printf("%-34s %8.3f ms  %6.3f %s per SM per clk (at %d MHz)\n", name, ms, warps*ITERS*per_iter_units/cyc, unit, clk/1000);
    cudaFree(out);
}

int main()
{
    run<0>("DFMA x8", 1, 8, "warp-DFMA");
    run<1>("LDS.64 x4 stride 1 (SoA)", 1, 4, "warp-LDS.64");
    run<1>("LDS.64 x4 stride 5 (AoS)", 5, 4, "warp-LDS.64");
    run<2>("LDS.128 x2", 1, 2, "warp-LDS.128");
    run<3>("SHFL 64-bit x4", 1, 4, "warp-shfl64");
    run<4>("DFMA x8 + LDS.64 x4", 5, 8, "warp-DFMA");
    run<5>("DFMA x8 + SHFL64 x4", 1, 8, "warp-DFMA");
    run<6>("LDS.64 x4 + SHFL64 x4", 5, 4, "warp-LDS.64 (and as many shfl64)");
    run<7>("LDS.64 x4, one lane per warp", 5, 4, "warp-LDS.64");
    run<8>("STS.64 x4 stride 5", 5, 4, "warp-STS.64");
    return 0;
}
