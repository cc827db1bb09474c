// Dependent-issue latency of the instructions on the per-step critical path (one warp per SM, one chain).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int KIND>
__global__ void k(int iters, double *sink, long long *cyc)
{
    double d = 1.5 + threadIdx.x * 1e-6; uint32_t u = threadIdx.x * 2654435761u + 1u;
    __shared__ double tab[64];
    tab[threadIdx.x & 63] = 1.0 + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (KIND == 0) d = fma(d, 1.0000001, 1e-9);
            else if (KIND == 1) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(d));
            else if (KIND == 2) { uint64_t p = (uint64_t)u * 0xD2511F53u; u = (uint32_t)(p >> 32) ^ (uint32_t)p; }
            else if (KIND == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u) : "r"(u >> 1), "r"(0x9E3779B9u));
            else if (KIND == 4) d = tab[(__double2loint(d) >> 3) & 63] + 1e-9;          // LDS + DADD
            else if (KIND == 5) d = d + 1.0000001;
            else if (KIND == 6) d = d * 1.0000001;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    if (d + u == 1.2345) sink[0] = d;
}
template <int KIND> void run(const char *name, int per = 8)
{
    double *sink; long long *cyc; cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<KIND><<<1, 32>>>(iters, sink, cyc); k<KIND><<<1, 32>>>(iters, sink, cyc);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-40s %7.2f cycles per dependent op\n", name, (double)h / (iters * per));
}
int main()
{
    run<0>("DFMA -> DFMA"); run<5>("DADD -> DADD"); run<6>("DMUL -> DMUL"); run<1>("MUFU.RSQ64H -> MUFU.RSQ64H");
    run<2>("IMAD.WIDE -> LOP3 -> (pair)"); run<3>("SHF+LOP3 chain (pair)"); run<4>("LDS -> DADD -> (pair)");
    return 0;
}
