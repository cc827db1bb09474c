// Microbenchmark: does an FP64 instruction (1 warp-instr per 2 cycles per SM sub-partition on B200)
// block the issue port for the second cycle, or can integer/FP32 work issue in its shadow?
// Runs NI integer ops per DFMA (independent chains) and reports cycles per DFMA per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu && ./fp64_issue
#include <cstdio>
#include <cuda_runtime.h>

template <int NI, int KIND>   // KIND 0: IMAD  1: LOP3/IADD (alu)  2: FFMA
__global__ void __launch_bounds__(256) k(int iters, double *sink, unsigned *isink)
{
    double d[4];
    unsigned u[8];
    float f[8];
    for (int i = 0; i < 4; ++i) d[i] = 1.0 + threadIdx.x * 1e-9 + i;
    for (int i = 0; i < 8; ++i) { u[i] = threadIdx.x * 2654435761u + i; f[i] = 1.0f + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            d[c] = fma(d[c], 1.0000000001, 1e-12);
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                const int q = (c * NI + j) & 7;
                if (KIND == 0) u[q] = u[q] * 0x9E3779B9u + 0x7F4A7C15u;
                else if (KIND == 1) u[q] = (u[q] ^ (u[(q + 1) & 7] + 0x85EBCA6Bu));
                else f[q] = fmaf(f[q], 1.0000001f, 1e-7f);
            }
        }
    }
    double s = 0; unsigned t = 0; float g = 0;
    for (int i = 0; i < 4; ++i) s += d[i];
    for (int i = 0; i < 8; ++i) { t ^= u[i]; g += f[i]; }
    if (s == 123.456) sink[0] = s + g;
    if (t == 0x12345678u) isink[0] = t;
}

template <int NI, int KIND>
void run(const char *name, int sms, int khz)
{
    double *sink; unsigned *isink;
    cudaMalloc(&sink, 8); cudaMalloc(&isink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 15, blocks = sms * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<NI, KIND><<<blocks, 256>>>(iters, sink, isink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    // warp-instructions of DFMA per sub-partition
    const double dfma_per_smsp = (double)iters * 4 * (256 / 32) * 8 / 4;   // blocks/SM=8, warps/block=8, 4 SMSPs
    const double cycles = best * 1e-3 * khz * 1e3;
    printf("%-6s NI=%d  %.3f ms  cycles/DFMA/SMSP = %.2f   (issue slots per DFMA group = %d)\n", name, NI, best, cycles / dfma_per_smsp, 1 + NI);
    cudaFree(sink); cudaFree(isink);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, %d kHz\n", p.name, p.multiProcessorCount, khz);
    run<0, 0>("none", p.multiProcessorCount, khz);
    run<1, 0>("imad", p.multiProcessorCount, khz); run<2, 0>("imad", p.multiProcessorCount, khz); run<3, 0>("imad", p.multiProcessorCount, khz);
    run<1, 1>("alu", p.multiProcessorCount, khz); run<2, 1>("alu", p.multiProcessorCount, khz); run<3, 1>("alu", p.multiProcessorCount, khz);
    run<1, 2>("ffma", p.multiProcessorCount, khz); run<2, 2>("ffma", p.multiProcessorCount, khz); run<3, 2>("ffma", p.multiProcessorCount, khz);
    return 0;
}
