// Microbenchmark: which part of the k_terminal loop keeps the issue rate at ~0.65 IPC?
// Runs the CIR/heun loop with components removed and reports cycles per warp-step per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o terminal_parts tools/microbench/terminal_parts.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#define ICDF_TABLE_DEFINE __device__ const unsigned long long g_icdf_table
#include "../../sdetools_b200/csrc/models.cuh"
#include "../../sdetools_b200/csrc/philox.cuh"

using namespace sdegpu;
using M = Model<SDEGPU_MODEL_CIR>;

// MODE bit 0: Philox, bit 1: ICDF, bit 2: folded heun step, bit 3: generic heun step
template <int MODE, int THREADS, int NP>
__global__ void __launch_bounds__(THREADS, 1) k(const double *icdf, const double *coef, PhiloxKey key, int nsteps, double *sink)
{
    extern __shared__ __align__(16) double sh_tab[];
    icdf_stage_hot(sh_tab, icdf, threadIdx.x, THREADS);
    __syncthreads();
    KParams P; P.p[0] = 1; P.p[1] = 1; P.p[2] = 1; P.p[3] = 0;
    double x[NP][1];
    uint32_t r[NP][4];
    const uint64_t gp = (uint64_t)blockIdx.x * THREADS + threadIdx.x;
    for (int q = 0; q < NP; ++q) { x[q][0] = 1.0; r[q][0] = (uint32_t)gp * 2654435761u + q; r[q][1] = r[q][0] ^ 0x9E3779B9u; r[q][2] = r[q][1] * 7u; r[q][3] = r[q][2] + 13u; }
    for (int i = 0; i + 1 < nsteps; i += 2) {
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            uint32_t n0, n1, n2, n3;
            if (MODE & 1) philox_block(key, gp + q * 1000003u, (uint32_t)(i >> 1) + 1u, 0u, n0, n1, n2, n3);
            else { n0 = r[q][1] + 0x1234567u; n1 = r[q][2] ^ (r[q][0] >> 3); n2 = r[q][3] + r[q][0]; n3 = r[q][0] * 5u + 1u; }
            double z0, z1;
            if (MODE & 2) { z0 = icdf_normal<true>(r[q][0], r[q][1], sh_tab); z1 = icdf_normal<true>(r[q][2], r[q][3], sh_tab); }
            else { z0 = __hiloint2double(0x3FE00000 | (r[q][1] >> 12), r[q][0]) - 0.75; z1 = __hiloint2double(0x3FE00000 | (r[q][3] >> 12), r[q][2]) - 0.75; }
            if (MODE & 4) {
                FastStep<M, SDEGPU_HEUN, SDEGPU_PROJ_ABS>::apply(P, x[q], make_double2(0, 0), coef + (size_t)i * FAST_NCOEF, z0);
                FastStep<M, SDEGPU_HEUN, SDEGPU_PROJ_ABS>::apply(P, x[q], make_double2(0, 0), coef + (size_t)(i + 1) * FAST_NCOEF, z1);
            } else if (MODE & 8) {
                sde_step<M, SDEGPU_HEUN, SDEGPU_PROJ_ABS, Fast>(P, x[q], 1e-4, 1e-2 * z0);
                sde_step<M, SDEGPU_HEUN, SDEGPU_PROJ_ABS, Fast>(P, x[q], 1e-4, 1e-2 * z1);
            } else x[q][0] += z0 * z1;
            r[q][0] = n0; r[q][1] = n1; r[q][2] = n2; r[q][3] = n3;
        }
    }
    double s = 0;
    for (int q = 0; q < NP; ++q) s += x[q][0];
    if (s == 1.2345) sink[0] = s;
}

template <int MODE, int THREADS, int NP>
void run(const char *name, const double *icdf, const double *coef, int sms, int khz)
{
    const int nsteps = 4000;
    double *sink; cudaMalloc(&sink, 8);
    auto kern = k<MODE, THREADS, NP>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ICDF_HOT_TABLE_BYTES);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        kern<<<sms, THREADS, ICDF_HOT_TABLE_BYTES>>>(icdf, coef, make_philox_key(42), nsteps, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double warp_steps_per_smsp = (double)nsteps * (THREADS / 32) * NP / 4.0;
    printf("%-28s threads %4d np %d: %8.3f ms  %6.1f cycles per warp-step per SMSP  (%s)\n", name, THREADS, NP, best,
           best * 1e-3 * khz * 1e3 / warp_steps_per_smsp, cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double *icdf; cudaGetSymbolAddress((void **)&icdf, g_icdf_table);
    std::vector<double> hc(4000 * FAST_NCOEF);
    double par[3] = {1, 1, 1};
    for (int i = 0; i < 4000; ++i) fast_step_coefficients(SDEGPU_MODEL_CIR, par, 1e-4, 1e-2, &hc[i * FAST_NCOEF]);
    double *coef; cudaMalloc(&coef, hc.size() * 8); cudaMemcpy(coef, hc.data(), hc.size() * 8, cudaMemcpyHostToDevice);
    const int sms = p.multiProcessorCount;
    printf("%s %d SMs %d kHz\n", p.name, sms, khz);
    run<1, 512, 2>("philox only", icdf, coef, sms, khz);
    run<2, 512, 2>("icdf only", icdf, coef, sms, khz);
    run<3, 512, 2>("philox+icdf", icdf, coef, sms, khz);
    run<4, 512, 2>("folded step only", icdf, coef, sms, khz);
    run<8, 512, 2>("generic step only", icdf, coef, sms, khz);
    run<6, 512, 2>("icdf+folded", icdf, coef, sms, khz);
    run<5, 512, 2>("philox+folded", icdf, coef, sms, khz);
    run<7, 512, 2>("all (folded)", icdf, coef, sms, khz);
    run<11, 512, 2>("all (generic)", icdf, coef, sms, khz);
    run<7, 512, 1>("all (folded) np1", icdf, coef, sms, khz);
    run<7, 1024, 1>("all (folded) 1024 np1", icdf, coef, sms, khz);
    run<7, 256, 2>("all (folded) 256 np2", icdf, coef, sms, khz);
    run<7, 256, 4>("all (folded) 256 np4", icdf, coef, sms, khz);
    return 0;
}
