// Linear SDE dX = A X dt + G dB with matrix-valued g (nB channels, R/sdetools.R:425-428):
// the drift is `A %*% x` (R/sdetools.R:360-368; vignettes/NoisyOscillator.Rmd:38-54).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

namespace sdegpu {

struct LinearArgs {
    const double *A;          // device, d x d column-major
    const double *G;          // device, d x nB column-major
    const double2 *dtsq;      // [nt-1] {dt, sqrt(dt)}
    const double *icdf;       // device, half-normal quantile table (philox.cuh)
    const double *x0;         // per-path x0[p*d+j] or nullptr
    const double *x0_shared;  // device, d doubles (used when x0 == nullptr)
    const double *B;          // device, B[(p*nB+k)*nt+i] or nullptr (fused generator)
    double *X, *XT;           // X[(p*d+j)*nt+i], XT[p*d+j]
    double *moments;          // [5*d] {n,S1..S4} per component about center, or nullptr
    const double *center;     // device, d doubles
    const double *forcing;    // device, c[j*nt + i]: drift A x + c(t_i) (input term F u(t) of R/sdetools.R:211,:363), or nullptr
    uint64_t npaths, path_offset;
    PhiloxKey seed;
    int nt, d, nB, scheme, strict, accumulate;
};

int launch_linear(const LinearArgs &a, int sm_count, cudaStream_t stream);
// FP64 tensor-core (DMMA) path: euler, fused generator, nB = d, diagonal G, d a multiple of 32 up to 256
bool dmma_eligible(const LinearArgs &a, const double *hostG);
int launch_linear_dmma(const LinearArgs &a, double *scratch, int sm_count, cudaStream_t stream);
// 0: not eligible, 1: diagonal G, 2: dense G (any nB <= d that fits two tiles in shared memory; scratch dp*dp + dp*4*ceil(nB/4))
int dmma_mode(const LinearArgs &a, const double *hostG);
int launch_linear_dmma_dense(const LinearArgs &a, double *scratch, int sm_count, cudaStream_t stream);
// general tensor kernel (linear_tc.cu): euler and heun, supplied B or generator, trajectory, input term; 8 <= d <= 256, FAST
bool linear_tc_eligible(const LinearArgs &a, const double *hostG);
size_t linear_tc_scratch_doubles(const LinearArgs &a);
int launch_linear_tc(const LinearArgs &a, const double *hostG, double *scratch, int sm_count, cudaStream_t stream);

}  // namespace sdegpu
