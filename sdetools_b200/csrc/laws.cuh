// Host-side launch interface of the exact transition-law samplers (laws.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

namespace sdegpu {

struct LawArgs {
    const double *icdf;          // device, half-normal quantile table (philox.cuh)
    const double *x0;            // device, per-draw start x0[p*x0_stride], or nullptr -> x0s
    double x0s;
    double *out;                 // device, out[p] = state after nsteps transitions
    double *path;                // device or nullptr, path[p*(nsteps+1)+i] = state after i transitions
    uint64_t n, path_offset, x0_stride;
    PhiloxKey seed;
    int law, nsteps;             // law: 0 OU, 1 CIR
    double decay;                // exp(-lambda t)
    double xi, sd;               // OU: mean = xi + (x - xi)*decay, sd = sigma*sqrt((1-exp(-2 lambda t))/(2 lambda))
    double two_c, df;            // CIR: 2c = 4 lambda/(gamma^2 (1-decay)), df = 4 lambda xi/gamma^2
};

int launch_rlaw(const LawArgs &q, int sm_count, cudaStream_t stream);
// what = 0 density, 1 distribution function of the transition law at xs[p] given x0[p]; uses law/decay/xi/sd/two_c/df, n, x0
int launch_law_eval(const LawArgs &q, int what, const double *xs, double *out, int sm_count, cudaStream_t stream);

}  // namespace sdegpu
