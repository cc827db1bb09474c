// Host-side launch interface of integrals.cu (shared with api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

namespace sdegpu {

struct IntegralArgs {
    const double *a, *b;         // stochint: f, g       crossvar: X, Y
    double *l, *r, *c;           // stochint outputs (any may be null); crossvar writes `l`
    uint64_t n, ncols;
};
// mode 0: stochint rules l/r/c (itointegral = l), mode 1: cross-variation (QuadraticVariation with a == b)
int launch_integral(int mode, const IntegralArgs &q, int sm_count, cudaStream_t stream);

// Column c: dt <- c(times[1],diff(times)); dB <- u*dt + (sigma*sqrt(dt))*Z; B <- B0+cumsum(dB)
struct BMArgs {
    const double2 *dtsq;     // [nt] {dt_i, sqrt(dt_i)} with dt_0 = times[0]
    const double *icdf;      // device, half-normal quantile table (philox.cuh)
    double *B;               // B[c*nt+i]
    double sigma, B0, u;
    uint64_t ncols, path_offset;
    PhiloxKey seed;
    int nt;
};
int launch_rvbm(const BMArgs &q, int sm_count, cudaStream_t stream);

struct BMFuncArgs {
    BMArgs bm;                   // bm.B unused
    double level;                // first passage level
    double *vmax, *vmin;         // [ncols] max_i B_i, min_i B_i           (any may be null)
    double *tmax, *tmin;         // [ncols] times[i] of the first maximum / minimum
    double *thit;                // [ncols] first times[i] with B_i >= level (level >= B0) or <= level (level < B0); NaN: never
    const double *times;         // device [nt]
};
int launch_bm_functionals(const BMFuncArgs &q, int sm_count, cudaStream_t stream);

}  // namespace sdegpu
