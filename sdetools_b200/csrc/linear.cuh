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
    double *mom_partials;     // [mom_capacity][5*d] per-block sums: every block owns a slice, summed in block order afterwards
    int mom_capacity;         // blocks the slice buffer holds (launchers clamp their grids to it when moments are wanted)
    const double *center;     // device, d doubles
    const double *forcing;    // device, c[j*nt + i]: drift A x + c(t_i) (input term F u(t) of R/sdetools.R:211,:363), or nullptr
    uint64_t npaths, path_offset;
    PhiloxKey seed;
    int nt, d, nB, scheme, strict, accumulate;
};

int launch_linear(const LinearArgs &a, int sm_count, cudaStream_t stream);
// moments <- (accumulate ? moments : 0) + sum over blocks, in block order (deterministic), of the slices the kernel wrote
int reduce_moment_partials(const LinearArgs &a, int nblocks, cudaStream_t stream);

// Per-component power sums of one tile of terminal states held in shared memory as tile[k*ld + n] (component k, path n):
// thread tid owns components tid, tid+nthreads, ... and adds their sums over the tile's np paths, in path order, to its
// block's slice.  No atomics: a component's sums have one writer, so they do not depend on scheduling.
__device__ __forceinline__ void tile_moments(const LinearArgs &a, const double *tile, int ld, int np, int tid, int nthreads, bool first,
                                             int slice_index = -1)
{
    double *slice = a.mom_partials + (size_t)(slice_index < 0 ? (int)blockIdx.x : slice_index) * 5 * a.d;
    for (int k = tid; k < a.d; k += nthreads) {
        double s[5] = {0, 0, 0, 0, 0};
        const double cen = a.center[k];
        for (int n = 0; n < np; ++n) {
            const double dd = tile[k * ld + n] - cen;
            if (isfinite(dd)) {
                const double d2 = dd * dd;
                s[0] += 1.0; s[1] += dd; s[2] += d2; s[3] += d2 * dd; s[4] += d2 * d2;
            }
        }
#pragma unroll
        for (int m = 0; m < 5; ++m) slice[k * 5 + m] = first ? s[m] : slice[k * 5 + m] + s[m];
    }
}
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
