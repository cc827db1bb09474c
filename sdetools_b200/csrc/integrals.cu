// Stochastic integrals on stored paths, batched over columns:
//   stochint(f,g,rule)      R/sdetools.R:105-113   l = c(0,cumsum(f[-n]*dg)), r = c(0,cumsum(f[-1]*dg)), c = 0.5*(l+r)
//   itointegral(G,B)        R/sdetools.R:155-158   == stochint left rule
//   QuadraticVariation(X)   R/sdetools.R:134-137   c(0,cumsum(diff(X)^2))
//   CrossVariation(X,Y)     R/sdetools.R:184-187   c(0,cumsum(diff(X)*diff(Y)))
// plus rvBM (R/sdetools.R:16-22,40-46) whose B0+cumsum(dB) uses the same accumulator.
//
// A prefix sum with R's rounding is a strict recurrence per column, so the
// parallel axis is the column (= path): one thread per column, 80-bit
// accumulator in registers (float80.cuh), columns staged through shared-memory
// tiles so that global traffic stays coalesced along time.
#include <cuda_runtime.h>
#include <stdint.h>

#include "float80.cuh"
#include "philox.cuh"

namespace sdegpu {

constexpr int IT_COLS = 128;     // columns (= threads) per block
constexpr int IT_STEPS = 32;     // increments staged per tile
constexpr int IT_LD = IT_STEPS + 1;

struct IntegralArgs {
    const double *a, *b;         // stochint: f, g       crossvar: X, Y
    double *l, *r, *c;           // stochint outputs (any may be null); crossvar writes `l`
    uint64_t n, ncols;
};

template <int MODE>              // 0 stochint, 1 cross-variation
__global__ void __launch_bounds__(IT_COLS) k_integral(const IntegralArgs q)
{
    extern __shared__ __align__(16) double sm[];
    double *sA = sm, *sB = sm + IT_COLS * IT_LD, *sC = sB + IT_COLS * IT_LD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t ntiles = (q.ncols + IT_COLS - 1) / IT_COLS;
    const uint64_t nsteps = q.n - 1;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t c0 = tile * IT_COLS, col = c0 + tid;
        const bool valid = col < q.ncols;
        const int rows = (int)((q.ncols - c0 < (uint64_t)IT_COLS) ? (q.ncols - c0) : IT_COLS);
        F80 accl = f80_zero(), accr = f80_zero();
        if (valid) {                                            // c(0, ...)
            if (q.l) q.l[col * q.n] = 0.0;
            if (MODE == 0 && q.r) q.r[col * q.n] = 0.0;
            if (MODE == 0 && q.c) q.c[col * q.n] = 0.0;
        }
        for (uint64_t i0 = 0; i0 < nsteps; i0 += IT_STEPS) {
            const int nc = (int)((nsteps - i0 < (uint64_t)IT_STEPS) ? (nsteps - i0) : IT_STEPS);
            for (int rr = warp; rr < rows; rr += IT_COLS / 32) {
                const double *pa = q.a + (c0 + rr) * q.n + i0, *pb = q.b + (c0 + rr) * q.n + i0;
                for (int k = lane; k <= nc; k += 32) {
                    sA[rr * IT_LD + k] = __ldg(pa + k);
                    sB[rr * IT_LD + k] = __ldg(pb + k);
                }
            }
            __syncthreads();
            if (valid) {
                double a0 = sA[tid * IT_LD], b0 = sB[tid * IT_LD];
                for (int k = 0; k < nc; ++k) {
                    const double a1 = sA[tid * IT_LD + k + 1], b1 = sB[tid * IT_LD + k + 1];
                    if (MODE == 0) {
                        const double dg = __dsub_rn(b1, b0);
                        f80_add(accl, __dmul_rn(a0, dg));
                        f80_add(accr, __dmul_rn(a1, dg));
                        const double lv = f80_to_double(accl), rv = f80_to_double(accr);
                        sA[tid * IT_LD + k] = lv;
                        sB[tid * IT_LD + k] = rv;
                        sC[tid * IT_LD + k] = __dmul_rn(0.5, __dadd_rn(lv, rv));
                    } else {
                        f80_add(accl, __dmul_rn(__dsub_rn(a1, a0), __dsub_rn(b1, b0)));
                        sA[tid * IT_LD + k] = f80_to_double(accl);
                    }
                    a0 = a1; b0 = b1;
                }
            }
            __syncthreads();
            for (int rr = warp; rr < rows; rr += IT_COLS / 32) {
                const uint64_t off = (c0 + rr) * q.n + i0 + 1;
                for (int k = lane; k < nc; k += 32) {
                    if (q.l) q.l[off + k] = sA[rr * IT_LD + k];
                    if (MODE == 0 && q.r) q.r[off + k] = sB[rr * IT_LD + k];
                    if (MODE == 0 && q.c) q.c[off + k] = sC[rr * IT_LD + k];
                }
            }
            __syncthreads();
        }
    }
}

int launch_integral(int mode, const IntegralArgs &q, int sm_count, cudaStream_t stream)
{
    if (q.n == 0 || q.ncols == 0) return 0;
    const size_t smem = sizeof(double) * IT_COLS * IT_LD * 3;
    const uint64_t ntiles = (q.ncols + IT_COLS - 1) / IT_COLS;
    auto k0 = k_integral<0>;
    auto k1 = k_integral<1>;
    cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mode == 0 ? k0 : k1, IT_COLS, smem);
    uint64_t grid = (uint64_t)(per_sm < 1 ? 1 : per_sm) * sm_count;
    if (grid > ntiles) grid = ntiles;
    if (mode == 0) k0<<<(int)grid, IT_COLS, smem, stream>>>(q);
    else k1<<<(int)grid, IT_COLS, smem, stream>>>(q);
    return (int)cudaGetLastError();
}

// ---- rvBM ensembles ---------------------------------------------------------
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

__global__ void __launch_bounds__(IT_COLS) k_rvbm(const BMArgs q)
{
    extern __shared__ __align__(16) double sh_tab[];             // [quantile table][output tile]
    double *sm = sh_tab + ICDF_TABLE_DOUBLES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    icdf_stage(sh_tab, q.icdf, tid, IT_COLS);
    __syncthreads();
    const uint64_t ntiles = (q.ncols + IT_COLS - 1) / IT_COLS;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t c0 = tile * IT_COLS, col = c0 + tid;
        const bool valid = col < q.ncols;
        const int rows = (int)((q.ncols - c0 < (uint64_t)IT_COLS) ? (q.ncols - c0) : IT_COLS);
        F80 acc = f80_zero();
        double z0 = 0.0, z1 = 0.0;
        for (int i0 = 0; i0 < q.nt; i0 += IT_STEPS) {
            const int nc = (q.nt - i0 < IT_STEPS) ? (q.nt - i0) : IT_STEPS;
            if (valid) {
                for (int k = 0; k < nc; ++k) {
                    const int i = i0 + k;
                    if ((i & 1) == 0) normal_pair(q.seed, q.path_offset + col, (uint32_t)(i >> 1), 0u, sh_tab, z0, z1);
                    const double2 d = __ldg(&q.dtsq[i]);
                    // rnorm(mean=u*dt, sd=sigma*sqrt(dt)) = mean + sd*Z
                    const double dB = __dadd_rn(__dmul_rn(q.u, d.x), __dmul_rn(__dmul_rn(q.sigma, d.y), (i & 1) ? z1 : z0));
                    f80_add(acc, dB);
                    sm[tid * IT_LD + k] = __dadd_rn(q.B0, f80_to_double(acc));
                }
            }
            __syncthreads();
            for (int rr = warp; rr < rows; rr += IT_COLS / 32)
                for (int k = lane; k < nc; k += 32) q.B[(c0 + rr) * (uint64_t)q.nt + i0 + k] = sm[rr * IT_LD + k];
            __syncthreads();
        }
    }
}

int launch_rvbm(const BMArgs &q, int sm_count, cudaStream_t stream)
{
    if (q.ncols == 0 || q.nt == 0) return 0;
    const size_t smem = ICDF_TABLE_BYTES + sizeof(double) * IT_COLS * IT_LD;
    const uint64_t ntiles = (q.ncols + IT_COLS - 1) / IT_COLS;
    int per_sm = 1;
    cudaFuncSetAttribute(k_rvbm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rvbm, IT_COLS, smem);
    uint64_t grid = (uint64_t)(per_sm < 1 ? 1 : per_sm) * sm_count;
    if (grid > ntiles) grid = ntiles;
    k_rvbm<<<(int)grid, IT_COLS, smem, stream>>>(q);
    return (int)cudaGetLastError();
}

}  // namespace sdegpu
