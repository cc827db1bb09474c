// Linear SDE kernels.
//   k_linear_generic : any d, any nB, STRICT or FAST, supplied B or fused generator.
//       One thread per (path, state row); x lives in shared memory, A streams through L1/L2.
//       `A %*% x` is accumulated in column order k = 0..d-1 with separate multiply and add
//       (reference dgemv order) when strict.
#include "linear.cuh"

#include "../../include/sdegpu.h"
#include "philox.cuh"

namespace sdegpu {

constexpr int LIN_THREADS = 256;

template <bool STRICT>
__device__ __forceinline__ double madd(double acc, double a, double b)
{
    return STRICT ? __dadd_rn(acc, __dmul_rn(a, b)) : fma(a, b, acc);
}

// dynamic smem: xs[PT*d] ys[PT*d] ns[PT*d] dBs[PT*nB] zs[PT*nB] [quantile table if !B]
template <bool STRICT>
__global__ void __launch_bounds__(LIN_THREADS) k_linear_generic(const LinearArgs a, const int PT)
{
    extern __shared__ __align__(16) double sm[];
    const int d = a.d, nB = a.nB, nsteps = a.nt - 1;
    double *xs = sm, *ys = xs + (size_t)PT * d, *ns = ys + (size_t)PT * d;
    double *dBs = ns + (size_t)PT * d, *zs = dBs + (size_t)PT * nB;
    double *sh_tab = sm + (((size_t)3 * PT * d + (size_t)2 * PT * nB + 1) & ~(size_t)1);    // 16-byte aligned from the base
    const int tid = threadIdx.x;
    if (!a.B) icdf_stage(sh_tab, a.icdf, tid, LIN_THREADS);
    __syncthreads();
    const bool heun = a.scheme == SDEGPU_HEUN;
    const uint64_t ngroups = (a.npaths + PT - 1) / PT;
    for (uint64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const uint64_t p0 = grp * PT;
        const int np = (int)((a.npaths - p0 < (uint64_t)PT) ? (a.npaths - p0) : PT);
        for (int e = tid; e < np * d; e += LIN_THREADS) {
            const int q = e / d, j = e - q * d;
            const double v = a.x0 ? a.x0[(p0 + q) * d + j] : a.x0_shared[j];
            xs[e] = v;
            if (a.X) a.X[((p0 + q) * d + j) * (uint64_t)a.nt] = v;          // X[1,] <- x0 (:441)
        }
        __syncthreads();
        for (int i = 0; i < nsteps; ++i) {
            const double2 ds = __ldg(&a.dtsq[i]);
            // increments of this step for every (path, channel)
            for (int e = tid; e < np * nB; e += LIN_THREADS) {
                const int q = e / nB, k = e - q * nB;
                if (a.B) {
                    const double *b = a.B + ((p0 + q) * nB + k) * (uint64_t)a.nt + i;
                    dBs[e] = __dsub_rn(b[1], b[0]);                           // dB <- apply(B,2,diff) (:437)
                } else if (nB == 1) {
                    // one channel: the scalar models' stream (four normals per block), so a linear model with one
                    // noise channel sees the same Brownian motion as a catalogue model under the same seed
                    double zq[4];
                    normal_quad(a.seed, a.path_offset + p0 + q, (uint32_t)(i >> 2), 0u, sh_tab, zq);
                    const int s4 = i & 3;
                    dBs[e] = ds.y * (s4 == 0 ? zq[0] : (s4 == 1 ? zq[1] : (s4 == 2 ? zq[2] : zq[3])));
                } else {
                    // matrix-valued g: one Philox block per (path, step, channel pair), shared with the DMMA kernel
                    double z0, z1;
                    normal_pair(a.seed, a.path_offset + p0 + q, (uint32_t)i, 0x80000000u + (uint32_t)(k >> 1), sh_tab, z0, z1);
                    dBs[e] = ds.y * ((k & 1) ? z1 : z0);
                }
            }
            __syncthreads();
            // Y = (x + (A x) dt) + G dB        (euler :453 / heun predictor :318)
            for (int e = tid; e < np * d; e += LIN_THREADS) {
                const int q = e / d, j = e - q * d;
                const double *x = xs + q * d, *db = dBs + q * nB;
                double f = STRICT ? __dmul_rn(a.A[j], x[0]) : a.A[j] * x[0];
                for (int k = 1; k < d; ++k) f = madd<STRICT>(f, a.A[(size_t)k * d + j], x[k]);
                if (a.forcing) f = __dadd_rn(f, __ldg(a.forcing + (size_t)j * a.nt + i));            // A %*% x + c(t_i)
                double gdb = STRICT ? __dmul_rn(a.G[j], db[0]) : a.G[j] * db[0];
                for (int k = 1; k < nB; ++k) gdb = madd<STRICT>(gdb, a.G[(size_t)k * d + j], db[k]);
                ys[e] = STRICT ? __dadd_rn(__dadd_rn(x[j], __dmul_rn(f, ds.x)), gdb) : (x[j] + f * ds.x) + gdb;
            }
            __syncthreads();
            const double *fresh = ys;
            if (heun) {
                // corrector (:323): x' = (x + (0.5*(fX+fY))*dt) + 0.5*((gX+gY) %*% dB); g is constant, gX+gY = G+G
                for (int e = tid; e < np * d; e += LIN_THREADS) {
                    const int q = e / d, j = e - q * d;
                    const double *x = xs + q * d, *y = ys + q * d, *db = dBs + q * nB;
                    double fx = STRICT ? __dmul_rn(a.A[j], x[0]) : a.A[j] * x[0];
                    double fy = STRICT ? __dmul_rn(a.A[j], y[0]) : a.A[j] * y[0];
                    for (int k = 1; k < d; ++k) {
                        const double ajk = a.A[(size_t)k * d + j];
                        fx = madd<STRICT>(fx, ajk, x[k]);
                        fy = madd<STRICT>(fy, ajk, y[k]);
                    }
                    if (a.forcing) {                                             // ff(times[i],X), ff(times[i+1],Y)  (:316,:321)
                        fx = __dadd_rn(fx, __ldg(a.forcing + (size_t)j * a.nt + i));
                        fy = __dadd_rn(fy, __ldg(a.forcing + (size_t)j * a.nt + i + 1));
                    }
                    const double g0 = a.G[j];
                    double gdb = STRICT ? __dmul_rn(__dadd_rn(g0, g0), db[0]) : (g0 + g0) * db[0];
                    for (int k = 1; k < nB; ++k) {
                        const double gk = a.G[(size_t)k * d + j];
                        gdb = madd<STRICT>(gdb, STRICT ? __dadd_rn(gk, gk) : gk + gk, db[k]);
                    }
                    ns[e] = STRICT
                        ? __dadd_rn(__dadd_rn(x[j], __dmul_rn(__dmul_rn(0.5, __dadd_rn(fx, fy)), ds.x)), __dmul_rn(0.5, gdb))
                        : (x[j] + (0.5 * (fx + fy)) * ds.x) + 0.5 * gdb;
                }
                __syncthreads();
                fresh = ns;
            }
            for (int e = tid; e < np * d; e += LIN_THREADS) {
                const double v = fresh[e];
                xs[e] = v;
                if (a.X) {
                    const int q = e / d, j = e - q * d;
                    a.X[((p0 + q) * d + j) * (uint64_t)a.nt + i + 1] = v;
                }
            }
            __syncthreads();
        }
        for (int e = tid; e < np * d; e += LIN_THREADS) {
            const int q = e / d, j = e - q * d;
            if (a.XT) a.XT[(p0 + q) * d + j] = xs[e];
        }
        if (a.moments) {                                         // thread j owns component j: sums over the tile's paths in path order
            double *slice = a.mom_partials + (size_t)blockIdx.x * 5 * d;
            for (int j = tid; j < d; j += LIN_THREADS) {
                double s[5] = {0, 0, 0, 0, 0};
                for (int q = 0; q < np; ++q) {
                    const double dd = xs[q * d + j] - a.center[j];
                    if (isfinite(dd)) {
                        const double d2 = dd * dd;
                        s[0] += 1.0; s[1] += dd; s[2] += d2; s[3] += d2 * dd; s[4] += d2 * d2;
                    }
                }
                for (int m = 0; m < 5; ++m) slice[j * 5 + m] = grp == blockIdx.x ? s[m] : slice[j * 5 + m] + s[m];
            }
        }
        __syncthreads();
    }
}

// ---- small systems (d, nB <= 4: the noisy oscillator of vignettes/NoisyOscillator.Rmd and its kin) ------------
// One thread per path with x, A and G in registers — no shared-memory state, no barriers.  Same operations in
// the same order as k_linear_generic (bit-identical when STRICT) and the same generator stream.
constexpr int LIN_SMALL_MAX = 4;

template <int D, int NB, bool STRICT>
__global__ void __launch_bounds__(LIN_THREADS) k_linear_small(const LinearArgs a)
{
    extern __shared__ __align__(16) double sh_tab[];
    if (!a.B) icdf_stage(sh_tab, a.icdf, threadIdx.x, LIN_THREADS);
    __syncthreads();
    double Ar[D][D], Gr[D][NB];                                  // Ar[j][k] = A[j,k]
#pragma unroll
    for (int j = 0; j < D; ++j) {
#pragma unroll
        for (int k = 0; k < D; ++k) Ar[j][k] = __ldg(&a.A[k * D + j]);
#pragma unroll
        for (int k = 0; k < NB; ++k) Gr[j][k] = __ldg(&a.G[k * D + j]);
    }
    const bool heun = a.scheme == SDEGPU_HEUN;
    const int nsteps = a.nt - 1;
    const uint64_t stride = (uint64_t)gridDim.x * LIN_THREADS;
    const uint64_t rounds = (a.npaths + stride - 1) / stride;
    double mom[D][5];
#pragma unroll
    for (int j = 0; j < D; ++j)
#pragma unroll
        for (int m = 0; m < 5; ++m) mom[j][m] = 0.0;
    for (uint64_t r = 0; r < rounds; ++r) {
        const uint64_t p = r * stride + (uint64_t)blockIdx.x * LIN_THREADS + threadIdx.x;
        if (p >= a.npaths) continue;
        double x[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            x[j] = a.x0 ? a.x0[p * D + j] : a.x0_shared[j];
            if (a.X) a.X[(p * D + j) * (uint64_t)a.nt] = x[j];                 // X[1,] <- x0 (:441)
        }
        double zq[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < nsteps; ++i) {
            const double2 ds = __ldg(&a.dtsq[i]);
            double dB[NB];
            if (a.B) {
#pragma unroll
                for (int k = 0; k < NB; ++k) {
                    const double *b = a.B + (p * NB + k) * (uint64_t)a.nt + i;
                    dB[k] = __dsub_rn(__ldg(b + 1), __ldg(b));                  // dB <- apply(B,2,diff) (:437)
                }
            } else if (NB == 1) {                                               // the scalar models' stream: 4 normals per block
                if ((i & 3) == 0) normal_quad(a.seed, a.path_offset + p, (uint32_t)(i >> 2), 0u, sh_tab, zq);
                const int s4 = i & 3;
                dB[0] = __dmul_rn(ds.y, s4 == 0 ? zq[0] : (s4 == 1 ? zq[1] : (s4 == 2 ? zq[2] : zq[3])));
            } else {
#pragma unroll
                for (int k = 0; k < NB; k += 2) {                               // one Philox block per channel pair
                    double z0, z1;
                    normal_pair(a.seed, a.path_offset + p, (uint32_t)i, 0x80000000u + (uint32_t)(k >> 1), sh_tab, z0, z1);
                    dB[k] = __dmul_rn(ds.y, z0);
                    if (k + 1 < NB) dB[k + 1] = __dmul_rn(ds.y, z1);
                }
            }
            double y[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {                                       // (x + (A x) dt) + G dB  (:453 / :318)
                double f = STRICT ? __dmul_rn(Ar[j][0], x[0]) : Ar[j][0] * x[0];
#pragma unroll
                for (int k = 1; k < D; ++k) f = madd<STRICT>(f, Ar[j][k], x[k]);
                if (a.forcing) f = __dadd_rn(f, __ldg(a.forcing + (size_t)j * a.nt + i));            // A %*% x + c(t_i)
                double gdb = STRICT ? __dmul_rn(Gr[j][0], dB[0]) : Gr[j][0] * dB[0];
#pragma unroll
                for (int k = 1; k < NB; ++k) gdb = madd<STRICT>(gdb, Gr[j][k], dB[k]);
                y[j] = STRICT ? __dadd_rn(__dadd_rn(x[j], __dmul_rn(f, ds.x)), gdb) : (x[j] + f * ds.x) + gdb;
            }
            if (heun) {                                                         // corrector (:323), gX + gY = G + G
                double n[D];
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    double fx = STRICT ? __dmul_rn(Ar[j][0], x[0]) : Ar[j][0] * x[0];
                    double fy = STRICT ? __dmul_rn(Ar[j][0], y[0]) : Ar[j][0] * y[0];
#pragma unroll
                    for (int k = 1; k < D; ++k) {
                        fx = madd<STRICT>(fx, Ar[j][k], x[k]);
                        fy = madd<STRICT>(fy, Ar[j][k], y[k]);
                    }
                    if (a.forcing) {                                             // ff(times[i],X), ff(times[i+1],Y)  (:316,:321)
                        fx = __dadd_rn(fx, __ldg(a.forcing + (size_t)j * a.nt + i));
                        fy = __dadd_rn(fy, __ldg(a.forcing + (size_t)j * a.nt + i + 1));
                    }
                    const double g0 = Gr[j][0];
                    double gdb = STRICT ? __dmul_rn(__dadd_rn(g0, g0), dB[0]) : (g0 + g0) * dB[0];
#pragma unroll
                    for (int k = 1; k < NB; ++k) {
                        const double gk = Gr[j][k];
                        gdb = madd<STRICT>(gdb, STRICT ? __dadd_rn(gk, gk) : gk + gk, dB[k]);
                    }
                    n[j] = STRICT
                        ? __dadd_rn(__dadd_rn(x[j], __dmul_rn(__dmul_rn(0.5, __dadd_rn(fx, fy)), ds.x)), __dmul_rn(0.5, gdb))
                        : (x[j] + (0.5 * (fx + fy)) * ds.x) + 0.5 * gdb;
                }
#pragma unroll
                for (int j = 0; j < D; ++j) y[j] = n[j];
            }
#pragma unroll
            for (int j = 0; j < D; ++j) {
                x[j] = y[j];
                if (a.X) a.X[(p * D + j) * (uint64_t)a.nt + i + 1] = x[j];
            }
        }
#pragma unroll
        for (int j = 0; j < D; ++j) {
            if (a.XT) a.XT[p * D + j] = x[j];
            if (a.moments) {
                const double dd = x[j] - a.center[j];
                if (isfinite(dd)) {
                    const double d2 = dd * dd;
                    mom[j][0] += 1.0; mom[j][1] += dd; mom[j][2] += d2; mom[j][3] += d2 * dd; mom[j][4] += d2 * d2;
                }
            }
        }
    }
    if (a.moments) {                                             // per-thread sums -> warp (shuffle tree) -> warps in index order: no atomics
        __shared__ double sh_mom[LIN_THREADS / 32][D * 5];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                double v = mom[j][m];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if ((threadIdx.x & 31) == 0) sh_mom[threadIdx.x >> 5][j * 5 + m] = v;
            }
        __syncthreads();
        if (threadIdx.x < D * 5) {
            double v = 0.0;
            for (int w = 0; w < LIN_THREADS / 32; ++w) v += sh_mom[w][threadIdx.x];
            a.mom_partials[(size_t)blockIdx.x * 5 * D + threadIdx.x] = v;
        }
    }
}

// ---- medium systems (4 < d <= 32, nB <= 32): one lane per (path, row), a warp holds 32/DP paths ------------------
// DP = d rounded up to 8, 16 or 32.  Lane (q, j) keeps x_j, row j of A and row j of G in registers; A x and G dB are
// accumulated over k in the reference's order with x_k and dB_k fetched from lane (q, k) by shuffle — no shared
// memory state, no block barriers (k_linear_generic needs three or four per step).  Same operations, same order,
// same generator stream as the other linear kernels.
template <int DP, bool STRICT>
__global__ void __launch_bounds__(LIN_THREADS, DP <= 16 ? 2 : 1) k_linear_warp(const LinearArgs a)
{
    extern __shared__ __align__(16) double sh_tab[];
    if (!a.B) icdf_stage(sh_tab, a.icdf, threadIdx.x, LIN_THREADS);
    __syncthreads();
    constexpr int PPW = 32 / DP;                                 // paths per warp
    const int d = a.d, nB = a.nB, nsteps = a.nt - 1;
    const int lane = threadIdx.x & 31, q = lane / DP, j = lane - q * DP;
    const bool row = j < d;                                      // padding lanes compute nothing that is kept
    double Ar[DP], Gr[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) {
        Ar[k] = (row && k < d) ? __ldg(&a.A[(size_t)k * d + j]) : 0.0;
        Gr[k] = (row && k < nB) ? __ldg(&a.G[(size_t)k * d + j]) : 0.0;
    }
    const bool heun = a.scheme == SDEGPU_HEUN;
    const uint64_t nwarps = (uint64_t)gridDim.x * (LIN_THREADS / 32);
    const uint64_t warp = (uint64_t)blockIdx.x * (LIN_THREADS / 32) + (threadIdx.x >> 5);
    const uint64_t ngroups = (a.npaths + PPW - 1) / PPW;
    double mom[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (uint64_t grp = warp; grp < ngroups; grp += nwarps) {
        const uint64_t p = grp * PPW + q;
        const bool valid = p < a.npaths, mine = valid && row;
        const uint64_t pc = valid ? p : a.npaths - 1;            // padding paths recompute the last one and discard it
        double x = row ? (a.x0 ? a.x0[pc * d + j] : a.x0_shared[j]) : 0.0;
        if (mine && a.X) a.X[(p * d + j) * (uint64_t)a.nt] = x;                  // X[1,] <- x0 (:441)
        double zq[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < nsteps; ++i) {
            const double2 ds = __ldg(&a.dtsq[i]);
            double dBj = 0.0;                                    // increment of channel j (lanes j < nB)
            if (j < nB) {
                if (a.B) {
                    const double *b = a.B + (pc * nB + j) * (uint64_t)a.nt + i;
                    dBj = __dsub_rn(__ldg(b + 1), __ldg(b));                      // dB <- apply(B,2,diff) (:437)
                } else if (nB == 1) {
                    if ((i & 3) == 0) normal_quad(a.seed, a.path_offset + pc, (uint32_t)(i >> 2), 0u, sh_tab, zq);
                    const int s4 = i & 3;
                    dBj = __dmul_rn(ds.y, s4 == 0 ? zq[0] : (s4 == 1 ? zq[1] : (s4 == 2 ? zq[2] : zq[3])));
                }
            }
            if (!a.B && nB > 1) {                                // one Philox block per channel pair: the even lane draws
                double z0 = 0.0, z1 = 0.0;                       // both normals and hands the second to its neighbour
                if (j < nB && (j & 1) == 0)
                    normal_pair(a.seed, a.path_offset + pc, (uint32_t)i, 0x80000000u + (uint32_t)(j >> 1), sh_tab, z0, z1);
                const double zn = __shfl_sync(0xffffffffu, z1, lane & ~1);
                if (j < nB) dBj = __dmul_rn(ds.y, (j & 1) ? zn : z0);
            }
            const int base = q * DP;
            // Y = (x + (A x) dt) + G dB        (euler :453 / heun predictor :318), accumulation order k = 0..d-1
            double f = 0.0, gdb = 0.0;
            if (STRICT) {
#pragma unroll
                for (int k = 0; k < DP; ++k) {
                    const double xk = __shfl_sync(0xffffffffu, x, base + k);
                    if (k == 0) f = __dmul_rn(Ar[0], xk);
                    else if (k < d) f = madd<true>(f, Ar[k], xk);
                }
#pragma unroll
                for (int k = 0; k < DP; ++k) {
                    const double bk = __shfl_sync(0xffffffffu, dBj, base + k);
                    if (k == 0) gdb = __dmul_rn(Gr[0], bk);
                    else if (k < nB) gdb = madd<true>(gdb, Gr[k], bk);
                }
            } else {                                             // FAST: four partial sums (padding rows/columns of A, G are 0)
                double f4[4] = {0.0, 0.0, 0.0, 0.0}, g4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int k = 0; k < DP; ++k) {
                    f4[k & 3] = fma(Ar[k], __shfl_sync(0xffffffffu, x, base + k), f4[k & 3]);
                    g4[k & 3] = fma(Gr[k], __shfl_sync(0xffffffffu, dBj, base + k), g4[k & 3]);
                }
                f = (f4[0] + f4[1]) + (f4[2] + f4[3]);
                gdb = (g4[0] + g4[1]) + (g4[2] + g4[3]);
            }
            const double c0 = (a.forcing && row) ? __ldg(a.forcing + (size_t)j * a.nt + i) : 0.0;
            const double c1 = (a.forcing && row) ? __ldg(a.forcing + (size_t)j * a.nt + i + 1) : 0.0;
            if (a.forcing) f = __dadd_rn(f, c0);                 // A %*% x + c(t_i)
            double y = STRICT ? __dadd_rn(__dadd_rn(x, __dmul_rn(f, ds.x)), gdb) : (x + f * ds.x) + gdb;
            if (heun) {                                          // corrector (:323), gX + gY = G + G
                double fy = 0.0, g2 = 0.0;
#pragma unroll
                for (int k = 0; k < DP; ++k) {
                    const double yk = __shfl_sync(0xffffffffu, y, base + k);
                    if (k == 0) fy = STRICT ? __dmul_rn(Ar[0], yk) : Ar[0] * yk;
                    else if (k < d) fy = madd<STRICT>(fy, Ar[k], yk);
                }
                if (a.forcing) fy = __dadd_rn(fy, c1);           // ff(times[i+1], Y)  (:321)
#pragma unroll
                for (int k = 0; k < DP; ++k) {
                    const double bk = __shfl_sync(0xffffffffu, dBj, base + k);
                    const double gk = STRICT ? __dadd_rn(Gr[k], Gr[k]) : Gr[k] + Gr[k];
                    if (k == 0) g2 = STRICT ? __dmul_rn(gk, bk) : gk * bk;
                    else if (k < nB) g2 = madd<STRICT>(g2, gk, bk);
                }
                y = STRICT ? __dadd_rn(__dadd_rn(x, __dmul_rn(__dmul_rn(0.5, __dadd_rn(f, fy)), ds.x)), __dmul_rn(0.5, g2))
                           : (x + (0.5 * (f + fy)) * ds.x) + 0.5 * g2;
            }
            x = row ? y : 0.0;
            if (mine && a.X) a.X[(p * d + j) * (uint64_t)a.nt + i + 1] = x;
        }
        if (mine) {
            if (a.XT) a.XT[p * d + j] = x;
            if (a.moments) {
                const double dd = x - a.center[j];
                if (isfinite(dd)) {
                    const double d2 = dd * dd;
                    mom[0] += 1.0; mom[1] += dd; mom[2] += d2; mom[3] += d2 * dd; mom[4] += d2 * d2;
                }
            }
        }
    }
    if (a.moments) {                                             // lanes (q, j) -> component j: paths q in order, then warps in index order
        __shared__ double sh_mom[LIN_THREADS / 32][DP * 5];
        __syncthreads();
#pragma unroll
        for (int m = 0; m < 5; ++m) {
            double v = mom[m];
#pragma unroll
            for (int qq = 1; qq < PPW; ++qq) {
                const double o = __shfl_sync(0xffffffffu, mom[m], qq * DP + j);
                if (q == 0) v += o;
            }
            if (q == 0) sh_mom[threadIdx.x >> 5][j * 5 + m] = v;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < d * 5; e += LIN_THREADS) {
            double v = 0.0;
            for (int w = 0; w < LIN_THREADS / 32; ++w) v += sh_mom[w][e];
            a.mom_partials[(size_t)blockIdx.x * 5 * d + e] = v;
        }
    }
}

template <int DP>
static int launch_warp_dp(const LinearArgs &a, int sm_count, cudaStream_t stream)
{
    auto ks = k_linear_warp<DP, true>;
    auto kf = k_linear_warp<DP, false>;
    const size_t smem = a.B ? 0 : (size_t)ICDF_TABLE_BYTES;
    if (smem > 48 * 1024) {
        cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, a.strict ? ks : kf, LIN_THREADS, smem);
    uint64_t grid = (uint64_t)(per_sm < 1 ? 1 : per_sm) * sm_count;
    const uint64_t ngroups = (a.npaths + (32 / DP) - 1) / (32 / DP), nblk = (ngroups + LIN_THREADS / 32 - 1) / (LIN_THREADS / 32);
    if (grid > nblk) grid = nblk;
    if (a.moments && grid > (uint64_t)a.mom_capacity) grid = (uint64_t)a.mom_capacity;
    if (a.strict) ks<<<(int)grid, LIN_THREADS, smem, stream>>>(a);
    else kf<<<(int)grid, LIN_THREADS, smem, stream>>>(a);
    if (a.moments) reduce_moment_partials(a, (int)grid, stream);
    return (int)cudaGetLastError();
}

template <int D, int NB>
static int launch_small_dn(const LinearArgs &a, int sm_count, cudaStream_t stream)
{
    auto ks = k_linear_small<D, NB, true>;
    auto kf = k_linear_small<D, NB, false>;
    const size_t smem = a.B ? 0 : (size_t)ICDF_TABLE_BYTES;
    if (smem > 48 * 1024) {
        cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, a.strict ? ks : kf, LIN_THREADS, smem);
    uint64_t grid = (uint64_t)(per_sm < 1 ? 1 : per_sm) * sm_count;
    const uint64_t nblk = (a.npaths + LIN_THREADS - 1) / LIN_THREADS;
    if (grid > nblk) grid = nblk;
    if (a.moments && grid > (uint64_t)a.mom_capacity) grid = (uint64_t)a.mom_capacity;
    if (a.strict) ks<<<(int)grid, LIN_THREADS, smem, stream>>>(a);
    else kf<<<(int)grid, LIN_THREADS, smem, stream>>>(a);
    if (a.moments) reduce_moment_partials(a, (int)grid, stream);
    return (int)cudaGetLastError();
}

template <int D>
static int launch_small_d(const LinearArgs &a, int sm_count, cudaStream_t stream)
{
    switch (a.nB) {
    case 1: return launch_small_dn<D, 1>(a, sm_count, stream);
    case 2: return launch_small_dn<D, 2>(a, sm_count, stream);
    case 3: return launch_small_dn<D, 3>(a, sm_count, stream);
    default: return launch_small_dn<D, 4>(a, sm_count, stream);
    }
}

// moments[v] <- (accumulate ? moments[v] : 0) + sum_b partials[b][v], b in block order
__global__ void k_reduce_moment_partials(const double *__restrict__ partials, int nblocks, int nvals, double *__restrict__ out, int accumulate)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvals) return;
    double s = accumulate ? out[v] : 0.0;
    for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * nvals + v];
    out[v] = s;
}

int reduce_moment_partials(const LinearArgs &a, int nblocks, cudaStream_t stream)
{
    const int nvals = 5 * a.d;
    k_reduce_moment_partials<<<(nvals + 127) / 128, 128, 0, stream>>>(a.mom_partials, nblocks, nvals, a.moments, a.accumulate);
    return (int)cudaGetLastError();
}

int launch_linear(const LinearArgs &a, int sm_count, cudaStream_t stream)
{
    if (a.d <= LIN_SMALL_MAX && a.nB <= LIN_SMALL_MAX && a.npaths > 0) {
        switch (a.d) {
        case 1: return launch_small_d<1>(a, sm_count, stream);
        case 2: return launch_small_d<2>(a, sm_count, stream);
        case 3: return launch_small_d<3>(a, sm_count, stream);
        default: return launch_small_d<4>(a, sm_count, stream);
        }
    }
    if (a.d <= 32 && a.nB <= a.d && a.npaths > 0) {
        if (a.d <= 8) return launch_warp_dp<8>(a, sm_count, stream);
        if (a.d <= 16) return launch_warp_dp<16>(a, sm_count, stream);
        return launch_warp_dp<32>(a, sm_count, stream);
    }
    int PT = LIN_THREADS / a.d;                        // paths per block: one thread per (path, row) when d <= 256
    if (PT < 1) PT = 1;
    if (PT > 64) PT = 64;
    auto need = [&](int pt) {
        return sizeof(double) * ((size_t)3 * pt * a.d + (size_t)2 * pt * a.nB + 2) + (a.B ? 0 : (size_t)ICDF_TABLE_BYTES);
    };
    size_t smem = need(PT);
    while (smem > 200 * 1024 && PT > 1) { PT /= 2; smem = need(PT); }
    if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
    const uint64_t ngroups = (a.npaths + PT - 1) / PT;
    auto ks = k_linear_generic<true>;
    auto kf = k_linear_generic<false>;
    if (smem > 48 * 1024) {
        cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, a.strict ? ks : kf, LIN_THREADS, smem);
    uint64_t grid = (uint64_t)(per_sm < 1 ? 1 : per_sm) * sm_count;
    if (grid > ngroups) grid = ngroups;
    if (a.moments && grid > (uint64_t)a.mom_capacity) grid = (uint64_t)a.mom_capacity;
    if (a.strict) ks<<<(int)grid, LIN_THREADS, smem, stream>>>(a, PT);
    else kf<<<(int)grid, LIN_THREADS, smem, stream>>>(a, PT);
    if (a.moments) reduce_moment_partials(a, (int)grid, stream);
    return (int)cudaGetLastError();
}

}  // namespace sdegpu
