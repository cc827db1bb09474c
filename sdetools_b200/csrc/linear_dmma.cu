// High-dimensional linear SDE  dX = A X dt + diag(g) dB  (BASELINE.json config 4: d = 256, 10^6 paths).
// Across a tile of paths the drift `A %*% x` (R/sdetools.R:360-368, vignettes/NoisyOscillator.Rmd:38-54)
// is a GEMM  Y[d x P] = A[d x d] . X[d x P], so it runs on the FP64 tensor path:
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).  tcgen05 has no FP64 kind, so this legacy-encoded DMMA *is*
// Blackwell's FP64 tensor instruction.
//
//  * one CTA owns 64 paths for all time steps; X (d x 64, padded to 68) stays in shared memory,
//    the accumulators Y stay in registers (4 x 8 DMMA tiles = 32 rows x 64 paths per warp);
//  * A is re-laid out once per run into per-warp fragment order, so each k-step is four fully
//    coalesced 256-byte loads per warp that stream from L2 (512 KB per CTA per step), prefetched
//    one k-step ahead under 32 DMMAs;
//  * the Euler update  X' = (X + Y dt) + g sqrt(dt) Z  (R/sdetools.R:453) is the epilogue of the GEMM;
//    Z comes from the fused Philox generator, one block per (path, step, channel pair), the two
//    rows of a pair live in lanes L and L^4 and swap one normal each by shuffle.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <stdint.h>

#include "../../include/sdegpu.h"
#include "linear.cuh"

namespace sdegpu {

constexpr int DM_PT = 64;            // paths per CTA
constexpr int DM_LDX = DM_PT + 4;    // leading dimension of the X tile: 4 mod 16 keeps B-fragment loads conflict-free
constexpr int DM_MA = 4, DM_NB = 8;  // DMMA tiles per warp: 32 rows x 64 paths

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// M (dr x kc, column-major, leading dimension dr), zero-padded to d rows (a multiple of 32) x 4*KK columns
//   -> fragment order: Mpk[((w*KK + kk)*MA + mi)*32 + lane] = M[32w + 8mi + lane/4, 4kk + lane%4]
__global__ void k_pack_A(const double *__restrict__ A, double *__restrict__ Apk, int d, int dr, int KK, int kc)
{
    const size_t n = (size_t)d * 4 * KK;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int lane = (int)(idx & 31);
        size_t t = idx >> 5;
        const int mi = (int)(t % DM_MA); t /= DM_MA;
        const int kk = (int)(t % KK);
        const int w = (int)(t / KK);
        const int row = 32 * w + 8 * mi + (lane >> 2), col = 4 * kk + (lane & 3);
        Apk[idx] = (row < dr && col < kc) ? A[(size_t)col * dr + row] : 0.0;
    }
}

template <int NWARPS>
__global__ void __launch_bounds__(32 * NWARPS, 1) k_linear_dmma(const LinearArgs a, const double *__restrict__ Apk,
                                                                const double *__restrict__ gdiag)
{
    constexpr int D = 32 * NWARPS, KK = D / 4;
    extern __shared__ __align__(16) double sm[];
    double *sh_tab = sm;                                         // quantile table
    double *xs = sm + ICDF_TABLE_DOUBLES;                        // X tile, xs[k*DM_LDX + n]
    const int tid = threadIdx.x, lane = tid & 31, w = __shfl_sync(0xffffffffu, tid >> 5, 0), gid = lane >> 2, tig = lane & 3;   // w through a shuffle: known to be warp-uniform
    icdf_stage(sh_tab, a.icdf, tid, 32 * NWARPS);
    const int nsteps = a.nt - 1, dr = a.d;                       // dr: the system's dimension; rows dr..D-1 are zero padding
    const uint64_t ngroups = (a.npaths + DM_PT - 1) / DM_PT;
    const double *Aw = Apk + (size_t)w * KK * DM_MA * 32 + lane;
    double grow[DM_MA];                                          // diffusion intensity of this lane's rows
#pragma unroll
    for (int mi = 0; mi < DM_MA; ++mi) grow[mi] = gdiag[32 * w + 8 * mi + gid];

    for (uint64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const uint64_t p0 = grp * DM_PT;
        const int np = (int)((a.npaths - p0 < (uint64_t)DM_PT) ? (a.npaths - p0) : DM_PT);
        __syncthreads();
        for (int e = tid; e < D * DM_PT; e += 32 * NWARPS) {
            const int n = e / D, k = e - n * D;                  // consecutive threads walk one path's state vector
            xs[k * DM_LDX + n] = (n < np && k < dr) ? (a.x0 ? a.x0[(p0 + n) * dr + k] : a.x0_shared[k]) : 0.0;
        }
        __syncthreads();
        for (int i = 0; i < nsteps; ++i) {
            const double2 ds = __ldg(&a.dtsq[i]);
            double acc[DM_MA][DM_NB][2];
#pragma unroll
            for (int mi = 0; mi < DM_MA; ++mi)
#pragma unroll
                for (int ni = 0; ni < DM_NB; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
            double an[DM_MA];
#pragma unroll
            for (int mi = 0; mi < DM_MA; ++mi) an[mi] = __ldg(Aw + mi * 32);
#pragma unroll 2
            for (int kk = 0; kk < KK; ++kk) {
                double ac[DM_MA], b[DM_NB];
#pragma unroll
                for (int mi = 0; mi < DM_MA; ++mi) ac[mi] = an[mi];
                if (kk + 1 < KK) {
#pragma unroll
                    for (int mi = 0; mi < DM_MA; ++mi) an[mi] = __ldg(Aw + ((size_t)(kk + 1) * DM_MA + mi) * 32);
                }
                const double *xb = xs + (4 * kk + tig) * DM_LDX + gid;
#pragma unroll
                for (int ni = 0; ni < DM_NB; ++ni) b[ni] = xb[8 * ni];
#pragma unroll
                for (int mi = 0; mi < DM_MA; ++mi)
#pragma unroll
                    for (int ni = 0; ni < DM_NB; ++ni) dmma884(acc[mi][ni], ac[mi], b[ni]);
            }
            __syncthreads();                                     // every warp is done reading X_i
            // epilogue: X_{i+1} = (X_i + Y dt) + g sqrt(dt) Z      rows 32w+8mi+gid, paths 8ni+2tig+{0,1}
#pragma unroll
            for (int mi = 0; mi < DM_MA; ++mi) {
                const int r = 32 * w + 8 * mi + gid;
                const double gs = grow[mi] * ds.y;
#pragma unroll
                for (int ni = 0; ni < DM_NB; ++ni) {
                    const int c = 8 * ni + 2 * tig;
                    // channel pair (r & ~1, r | 1): even-row lane draws for path c, odd-row lane for path c+1
                    double za, zb;
                    normal_pair(a.seed, a.path_offset + p0 + c + (gid & 1), (uint32_t)i, 0x80000000u + (uint32_t)(r >> 1), sh_tab, za, zb);
                    // even lane holds {z(r,c), z(r+1,c)}, odd lane holds {z(r-1,c+1), z(r,c+1)}
                    const double send = (gid & 1) ? za : zb;
                    const double recv = __shfl_xor_sync(0xffffffffu, send, 4);
                    const double z0 = (gid & 1) ? recv : za;     // this lane's row, path c
                    const double z1 = (gid & 1) ? zb : recv;     // this lane's row, path c+1
                    double2 xo = *reinterpret_cast<double2 *>(xs + r * DM_LDX + c);
                    xo.x = (xo.x + acc[mi][ni][0] * ds.x) + gs * z0;
                    xo.y = (xo.y + acc[mi][ni][1] * ds.x) + gs * z1;
                    *reinterpret_cast<double2 *>(xs + r * DM_LDX + c) = xo;
                }
            }
            __syncthreads();
        }
        // terminal state and per-component power sums
        for (int e = tid; e < D * np; e += 32 * NWARPS) {
            const int n = e / D, k = e - n * D;
            const double v = xs[k * DM_LDX + n];
            if (a.XT && k < dr) a.XT[(p0 + n) * dr + k] = v;
        }
        if (a.moments) tile_moments(a, xs, DM_LDX, np, tid, 32 * NWARPS, grp == blockIdx.x);
    }
}

// Dense G (correlated noise; the exact-transition model of laws.linear_exact): the increment term is a second GEMM,
// G (d x nB) times a tile Z (nB x 64) of scaled normals kept in shared memory next to X, accumulated into the same
// registers:  acc = (A X) dt + G (sqrt(dt) Z),  X' = X + acc.  Same generator stream as every other linear kernel
// (one Philox block per path, step and channel pair).  Two tiles in shared memory: d + nB <= ~320 rows.
template <int NWARPS>
__global__ void __launch_bounds__(32 * NWARPS, 1) k_linear_dmma_dense(const LinearArgs a, const double *__restrict__ Apk,
                                                                      const double *__restrict__ Gpk, const int KKG)
{
    constexpr int D = 32 * NWARPS, KK = D / 4;
    extern __shared__ __align__(16) double sm[];
    double *sh_tab = sm;
    double *xs = sm + ICDF_TABLE_DOUBLES;                        // X tile, xs[k*DM_LDX + n]
    double *zs = xs + D * DM_LDX;                                // Z tile, zs[k*DM_LDX + n], k < 4*KKG
    const int tid = threadIdx.x, lane = tid & 31, w = __shfl_sync(0xffffffffu, tid >> 5, 0), gid = lane >> 2, tig = lane & 3;   // w through a shuffle: known to be warp-uniform
    icdf_stage(sh_tab, a.icdf, tid, 32 * NWARPS);
    const int nsteps = a.nt - 1, dr = a.d, nB = a.nB, npairs = 2 * KKG;            // channel pairs incl. zero padding
    const uint64_t ngroups = (a.npaths + DM_PT - 1) / DM_PT;
    const double *Aw = Apk + (size_t)w * KK * DM_MA * 32 + lane;
    const double *Gw = Gpk + (size_t)w * KKG * DM_MA * 32 + lane;
    for (uint64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const uint64_t p0 = grp * DM_PT;
        const int np = (int)((a.npaths - p0 < (uint64_t)DM_PT) ? (a.npaths - p0) : DM_PT);
        __syncthreads();
        for (int e = tid; e < D * DM_PT; e += 32 * NWARPS) {
            const int n = e / D, k = e - n * D;
            xs[k * DM_LDX + n] = (n < np && k < dr) ? (a.x0 ? a.x0[(p0 + n) * dr + k] : a.x0_shared[k]) : 0.0;
        }
        for (int i = 0; i < nsteps; ++i) {
            const double2 ds = __ldg(&a.dtsq[i]);
            // Z tile: sqrt(dt) times the normals of (channel pair, path); padding channels and paths are zero
            for (int e = tid; e < npairs * DM_PT; e += 32 * NWARPS) {
                const int kp = e / DM_PT, n = e - kp * DM_PT;
                double z0 = 0.0, z1 = 0.0;
                if (nB == 1) {                                   // a single channel draws from the scalar models' stream
                    if (kp == 0 && n < np) {
                        double zq[4];
                        normal_quad(a.seed, a.path_offset + p0 + n, (uint32_t)(i >> 2), 0u, sh_tab, zq);
                        const int s4 = i & 3;
                        z0 = s4 == 0 ? zq[0] : (s4 == 1 ? zq[1] : (s4 == 2 ? zq[2] : zq[3]));
                    }
                } else if (2 * kp < nB && n < np)
                    normal_pair(a.seed, a.path_offset + p0 + n, (uint32_t)i, 0x80000000u + (uint32_t)kp, sh_tab, z0, z1);
                zs[(2 * kp) * DM_LDX + n] = ds.y * z0;
                zs[(2 * kp + 1) * DM_LDX + n] = (2 * kp + 1 < nB) ? ds.y * z1 : 0.0;
            }
            __syncthreads();                                     // X_i and Z_i are in place
            double acc[DM_MA][DM_NB][2];
#pragma unroll
            for (int mi = 0; mi < DM_MA; ++mi)
#pragma unroll
                for (int ni = 0; ni < DM_NB; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
            double an[DM_MA];
#pragma unroll
            for (int mi = 0; mi < DM_MA; ++mi) an[mi] = __ldg(Aw + mi * 32);
#pragma unroll 2
            for (int kk = 0; kk < KK; ++kk) {
                double ac[DM_MA], b[DM_NB];
#pragma unroll
                for (int mi = 0; mi < DM_MA; ++mi) ac[mi] = an[mi];
                if (kk + 1 < KK) {
#pragma unroll
                    for (int mi = 0; mi < DM_MA; ++mi) an[mi] = __ldg(Aw + ((size_t)(kk + 1) * DM_MA + mi) * 32);
                }
                const double *xb = xs + (4 * kk + tig) * DM_LDX + gid;
#pragma unroll
                for (int ni = 0; ni < DM_NB; ++ni) b[ni] = xb[8 * ni];
#pragma unroll
                for (int mi = 0; mi < DM_MA; ++mi)
#pragma unroll
                    for (int ni = 0; ni < DM_NB; ++ni) dmma884(acc[mi][ni], ac[mi], b[ni]);
            }
#pragma unroll
            for (int mi = 0; mi < DM_MA; ++mi)
#pragma unroll
                for (int ni = 0; ni < DM_NB; ++ni) { acc[mi][ni][0] *= ds.x; acc[mi][ni][1] *= ds.x; }
            for (int kk = 0; kk < KKG; ++kk) {
                double ac[DM_MA], b[DM_NB];
#pragma unroll
                for (int mi = 0; mi < DM_MA; ++mi) ac[mi] = __ldg(Gw + ((size_t)kk * DM_MA + mi) * 32);
                const double *zb = zs + (4 * kk + tig) * DM_LDX + gid;
#pragma unroll
                for (int ni = 0; ni < DM_NB; ++ni) b[ni] = zb[8 * ni];
#pragma unroll
                for (int mi = 0; mi < DM_MA; ++mi)
#pragma unroll
                    for (int ni = 0; ni < DM_NB; ++ni) dmma884(acc[mi][ni], ac[mi], b[ni]);
            }
            __syncthreads();                                     // every warp is done reading X_i and Z_i
#pragma unroll
            for (int mi = 0; mi < DM_MA; ++mi) {
                const int r = 32 * w + 8 * mi + gid;
#pragma unroll
                for (int ni = 0; ni < DM_NB; ++ni) {
                    const int c = 8 * ni + 2 * tig;
                    double2 xo = *reinterpret_cast<double2 *>(xs + r * DM_LDX + c);
                    xo.x += acc[mi][ni][0];
                    xo.y += acc[mi][ni][1];
                    *reinterpret_cast<double2 *>(xs + r * DM_LDX + c) = xo;
                }
            }
        }
        __syncthreads();
        for (int e = tid; e < D * np; e += 32 * NWARPS) {
            const int n = e / D, k = e - n * D;
            if (a.XT && k < dr) a.XT[(p0 + n) * dr + k] = xs[k * DM_LDX + n];
        }
        if (a.moments) tile_moments(a, xs, DM_LDX, np, tid, 32 * NWARPS, grp == blockIdx.x);
    }
}

static size_t dense_smem(int dp, int kkg)
{
    return sizeof(double) * ((size_t)ICDF_TABLE_DOUBLES + (size_t)(dp + 4 * kkg) * DM_LDX);
}

// 0: not eligible, 1: diagonal G (normals in the GEMM epilogue), 2: dense G (second GEMM on a tile of normals)
int dmma_mode(const LinearArgs &a, const double *hostG)
{
    if (a.B || a.forcing || a.scheme != SDEGPU_EULER || a.X || a.d <= 32 || a.d > 256 || a.nB < 1) return 0;
    if (const char *e = getenv("SDEGPU_LINEAR_DMMA")) if (atoi(e) == 0) return 0;          // 0: leave every size to k_linear_tc; 1: every eligible size here
    bool diag = a.nB == a.d;
    for (int c = 0; c < a.nB && diag; ++c)
        for (int r = 0; r < a.d; ++r)
            if (r != c && hostG[(size_t)c * a.d + r] != 0.0) { diag = false; break; }
    if (diag) return 1;
    const int dp = (a.d + 31) / 32 * 32, kkg = (a.nB + 3) / 4;
    return dense_smem(dp, kkg) <= 226 * 1024 ? 2 : 0;
}

// scratch: dp*dp + dp*4*kkg doubles (packed A, packed G)
int launch_linear_dmma_dense(const LinearArgs &a, double *scratch, int sm_count, cudaStream_t stream)
{
    const int dr = a.d, d = (dr + 31) / 32 * 32, kkg = (a.nB + 3) / 4;
    double *Apk = scratch, *Gpk = scratch + (size_t)d * d;
    k_pack_A<<<sm_count * 2, 256, 0, stream>>>(a.A, Apk, d, dr, d / 4, dr);
    k_pack_A<<<sm_count * 2, 256, 0, stream>>>(a.G, Gpk, d, dr, kkg, a.nB);
    const size_t smem = dense_smem(d, kkg);
    const uint64_t ngroups = (a.npaths + DM_PT - 1) / DM_PT;
#define SDEGPU_DMMA_DENSE_CASE(NW)                                                                             \
    case NW: {                                                                                                 \
        auto k = k_linear_dmma_dense<NW>;                                                                      \
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                       \
        int per_sm = 1;                                                                                        \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 32 * NW, smem);                              \
        const uint64_t cap = (uint64_t)(per_sm < 1 ? 1 : per_sm) * sm_count;                                   \
        int grid = (int)(ngroups < cap ? ngroups : cap);                                                       \
        if (a.moments && grid > a.mom_capacity) grid = a.mom_capacity;                                         \
        k<<<grid, 32 * NW, smem, stream>>>(a, Apk, Gpk, kkg);                                                  \
        if (a.moments) reduce_moment_partials(a, grid, stream);                                                \
    } break;
    switch (d / 32) {
        SDEGPU_DMMA_DENSE_CASE(2) SDEGPU_DMMA_DENSE_CASE(3) SDEGPU_DMMA_DENSE_CASE(4) SDEGPU_DMMA_DENSE_CASE(5)
        SDEGPU_DMMA_DENSE_CASE(6) SDEGPU_DMMA_DENSE_CASE(7) SDEGPU_DMMA_DENSE_CASE(8)
    default: return (int)cudaErrorInvalidValue;
    }
#undef SDEGPU_DMMA_DENSE_CASE
    return (int)cudaGetLastError();
}

bool dmma_eligible(const LinearArgs &a, const double *hostG) { return dmma_mode(a, hostG) == 1; }

// scratch: dp*dp + dp doubles (packed A, diagonal of G), dp = d rounded up to a multiple of 32
int launch_linear_dmma(const LinearArgs &a, double *scratch, int sm_count, cudaStream_t stream)
{
    const int dr = a.d, d = (dr + 31) / 32 * 32;
    double *Apk = scratch, *gdiag = scratch + (size_t)d * d;
    k_pack_A<<<sm_count * 2, 256, 0, stream>>>(a.A, Apk, d, dr, d / 4, dr);
    cudaMemsetAsync(gdiag, 0, sizeof(double) * d, stream);
    cudaMemcpy2DAsync(gdiag, sizeof(double), a.G, sizeof(double) * (dr + 1), sizeof(double), dr, cudaMemcpyDeviceToDevice, stream);
    const size_t smem = sizeof(double) * ((size_t)ICDF_TABLE_DOUBLES + (size_t)d * DM_LDX);
    const uint64_t ngroups = (a.npaths + DM_PT - 1) / DM_PT;
    // small d leaves room for several CTAs per SM (d = 64: two warps per CTA): as many as registers and shared memory allow
#define SDEGPU_DMMA_CASE(NW)                                                                                   \
    case NW: {                                                                                                 \
        auto k = k_linear_dmma<NW>;                                                                            \
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                       \
        int per_sm = 1;                                                                                        \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 32 * NW, smem);                              \
        const uint64_t cap = (uint64_t)(per_sm < 1 ? 1 : per_sm) * sm_count;                                   \
        int grid = (int)(ngroups < cap ? ngroups : cap);                                                       \
        if (a.moments && grid > a.mom_capacity) grid = a.mom_capacity;                                         \
        k<<<grid, 32 * NW, smem, stream>>>(a, Apk, gdiag);                                                     \
        if (a.moments) reduce_moment_partials(a, grid, stream);                                                \
    } break;
    switch (d / 32) {
        SDEGPU_DMMA_CASE(1) SDEGPU_DMMA_CASE(2) SDEGPU_DMMA_CASE(3) SDEGPU_DMMA_CASE(4)
        SDEGPU_DMMA_CASE(5) SDEGPU_DMMA_CASE(6) SDEGPU_DMMA_CASE(7) SDEGPU_DMMA_CASE(8)
    default: return (int)cudaErrorInvalidValue;
    }
#undef SDEGPU_DMMA_CASE
    return (int)cudaGetLastError();
}

}  // namespace sdegpu
