// Ensemble kernels: P independent euler()/heun() calls, one thread per path,
// state in registers across all steps (replaces the sapply/replicate idiom around
// R/sdetools.R:375-470 / :240-338, e.g. vignettes/Killing.Rmd:183-189).
//
//   k_terminal : increments from the fused Philox generator, nothing but terminal
//                state / power sums / histogram leaves the SM   (FP64-pipe bound)
//   k_tiled    : supplied B read through shared-memory tiles and/or the full
//                trajectory X written through shared-memory tiles (HBM bound)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "models.cuh"
#include "philox.cuh"

namespace sdegpu {

constexpr int MAX_NX = 2;           // catalogue models handled here have nx <= 2
#ifndef SDEGPU_TERMINAL_THREADS
#define SDEGPU_TERMINAL_THREADS 640
#endif
#ifndef SDEGPU_TERMINAL_MIN_BLOCKS
#define SDEGPU_TERMINAL_MIN_BLOCKS 1
#endif
constexpr int TERMINAL_THREADS = SDEGPU_TERMINAL_THREADS;
constexpr int TILE_PATHS = 128;     // paths (= threads) per block in k_tiled
constexpr int TILE_STEPS = 32;      // time steps staged per tile
constexpr int TILE_LD = TILE_STEPS + 1;   // odd leading dimension: conflict-free column walks

struct RunArgs {
    KParams P;
    const double2 *dtsq;            // [nt-1] {dt_i, sqrt(dt_i)},  dt_i = times[i+1]-times[i]  (:446)
    const double *icdf;             // device, half-normal quantile table (philox.cuh)
    const double *x0;               // per-path x0 (device) or nullptr
    double x0s[MAX_NX];             // shared x0
    const double *B;                // device, B[p*nt+i]  (nB = 1)
    double *X;                      // device, X[(p*nx+j)*nt+i]
    double *XT;                     // device, XT[p*nx+j]
    double *partials;               // [gridDim.x][5*nx] block partial power sums, or nullptr
    unsigned long long *hist;       // [nbins+3] or nullptr
    double center[MAX_NX];
    double hist_lo, hist_hi, hist_invw;
    uint64_t npaths, path_offset;
    PhiloxKey seed;                 // round keys of the run's seed
    int nt, hist_nbins, hist_comp;
};

// ---- terminal statistics ---------------------------------------------------
// Histogram bin of R's hist(right=TRUE, include.lowest=TRUE) on equal-width breaks;
// index arithmetic uses individually rounded ops so the CPU oracle reproduces it.
__device__ __forceinline__ int hist_bin(double x, const RunArgs &a)
{
    if (x != x) return a.hist_nbins + 2;
    if (x < a.hist_lo) return 0;
    if (x > a.hist_hi) return a.hist_nbins + 1;
    int k = (int)ceil(__dmul_rn(__dsub_rn(x, a.hist_lo), a.hist_invw)) - 1;
    k = k < 0 ? 0 : (k >= a.hist_nbins ? a.hist_nbins - 1 : k);
    return k + 1;
}

template <int NX>
__device__ __forceinline__ void accumulate_terminal(const RunArgs &a, const double (&x)[NX], double (&acc)[NX][5],
                                                    unsigned int *sh_hist)
{
    if (a.partials) {
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            const double d = x[j] - a.center[j];
            if (isfinite(d)) {
                const double d2 = d * d;
                acc[j][0] += 1.0; acc[j][1] += d; acc[j][2] += d2; acc[j][3] += d2 * d; acc[j][4] += d2 * d2;
            }
        }
    }
    if (a.hist) atomicAdd(&sh_hist[hist_bin(x[a.hist_comp < NX ? a.hist_comp : 0], a)], 1u);
}

// Deterministic block reduction: shuffle tree per warp, then warps in index order.
template <int NX, int THREADS>
__device__ __forceinline__ void flush_block_stats(const RunArgs &a, double (&acc)[NX][5], unsigned int *sh_hist)
{
    constexpr int NW = THREADS / 32;
    __shared__ double sh_red[NW][NX * 5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (a.partials) {
#pragma unroll
        for (int j = 0; j < NX; ++j)
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                double v = acc[j][m];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
                if (lane == 0) sh_red[warp][j * 5 + m] = v;
            }
    }
    __syncthreads();
    if (a.partials && threadIdx.x < NX * 5) {
        double v = 0.0;
        for (int w = 0; w < NW; ++w) v += sh_red[w][threadIdx.x];
        a.partials[(size_t)blockIdx.x * (NX * 5) + threadIdx.x] = v;
    }
    if (a.hist)
        for (int i = threadIdx.x; i < a.hist_nbins + 3; i += THREADS)
            if (sh_hist[i]) atomicAdd(&a.hist[i], (unsigned long long)sh_hist[i]);
}

template <int NX>
__device__ __forceinline__ void load_x0(const RunArgs &a, uint64_t p, double (&x)[NX])
{
#pragma unroll
    for (int j = 0; j < NX; ++j) x[j] = a.x0 ? a.x0[p * NX + j] : a.x0s[j];
}

// ---- register-resident, fused-generator kernel ------------------------------
template <class M, int SCHEME, int PROJ>
__global__ void __launch_bounds__(TERMINAL_THREADS, SDEGPU_TERMINAL_MIN_BLOCKS) k_terminal(const RunArgs a)
{
    constexpr int NX = M::NX;
    extern __shared__ __align__(16) double sh_tab[];             // [quantile table][histogram]
    unsigned int *sh_hist = reinterpret_cast<unsigned int *>(sh_tab + ICDF_TABLE_DOUBLES);
    icdf_stage(sh_tab, a.icdf, threadIdx.x, TERMINAL_THREADS);
    if (a.hist)
        for (int i = threadIdx.x; i < a.hist_nbins + 3; i += TERMINAL_THREADS) sh_hist[i] = 0u;
    __syncthreads();
    double acc[NX][5];
#pragma unroll
    for (int j = 0; j < NX; ++j)
#pragma unroll
        for (int m = 0; m < 5; ++m) acc[j][m] = 0.0;

    const int nsteps = a.nt - 1;
    const uint64_t stride = (uint64_t)gridDim.x * TERMINAL_THREADS;
    for (uint64_t p = (uint64_t)blockIdx.x * TERMINAL_THREADS + threadIdx.x; p < a.npaths; p += stride) {
        double x[NX];
        load_x0<NX>(a, p, x);
        const uint64_t gp = a.path_offset + p;
        // Software-pipelined: the Philox block of steps (i+2, i+3) is generated while steps (i, i+1)
        // integrate, so the integer rounds issue in the shadow of the dependent FP64 chain.
        int i = 0;
        uint32_t r0, r1, r2, r3;
        philox_block(a.seed, gp, 0u, 0u, r0, r1, r2, r3);
        for (; i + 1 < nsteps; i += 2) {
            uint32_t n0, n1, n2, n3;
            philox_block(a.seed, gp, (uint32_t)(i >> 1) + 1u, 0u, n0, n1, n2, n3);
            const double z0 = icdf_normal(r0, r1, sh_tab), z1 = icdf_normal(r2, r3, sh_tab);
            const double2 d0 = __ldg(&a.dtsq[i]), d1 = __ldg(&a.dtsq[i + 1]);
            sde_step<M, SCHEME, PROJ, Fast>(a.P, x, d0.x, d0.y * z0);
            sde_step<M, SCHEME, PROJ, Fast>(a.P, x, d1.x, d1.y * z1);
            r0 = n0; r1 = n1; r2 = n2; r3 = n3;
        }
        if (i < nsteps) {
            const double z0 = icdf_normal(r0, r1, sh_tab);
            const double2 d0 = __ldg(&a.dtsq[i]);
            sde_step<M, SCHEME, PROJ, Fast>(a.P, x, d0.x, d0.y * z0);
        }
        if (a.XT) {
#pragma unroll
            for (int j = 0; j < NX; ++j) a.XT[p * NX + j] = x[j];
        }
        accumulate_terminal<NX>(a, x, acc, sh_hist);
    }
    flush_block_stats<NX, TERMINAL_THREADS>(a, acc, sh_hist);
}

// ---- full-trajectory kernel, fused generator (HBM-write bound) ----------------
// X[(p*nx+j)*nt + i] is time-fastest per path (the reference's column-major nt x nx matrix), so the
// 32 paths of a warp write 32 rows that are nt*8 bytes apart and nothing coalesces across lanes.
// Partial-sector stores make L2 fetch the rest of the sector from HBM (measured: 0.43 B read per B
// written with 16-byte stores), so each thread keeps four consecutive time points of its own row in
// registers and writes them as ONE aligned 32-byte sector (sm_100 256-bit store, STG.E.256).  Rows
// start at arbitrary element offsets (nt is usually odd), so the sector grid is shifted per row by
// d = (row start + 1) mod 4: the sector that completes in a 4-step group is made of the last d values
// of the previous group and the first 4-d of this one.  Paths are dealt to lanes with stride Q
// (Q*nx*nt = 0 mod 4) so that d is warp-uniform and the four-way switch does not diverge.  Only the
// first/last partial sector of a row is written element by element.
#ifndef SDEGPU_TRAJ_THREADS
#define SDEGPU_TRAJ_THREADS 512
#endif
constexpr int TRAJ_THREADS = SDEGPU_TRAJ_THREADS;

__device__ __forceinline__ void st_sector(double *p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

template <class M, int SCHEME, int PROJ>
__global__ void __launch_bounds__(TRAJ_THREADS, 1) k_traj(const RunArgs a, const int vec_ok, const int Q)
{
    constexpr int NX = M::NX;
    extern __shared__ __align__(16) double sh_tab[];             // [quantile table][histogram]
    unsigned int *sh_hist = reinterpret_cast<unsigned int *>(sh_tab + ICDF_TABLE_DOUBLES);
    icdf_stage(sh_tab, a.icdf, threadIdx.x, TRAJ_THREADS);
    if (a.hist)
        for (int i = threadIdx.x; i < a.hist_nbins + 3; i += TRAJ_THREADS) sh_hist[i] = 0u;
    __syncthreads();
    double acc[NX][5];
#pragma unroll
    for (int j = 0; j < NX; ++j)
#pragma unroll
        for (int m = 0; m < 5; ++m) acc[j][m] = 0.0;

    const int nsteps = a.nt - 1, ngroups = nsteps >> 2;
    const uint64_t stride = (uint64_t)gridDim.x * TRAJ_THREADS;
    const uint64_t W = 32u * (unsigned)Q;                        // paths per lane-interleaved group of Q warps
    for (uint64_t T = (uint64_t)blockIdx.x * TRAJ_THREADS + threadIdx.x;; T += stride) {
        const uint64_t within = T % W, p = (T - within) + (within >> 5) + (uint64_t)Q * (within & 31);
        if (T - within >= a.npaths) break;                       // whole group out of range (uniform)
        if (p >= a.npaths) continue;
        double x[NX], v[NX][4], pv[NX][4];
        double *row[NX];                                         // &X[row start + 1]: stream position 0
        int d[NX];
        load_x0<NX>(a, p, x);
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            const uint64_t e0 = (p * NX + j) * (uint64_t)a.nt;
            row[j] = a.X + e0 + 1;
            d[j] = vec_ok ? (int)((e0 + 1) & 3) : 4;             // 4: element-wise stores only
            if (d[j] != 1) a.X[e0] = x[j];                       // X[1,] <- x0 (:441); with d = 1 it opens the first sector
#pragma unroll
            for (int k = 0; k < 4; ++k) pv[j][k] = x[j];
        }
        const uint64_t gp = a.path_offset + p;
        for (int m = 0; m < ngroups; ++m) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {                        // one Philox block per two steps
                double z[2];
                normal_pair(a.seed, gp, 2u * (uint32_t)m + (uint32_t)h, 0u, sh_tab, z[0], z[1]);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int k = 2 * h + q;
                    const double2 dt = __ldg(&a.dtsq[4 * m + k]);
                    sde_step<M, SCHEME, PROJ, Fast>(a.P, x, dt.x, dt.y * z[q]);
#pragma unroll
                    for (int j = 0; j < NX; ++j) v[j][k] = x[j];
                }
            }
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                double *g = row[j] + 4 * (int64_t)m;             // stream position 4m
                if (d[j] == 4 || (m == 0 && d[j] >= 2)) {        // no vector store possible: the sector starts before the row
                    const int n = d[j] == 4 ? 4 : 4 - d[j];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < n) g[k] = v[j][k];
                } else if (d[j] == 0) st_sector(g, v[j][0], v[j][1], v[j][2], v[j][3]);
                else if (d[j] == 1) st_sector(g - 1, pv[j][3], v[j][0], v[j][1], v[j][2]);
                else if (d[j] == 2) st_sector(g - 2, pv[j][2], pv[j][3], v[j][0], v[j][1]);
                else st_sector(g - 3, pv[j][1], pv[j][2], pv[j][3], v[j][0]);
#pragma unroll
                for (int k = 0; k < 4; ++k) pv[j][k] = v[j][k];
            }
        }
        // values of the last group that opened a sector the row does not complete in a full group
        if (ngroups > 0) {
#pragma unroll
            for (int j = 0; j < NX; ++j)
                if (d[j] >= 1 && d[j] <= 3) {
#pragma unroll
                    for (int k = 1; k < 4; ++k)
                        if (k >= 4 - d[j]) row[j][4 * (int64_t)(ngroups - 1) + k] = pv[j][k];
                }
        }
        for (int i = 4 * ngroups; i < nsteps; ++i) {              // up to three trailing steps, element-wise
            double z0, z1;
            normal_pair(a.seed, gp, (uint32_t)(i >> 1), 0u, sh_tab, z0, z1);
            const double zz = (i & 1) ? z1 : z0;
            const double2 dt = __ldg(&a.dtsq[i]);
            sde_step<M, SCHEME, PROJ, Fast>(a.P, x, dt.x, dt.y * zz);
#pragma unroll
            for (int j = 0; j < NX; ++j) row[j][i] = x[j];
        }
        if (a.XT) {
#pragma unroll
            for (int j = 0; j < NX; ++j) a.XT[p * NX + j] = x[j];
        }
        accumulate_terminal<NX>(a, x, acc, sh_hist);
    }
    flush_block_stats<NX, TRAJ_THREADS>(a, acc, sh_hist);
}

// ---- tiled kernel: supplied B in, and/or full trajectory out ----------------
// Shared memory (dynamic): [quantile table]                               if !SRC_B
//                          [hist: nbins+3 u32, padded to 16 B]
//                          [sB: TILE_PATHS x TILE_LD doubles]            if SRC_B
//                          [sX: NX x TILE_PATHS x TILE_LD doubles]       if a.X
template <class M, int SCHEME, int PROJ, class A, bool SRC_B>
__global__ void __launch_bounds__(TILE_PATHS) k_tiled(const RunArgs a, const int hist_words)
{
    constexpr int NX = M::NX;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const double *sh_tab = reinterpret_cast<const double *>(smem_raw);
    unsigned char *after_tab = smem_raw + (SRC_B ? 0 : ICDF_TABLE_BYTES);
    unsigned int *sh_hist = reinterpret_cast<unsigned int *>(after_tab);
    double *sB = reinterpret_cast<double *>(after_tab + (size_t)hist_words * 4);
    double *sX = sB + (SRC_B ? TILE_PATHS * TILE_LD : 0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!SRC_B) icdf_stage(reinterpret_cast<double *>(smem_raw), a.icdf, tid, TILE_PATHS);
    if (a.hist) {
        for (int i = tid; i < a.hist_nbins + 3; i += TILE_PATHS) sh_hist[i] = 0u;
    }
    __syncthreads();
    double acc[NX][5];
#pragma unroll
    for (int j = 0; j < NX; ++j)
#pragma unroll
        for (int m = 0; m < 5; ++m) acc[j][m] = 0.0;

    const int nsteps = a.nt - 1;
    const uint64_t ntiles = (a.npaths + TILE_PATHS - 1) / TILE_PATHS;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t p0 = tile * TILE_PATHS, p = p0 + tid;
        const bool valid = p < a.npaths;
        const int rows = (int)((a.npaths - p0 < (uint64_t)TILE_PATHS) ? (a.npaths - p0) : TILE_PATHS);
        double x[NX];
        if (valid) load_x0<NX>(a, p, x);
        else {
#pragma unroll
            for (int j = 0; j < NX; ++j) x[j] = 0.0;
        }
        if (a.X && valid) {
#pragma unroll
            for (int j = 0; j < NX; ++j) a.X[(p * NX + j) * (uint64_t)a.nt] = x[j];      // X[1,] <- x0  (:441)
        }
        const uint64_t gp = a.path_offset + p;
        double z0 = 0.0, z1 = 0.0;
        for (int i0 = 0; i0 < nsteps; i0 += TILE_STEPS) {
            const int nc = (nsteps - i0 < TILE_STEPS) ? (nsteps - i0) : TILE_STEPS;
            if (SRC_B) {
                // rows of the B tile are contiguous in time: one warp streams one row at a time
                for (int r = warp; r < rows; r += TILE_PATHS / 32) {
                    const double *src = a.B + (p0 + r) * (uint64_t)a.nt + i0;
                    for (int c = lane; c <= nc; c += 32) sB[r * TILE_LD + c] = __ldg(src + c);
                }
                __syncthreads();
            }
            if (valid) {
#pragma unroll 2
                for (int c = 0; c < nc; ++c) {
                    const int i = i0 + c;
                    const double2 d = __ldg(&a.dtsq[i]);
                    double dB;
                    if (SRC_B) {
                        dB = A::sub(sB[tid * TILE_LD + c + 1], sB[tid * TILE_LD + c]);   // dB = diff(B)  (:437)
                    } else {
                        if ((c & 1) == 0) normal_pair(a.seed, gp, (uint32_t)(i >> 1), 0u, sh_tab, z0, z1);
                        dB = d.y * ((c & 1) ? z1 : z0);
                    }
                    sde_step<M, SCHEME, PROJ, A>(a.P, x, d.x, dB);
                    if (a.X) {
#pragma unroll
                        for (int j = 0; j < NX; ++j) sX[(j * TILE_PATHS + tid) * TILE_LD + c] = x[j];
                    }
                }
            }
            if (a.X) {
                __syncthreads();
                for (int r = warp; r < rows * NX; r += TILE_PATHS / 32) {
                    const int row = r / NX, j = r - row * NX;
                    double *dst = a.X + ((p0 + row) * NX + j) * (uint64_t)a.nt + i0 + 1;
                    const double *srow = sX + (j * TILE_PATHS + row) * TILE_LD;
                    for (int c = lane; c < nc; c += 32) dst[c] = srow[c];
                }
            }
            __syncthreads();
        }
        if (valid) {
            if (a.XT) {
#pragma unroll
                for (int j = 0; j < NX; ++j) a.XT[p * NX + j] = x[j];
            }
            accumulate_terminal<NX>(a, x, acc, sh_hist);
        }
    }
    flush_block_stats<NX, TILE_PATHS>(a, acc, sh_hist);
}

// ---- host-side launch signature shared by the per-model translation units ---
struct LaunchCfg {
    int scheme, proj, arith;
    bool src_b, want_x;
    int grid;            // 0: choose from occupancy
    int sm_count;
    cudaStream_t stream;
};
// returns cudaError_t as int; *grid_used gets the grid size actually launched
typedef int (*model_launcher)(const RunArgs &, const LaunchCfg &, int *grid_used);
int launch_model(int model, const RunArgs &, const LaunchCfg &, int *grid_used);

}  // namespace sdegpu
