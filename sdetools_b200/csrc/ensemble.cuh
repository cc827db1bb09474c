// Ensemble kernels: P independent euler()/heun() calls, one thread per path,
// state in registers across all steps (replaces the sapply/replicate idiom around
// R/sdetools.R:375-470 / :240-338, e.g. vignettes/Killing.Rmd:183-189).
//
//   k_terminal : increments from the fused Philox generator, nothing but terminal
//                state / power sums / histogram leaves the SM   (FP64-pipe bound)
//   k_traj     : fused generator, full trajectory X written as 32-byte sectors (HBM-write bound)
//   k_stream   : supplied B read, X written, both as per-thread 32-byte sectors (HBM bound; the
//                bit-exact STRICT path)
//   k_functionals   : integrals along the path (QV, Ito, Stratonovich, time) accumulated in the step loop with
//                     cumsum's 80-bit accumulator, nothing stored
//   k_killed        : killing rate r / stopping rule h with trajectory (NaN after tau)
//   k_killed_refill : the same without trajectory; lanes whose path ended take the warp's next path
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "float80.cuh"
#include "models.cuh"
#include "philox.cuh"

namespace sdegpu {

constexpr int MAX_NX = 2;           // catalogue models handled here have nx <= 2
#ifndef SDEGPU_TERMINAL_THREADS
#define SDEGPU_TERMINAL_THREADS 768
#endif
#ifndef SDEGPU_TERMINAL_NP
#define SDEGPU_TERMINAL_NP 1
#endif
#ifndef SDEGPU_TERMINAL_MIN_BLOCKS
#define SDEGPU_TERMINAL_MIN_BLOCKS 1
#endif
constexpr int TERMINAL_THREADS = SDEGPU_TERMINAL_THREADS;

struct RunArgs {
    KParams P;
    const double2 *dtsq;            // [nt-1] {dt_i, sqrt(dt_i)},  dt_i = times[i+1]-times[i]  (:446)
    const double *icdf;             // device, half-normal quantile table, generic form (philox.cuh)
    const double *icdf32;           // device, quantile table of the fast evaluation (register-resident kernels)
    const double *coef;             // [nt-1][FAST_NCOEF] folded step coefficients (models.cuh FastStep), non-uniform grids
    double ucoef[FAST_NCOEF];       // the one coefficient row of a uniform grid (constant-bank operands)
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
    // mortality / stopping (k_killed only)
    int kill_kind, kill_comp, stop_kind, stop_comp;
    double kill_a, kill_b, stop_lo, stop_hi, S0;
    const double *Stau;             // per-path thresholds or nullptr (drawn from Philox)
    const double *times;            // device copy of the time grid (tau = times[i+1])
    double *tau, *S_tau;
    double *Spath;                  // device, survival path S[p*nt+i] (k_killed) or nullptr
    double *func;                   // device, path integrals func[(p*nx+j)*4 + {QV, ito, strat, time}] (k_functionals)
    double *mfunc;                  // device, model integrals mfunc[(p*nx+j)*2 + {int f(X) dt, int g(X) dB}] (k_functionals)
    // time-dependent closures tabulated on the grid (sdegpu.h ABI 5): f = f0(x) + c(t), g = s(t) g0(x)
    const double *forcing;          // device, c[j*nt + i] or nullptr
    const double *gscale;           // device, s[i] or nullptr
};

// the tabulated time dependence of step i (t_i for the predictor, t_{i+1} for Heun's corrector)
template <int NX>
__device__ __forceinline__ void load_tdep(const RunArgs &a, int i, TDep<NX> &td)
{
    td.has_c = a.forcing != nullptr;
    td.has_s = a.gscale != nullptr;
#pragma unroll
    for (int j = 0; j < NX; ++j) {
        td.cX[j] = td.has_c ? __ldg(a.forcing + (size_t)j * a.nt + i) : 0.0;
        td.cY[j] = td.has_c ? __ldg(a.forcing + (size_t)j * a.nt + i + 1) : 0.0;
    }
    td.sX = td.has_s ? __ldg(a.gscale + i) : 1.0;
    td.sY = td.has_s ? __ldg(a.gscale + i + 1) : 1.0;
}

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
    if (a.hist) {
        double xh = x[0];                                        // selected without indexing x[] by a run-time value:
#pragma unroll                                                   // that would force the state array into local memory
        for (int j = 1; j < NX; ++j) xh = a.hist_comp == j ? x[j] : xh;
        atomicAdd(&sh_hist[hist_bin(xh, a)], 1u);
    }
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

// the same for a block size known only at run time: `scratch` holds (warps) x (5*NX) doubles of shared memory
template <int NX>
__device__ __forceinline__ void flush_block_stats_dyn(const RunArgs &a, double (&acc)[NX][5], unsigned int *sh_hist, double *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();                                             // the scratch area may alias rings that are still being read
    if (a.partials) {
#pragma unroll
        for (int j = 0; j < NX; ++j)
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                double v = acc[j][m];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
                if (lane == 0) scratch[warp * (NX * 5) + j * 5 + m] = v;
            }
    }
    __syncthreads();
    if (a.partials && threadIdx.x < NX * 5) {
        double v = 0.0;
        for (int w = 0; w < nw; ++w) v += scratch[w * (NX * 5) + threadIdx.x];
        a.partials[(size_t)blockIdx.x * (NX * 5) + threadIdx.x] = v;
    }
    if (a.hist)
        for (int i = threadIdx.x; i < a.hist_nbins + 3; i += blockDim.x)
            if (sh_hist[i]) atomicAdd(&a.hist[i], (unsigned long long)sh_hist[i]);
}

template <int NX>
__device__ __forceinline__ void load_x0(const RunArgs &a, uint64_t p, double (&x)[NX])
{
#pragma unroll
    for (int j = 0; j < NX; ++j) x[j] = a.x0 ? a.x0[p * NX + j] : a.x0s[j];
}

// ---- register-resident, fused-generator kernel ------------------------------
// One thread per path, grid-stride.  Per four steps: one Philox block (generated one iteration ahead, so its integer
// rounds issue under the dependent FP64 chain), four quantile evaluations, four steps.  UNIFORM: the grid's
// increments agree to rounding (sdegpu_run decides), so the step's coefficient row is a kernel parameter and its
// noise factor is folded into the staged quantile table; otherwise rows stream from global memory (L1-resident).
#ifndef SDEGPU_TERMINAL_REPL
#define SDEGPU_TERMINAL_REPL 8
#endif
constexpr int TERMINAL_REPL = SDEGPU_TERMINAL_REPL;

struct CoefRow {                                                // a step's row loaded as three 128-bit words
    double v[FAST_NCOEF];
    __device__ __forceinline__ double operator[](int k) const { return v[k]; }
    __device__ __forceinline__ void load(const double *p)
    {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p + 2)),
                      c = __ldg(reinterpret_cast<const double2 *>(p + 4));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y;
    }
};
struct CoefUniform {                                            // reads the kernel parameter: constant-bank operands
    const double *c;
    __device__ __forceinline__ double operator[](int k) const { return c[k]; }
};

template <class M, int SCHEME, int PROJ, bool UNIFORM>
__global__ void __launch_bounds__(TERMINAL_THREADS, SDEGPU_TERMINAL_MIN_BLOCKS) k_terminal(const __grid_constant__ RunArgs a)
{
    constexpr int NX = M::NX;
    using FS = FastStep<M, SCHEME, PROJ>;
    using TL = Icdf32Layout<TERMINAL_REPL>;
    extern __shared__ __align__(16) double sh_tab[];             // [quantile table, replicated][histogram]
    unsigned int *sh_hist = reinterpret_cast<unsigned int *>(sh_tab + TL::DOUBLES);
    icdf32_stage<TERMINAL_REPL>(sh_tab, a.icdf32, UNIFORM ? a.ucoef[0] : 1.0, threadIdx.x, TERMINAL_THREADS);
    if (a.hist)
        for (int i = threadIdx.x; i < a.hist_nbins + 3; i += TERMINAL_THREADS) sh_hist[i] = 0u;
    __syncthreads();
    NormalGen<TERMINAL_REPL> gen;
    gen.init(sh_tab, a.icdf, UNIFORM ? a.ucoef[0] : 1.0);
    double acc[NX][5];
#pragma unroll
    for (int j = 0; j < NX; ++j)
#pragma unroll
        for (int m = 0; m < 5; ++m) acc[j][m] = 0.0;

    const int nsteps = a.nt - 1;
    constexpr int NP = SDEGPU_TERMINAL_NP;                       // independent paths advanced together by one thread (ILP)
    const uint64_t stride = (uint64_t)gridDim.x * TERMINAL_THREADS;
    const CoefUniform cu{a.ucoef};
    for (uint64_t pb = (uint64_t)blockIdx.x * TERMINAL_THREADS + threadIdx.x; pb < a.npaths; pb += stride * NP) {
        double x[NP][NX];
        uint64_t pq[NP];
        uint32_t r[NP][4];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            pq[q] = pb + (uint64_t)q * stride;
            const uint64_t pl = pq[q] < a.npaths ? pq[q] : pb;   // a missing partner recomputes path pb and is discarded
            load_x0<NX>(a, pl, x[q]);
            philox_block(a.seed, a.path_offset + pl, 0u, 0u, r[q][0], r[q][1], r[q][2], r[q][3]);
        }
        // one loop counter: the block index (the Philox counter of the block generated ahead is blk + 1, rows start at 4 blk)
        const uint32_t nfull = (uint32_t)nsteps >> 2;
#if defined(SDEGPU_TERMINAL_UNROLL4)
#pragma unroll 4
#elif defined(SDEGPU_TERMINAL_UNROLL3)
#pragma unroll 3
#elif !defined(SDEGPU_TERMINAL_UNROLL1)
#pragma unroll 2
#endif
        for (uint32_t blk = 0; blk < nfull; ++blk) {
            const int i = (int)(blk << 2);
            CoefRow row[4];
            if (!UNIFORM) {
#pragma unroll
                for (int k = 0; k < 4; ++k) row[k].load(a.coef + (size_t)(i + k) * FAST_NCOEF);
            }
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const uint64_t gp = a.path_offset + (pq[q] < a.npaths ? pq[q] : pb);
                uint32_t n[4];
#ifdef SDEGPU_EXP_NOPHILOX                                       // decomposition experiment: a one-IMAD word source
#pragma unroll
                for (int k = 0; k < 4; ++k) n[k] = r[q][k] * 1664525u + 1013904223u + (uint32_t)gp;
#else
                // (PhiloxPath — the path's share of the rounds computed once, 17 instead of 20 wide multiplies per block — was measured
                // here: 0.4 % slower, profiles/r2_tune_philox_path.txt; the wide multiplies are not what this kernel waits for)
                philox_block(a.seed, gp, blk + 1u, 0u, n[0], n[1], n[2], n[3]);
#endif
                double z[4];
#ifdef SDEGPU_EXP_NOICDF                                         // decomposition experiment: a uniform instead of a normal
#pragma unroll
                for (int k = 0; k < 4; ++k) z[k] = (__hiloint2double(0x43300000, (int)r[q][k]) - 4503601774854144.0) * 4.6566128730773926e-10;
#else
                gen.quad(a.seed, gp, blk, 0u, r[q], z);
#endif
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#ifdef SDEGPU_EXP_NOSTEP                                         // decomposition experiment: the step is one DADD
                    x[q][0] += z[k];
#else
                    if (UNIFORM) FS::apply(a.P, x[q], cu, z[k]);
                    else FS::apply(a.P, x[q], row[k], row[k][0] * z[k]);
#endif
                    r[q][k] = n[k];
                }
            }
        }
        const int i = (int)(nfull << 2);
        if (i < nsteps) {                                        // up to three trailing steps of the last block
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const uint64_t gp = a.path_offset + (pq[q] < a.npaths ? pq[q] : pb);
                double z[4];
                gen.quad(a.seed, gp, (uint32_t)(i >> 2), 0u, r[q], z);
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (i + k < nsteps) {
                        if (UNIFORM) FS::apply(a.P, x[q], cu, z[k]);
                        else {
                            CoefRow row;
                            row.load(a.coef + (size_t)(i + k) * FAST_NCOEF);
                            FS::apply(a.P, x[q], row, row[0] * z[k]);
                        }
                    }
            }
        }
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            if (pq[q] < a.npaths) {
                if (a.XT) {
#pragma unroll
                    for (int j = 0; j < NX; ++j) a.XT[pq[q] * NX + j] = x[q][j];
                }
                accumulate_terminal<NX>(a, x[q], acc, sh_hist);
            }
        }
    }
    flush_block_stats<NX, TERMINAL_THREADS>(a, acc, sh_hist);
}

// ---- full-trajectory kernel, fused generator (HBM-write bound) ----------------
// X[(p*nx+j)*nt + i] is time-fastest per path (the reference's column-major nt x nx matrix, R/sdetools.R:439-441,464),
// so the 32 paths of a warp own 32*nx rows that are nt*8 bytes apart and nothing coalesces across lanes while the
// paths advance.  Round 1 wrote one 32-byte sector per lane per four steps straight from registers (4.6 TB/s).  The
// store microbenchmark (profiles/r1_tma_store_microbench.txt) says what the memory system prefers: few rows in flight
// per SM (256 rather than 512) and bursts of >= 128 bytes per row — 5.1-5.3 TB/s.  So: TMA-staged.
//   tile      per warp and component, a ring of 2*CH chunks; a chunk is 32 rows (one per lane) x 16 time points
//             (128 bytes per row, 4 KB), laid out as the 3-D TMA box {16 doubles, 32 rows, CH chunks} with the
//             128-byte hardware swizzle, so a lane's 16-byte store lands in bank quad (k ^ lane%8).
//   body      16 steps (four Philox blocks) advance in registers — the generator software-pipelined one block ahead, so that
//             the step's FP64 chain, the next block's inversion and the Philox rounds of the block after share one basic
//             block (SDEGPU_TRAJ_PIPE) — then every component's 16 values go to its tile as 128-bit shared stores (64-bit
//             ones for odd row phases); after CH bodies ONE elected lane issues one `cp.async.bulk.tensor.3d` per
//             component (SASS UTMASTG): CH*128 bytes per row, 32 rows.  The other half of the ring is being read
//             by the previous store meanwhile (cp.async.bulk.wait_group.read 0 before it is overwritten).
//   alignment rows start at arbitrary element offsets (nt is usually odd) but TMA wants 16-byte and the memory system
//             32-byte (sector) alignment: rows are split into S = lcm(4/gcd(nt,4), nx) classes by (row mod S); within
//             a class all rows start at the same offset mod 4, so one tensor map per class, based at the class's
//             first sector boundary h in 1..4 elements into the row, covers every row of the class with a stride of
//             S*nt*8 bytes.  Lanes take every (S/nx)-th path so that a warp's rows of one component are consecutive
//             rows of one class.  The h head elements, and the < 16+3 elements behind the last whole chunk, are written
//             element by element (two partial sectors per row).
// The step is the folded FAST step of k_terminal (same coefficient rows, same generator), so X[nt,] of a trajectory run
// equals XT of a terminal run bit for bit.
#ifndef SDEGPU_TRAJ_CH
#define SDEGPU_TRAJ_CH 2
#endif
#ifndef SDEGPU_TRAJ_REPL
#define SDEGPU_TRAJ_REPL 8
#endif
#ifndef SDEGPU_TRAJ_MAX_THREADS
#define SDEGPU_TRAJ_MAX_THREADS 256
#endif
#ifndef SDEGPU_TRAJ_PIPE
#define SDEGPU_TRAJ_PIPE 1                     // generator software-pipelined one Philox block ahead of the step chain
#endif
constexpr int TRAJ_REPL = SDEGPU_TRAJ_REPL;
// Per state dimension: a scalar model owns one row per path, so 12 warps (384 rows in flight) with single-chunk stores
// (128 B per row) still sit on the memory system's good side and give the generator and the step — which co-limit these
// models — three warps per scheduler (tuned, profiles/r2_tune_traj_scalar.txt: OU 4.8e11 -> 5.3e11, CIR/heun 2.7e11 ->
// 3.4e11 path-steps/s); two-component models keep 4 warps x 2 rows x 2-chunk stores (256 rows).
#ifndef SDEGPU_TRAJ_CH_NX1
#define SDEGPU_TRAJ_CH_NX1 1
#endif
#ifndef SDEGPU_TRAJ_MAX_THREADS_NX1
#define SDEGPU_TRAJ_MAX_THREADS_NX1 384
#endif
template <int NX> struct TrajCfg {
    static constexpr int CH = NX == 1 ? SDEGPU_TRAJ_CH_NX1 : SDEGPU_TRAJ_CH;          // chunks per bulk store
    static constexpr int NSLOT = 2 * CH;                                              // chunk slots in a component's ring
    static constexpr int MAX_THREADS = NX == 1 ? SDEGPU_TRAJ_MAX_THREADS_NX1 : SDEGPU_TRAJ_MAX_THREADS;
};
constexpr int TRAJ_MAX_CLASSES = 4;                              // lcm(4, nx) for nx <= 2

struct __align__(64) TrajMaps {
    unsigned char m[TRAJ_MAX_CLASSES][128];                      // CUtensorMap, one per row class
};
struct TrajGeom {
    int sp;                                                      // lanes take every sp-th path; classes = sp*nx
    int nchunks;                                                 // whole 16-point chunks every class has behind its head
    int head[TRAJ_MAX_CLASSES];                                  // elements in front of the class's first sector boundary, 1..4
};

__device__ __forceinline__ void st_sector(double *p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_f64x2(uint32_t addr, double v0, double v1)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v0), "d"(v1) : "memory");
}
// swizzled byte offset of time slot s (0..15) of the lane's row inside a chunk; l7 = (lane & 7) << 4
__device__ __forceinline__ uint32_t traj_slot(uint32_t lane_row, uint32_t l7, int s)
{
    return lane_row + ((((uint32_t)s >> 1) << 4) ^ l7) + (((uint32_t)s & 1u) << 3);
}

// The 16 values of body m (times 16m+1 .. 16m+16) of one component into the ring: slot u + D of chunk m, or slot
// 16 + u + D of chunk m-1 when u + D < 0 (D = 1 - head in 0,-1,-2,-3); in body 0 those are head elements of the row.
template <int D>
__device__ __forceinline__ void traj_put(const double (&v)[16], uint32_t cur, uint32_t prev, uint32_t lane_row, uint32_t l7,
                                         bool first, bool valid, double *grow)
{
#pragma unroll
    for (int u = 0; u < 16; ++u) {
        const int vs = u + D;                                    // slot in chunk m; negative: slot 16 + vs of chunk m-1
        if ((D & 1) != 0) {
            // odd shift: a 128-bit store would pair (v[1], v[2]), ... — registers that are not an aligned quad, which costs two
            // moves per value (36 per component and body); sixteen 64-bit stores move the same bytes in the same wavefronts
            if (vs < 0) {
                if (first) {
                    if (valid) grow[1 + u] = v[u];
                } else sts_f64(prev + traj_slot(lane_row, l7, 16 + vs), v[u]);
            } else sts_f64(cur + traj_slot(lane_row, l7, vs), v[u]);
            continue;
        }
        const bool even = (vs & 1) == 0, pair = even && u + 1 < 16;
        if (!even && u != 0) continue;                           // second half of the pair stored at u - 1
        if (vs < 0) {
            if (first) {
                if (valid) {
                    grow[1 + u] = v[u];
                    if (pair) grow[2 + u] = v[u + 1];
                }
            } else if (pair) sts_f64x2(prev + traj_slot(lane_row, l7, 16 + vs), v[u], v[u + 1]);
            else sts_f64(prev + traj_slot(lane_row, l7, 16 + vs), v[u]);
        } else if (pair) sts_f64x2(cur + traj_slot(lane_row, l7, vs), v[u], v[u + 1]);
        else sts_f64(cur + traj_slot(lane_row, l7, vs), v[u]);
    }
}

template <class M, int SCHEME, int PROJ, bool UNIFORM>
__global__ void __launch_bounds__(TrajCfg<M::NX>::MAX_THREADS, 1)
k_traj(const __grid_constant__ RunArgs a, const __grid_constant__ TrajMaps maps, const __grid_constant__ TrajGeom geo)
{
    constexpr int NX = M::NX, CH = TrajCfg<NX>::CH, NSLOT = TrajCfg<NX>::NSLOT;
    constexpr uint32_t CHUNK = 4096u, RING = NSLOT * CHUNK;
    using FS = FastStep<M, SCHEME, PROJ>;
    using TL = Icdf32Layout<TRAJ_REPL>;
    extern __shared__ __align__(1024) unsigned char sh_raw[];    // [rings: warps x NX x NSLOT chunks][quantile table][histogram][scratch]
    // the warp index through a shuffle from lane 0: the compiler then knows it is warp-uniform, so the task bookkeeping (task and
    // ring offsets, TMA coordinates) lives in uniform registers and the bulk stores need no register-to-uniform conversion loops
    const int nthreads = blockDim.x, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31, nwarps = nthreads >> 5;
    double *sh_tab = reinterpret_cast<double *>(sh_raw + (size_t)nwarps * NX * RING);
    unsigned int *sh_hist = reinterpret_cast<unsigned int *>(sh_tab + TL::DOUBLES);
    const int nhist = a.hist ? ((a.hist_nbins + 3 + 3) & ~3) : 0;
    icdf32_stage<TRAJ_REPL>(sh_tab, a.icdf32, UNIFORM ? a.ucoef[0] : 1.0, threadIdx.x, nthreads);
    for (int i = threadIdx.x; i < nhist; i += nthreads) sh_hist[i] = 0u;
    __syncthreads();
    NormalGen<TRAJ_REPL> gen;
    gen.init(sh_tab, a.icdf, UNIFORM ? a.ucoef[0] : 1.0);
    double acc[NX][5];
#pragma unroll
    for (int j = 0; j < NX; ++j)
#pragma unroll
        for (int m = 0; m < 5; ++m) acc[j][m] = 0.0;

    const int nt = a.nt, nsteps = nt - 1, sp = geo.sp, nbody = geo.nchunks;
    const uint32_t ring0 = smem_addr_u32(sh_raw) + (uint32_t)warp * NX * RING;
    const uint32_t lane_row = (uint32_t)lane << 7, l7 = ((uint32_t)lane & 7u) << 4;
    const CoefUniform cu{a.ucoef};
    const uint64_t ntasks = (a.npaths + 31) >> 5;                // task T: paths {32*sp*(T/sp) + T%sp + sp*L, L = 0..31}
    for (uint64_t T = (uint64_t)blockIdx.x * nwarps + warp; T < (ntasks + sp - 1) / sp * sp; T += (uint64_t)gridDim.x * nwarps) {
        const uint64_t grp = T / (unsigned)sp;
        const int w = (int)(T - grp * (unsigned)sp);
        const uint64_t pbase = grp * 32u * (unsigned)sp + (unsigned)w;
        if (pbase >= a.npaths) continue;
        const uint64_t p = pbase + (uint64_t)sp * lane;
        const bool valid = p < a.npaths;
        const uint64_t pl = valid ? p : pbase, gp = a.path_offset + pl;
        double x[NX];
        load_x0<NX>(a, pl, x);
        double *grow[NX];
        int hd[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            grow[j] = a.X + (pl * NX + j) * (uint64_t)nt;
            hd[j] = geo.head[w * NX + j];
            if (valid) grow[j][0] = x[j];                        // X[1,] <- x0 (:441)
        }
#if SDEGPU_TRAJ_PIPE
        // Software pipeline over Philox blocks (four steps each).  While block g integrates — a dependent FP64 chain —
        // the normals of block g+1 are evaluated from words generated during block g-1 and the words of block g+2 are
        // generated: three independent instruction streams in ONE basic block, which is what a kernel that runs one
        // warp per scheduler needs (before, the 16 inversions, the 16 steps and the stores ran one after the other and
        // a warp-step took ~130 cycles for ~45 instructions).  A tail word (1 block in ~1000) is evaluated on the
        // clamped segment first and replaced behind the block, so the common path has no branch in front of it.
        double zc[4];                                            // scale * normals of block g
        uint32_t wn[4];                                          // words of block g+1
        PhiloxPath pp;                                           // the path's share of the Philox rounds, computed once
        pp.init(a.seed, gp, 0u);
        {
            uint32_t w0[4];
            pp.block(a.seed, 0u, w0[0], w0[1], w0[2], w0[3]);
            gen.quad(a.seed, gp, 0u, 0u, w0, zc);
            pp.block(a.seed, 1u, wn[0], wn[1], wn[2], wn[3]);
        }
        int issued = 0;
        uint32_t cur = 0u, prev = (NSLOT - 1) * CHUNK;           // ring offsets of chunks m and m-1
        // two bodies per iteration for the Euler scheme: the ring bookkeeping and the `m % CH` tests resolve per copy (C3 0.770 -> 0.789,
        // OU trajectories 0.71 -> 0.74); Heun's longer body measured 5 % slower unrolled (CIR/heun 0.485 -> 0.461) and stays as it is
#pragma unroll (SCHEME == SDEGPU_EULER ? 2 : 1)
        for (int m = 0; m < nbody; ++m) {
            double v[NX][16];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t g = (uint32_t)(4 * m + b);
                double zn[4];
                uint32_t wk[4];
                const bool tail = gen.bulk_quad_clamped(wn, zn);
#pragma unroll
                for (int k = 0; k < 4; ++k) wk[k] = wn[k];
                pp.block(a.seed, g + 2u, wn[0], wn[1], wn[2], wn[3]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (UNIFORM) FS::apply(a.P, x, cu, zc[k]);
                    else {
                        CoefRow row;
                        row.load(a.coef + (size_t)(16 * m + 4 * b + k) * FAST_NCOEF);
                        FS::apply(a.P, x, row, row[0] * zc[k]);
                    }
#pragma unroll
                    for (int j = 0; j < NX; ++j) v[j][4 * b + k] = x[j];
                }
                if (tail) gen.fix_tail(a.seed, gp, g + 1u, 0u, wk, zn);
#pragma unroll
                for (int k = 0; k < 4; ++k) zc[k] = zn[k];
            }
#else
        uint32_t r[4][4];
#pragma unroll
        for (int b = 0; b < 4; ++b) philox_block(a.seed, gp, (uint32_t)b, 0u, r[b][0], r[b][1], r[b][2], r[b][3]);
        int issued = 0;
        uint32_t cur = 0u, prev = (NSLOT - 1) * CHUNK;           // ring offsets of chunks m and m-1
        for (int m = 0; m < nbody; ++m) {
            double v[NX][16], z[16];
            gen.template multi<4>(a.seed, gp, (uint32_t)(4 * m), 0u, r, z);
#pragma unroll
            for (int b = 0; b < 4; ++b) {                        // words of the next body: integer work under the FP64 chain
                philox_block(a.seed, gp, (uint32_t)(4 * (m + 1) + b), 0u, r[b][0], r[b][1], r[b][2], r[b][3]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (UNIFORM) FS::apply(a.P, x, cu, z[4 * b + k]);
                    else {
                        CoefRow row;
                        row.load(a.coef + (size_t)(16 * m + 4 * b + k) * FAST_NCOEF);
                        FS::apply(a.P, x, row, row[0] * z[4 * b + k]);
                    }
#pragma unroll
                    for (int j = 0; j < NX; ++j) v[j][4 * b + k] = x[j];
                }
            }
#endif
            if (m % CH == 0) {                                   // entering the ring half the store before last read from
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                const uint32_t rj = ring0 + (uint32_t)j * RING;
                // warp-uniform; two compare-and-branch levels instead of a switch: ptxas turns the switch into a jump table
                // (LDC + SHF + BRX), which costs a kernel that runs one warp per scheduler ~3 % of its time per component
                if (hd[j] <= 2) {
                    if (hd[j] == 1) traj_put<0>(v[j], rj + cur, rj + prev, lane_row, l7, m == 0, valid, grow[j]);
                    else traj_put<-1>(v[j], rj + cur, rj + prev, lane_row, l7, m == 0, valid, grow[j]);
                } else {
                    if (hd[j] == 3) traj_put<-2>(v[j], rj + cur, rj + prev, lane_row, l7, m == 0, valid, grow[j]);
                    else traj_put<-3>(v[j], rj + cur, rj + prev, lane_row, l7, m == 0, valid, grow[j]);
                }
            }
            if (m > 0 && m % CH == 0) {                          // chunks m-CH .. m-1 are complete for every component
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < NX; ++j)
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                                     ::"l"(&maps.m[w * NX + j]), "r"(0), "r"((int)(grp * 32u)), "r"(m - CH),
                                       "r"(ring0 + (uint32_t)j * RING + (uint32_t)((m - CH) % NSLOT) * CHUNK) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                issued = m;
            }
            prev = cur;
            cur = cur + CHUNK == RING ? 0u : cur + CHUNK;
        }
        // the steps behind the last whole body (< 16 + 3 of them): slots of chunk nbody-1 still go to the ring, the
        // rest of the row is written element by element
        for (int i = 16 * nbody; i < nsteps; i += 4) {
            double z[4];
#if SDEGPU_TRAJ_PIPE
            if (i == 16 * nbody) {                               // the pipeline's current block
#pragma unroll
                for (int k = 0; k < 4; ++k) z[k] = zc[k];
            } else {
                uint32_t w[4];
                philox_block(a.seed, gp, (uint32_t)(i >> 2), 0u, w[0], w[1], w[2], w[3]);
                gen.quad(a.seed, gp, (uint32_t)(i >> 2), 0u, w, z);
            }
#else
            if (i != 16 * nbody) philox_block(a.seed, gp, (uint32_t)(i >> 2), 0u, r[0][0], r[0][1], r[0][2], r[0][3]);
            gen.quad(a.seed, gp, (uint32_t)(i >> 2), 0u, r[0], z);
#endif
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (i + k < nsteps) {
                    if (UNIFORM) FS::apply(a.P, x, cu, z[k]);
                    else {
                        CoefRow row;
                        row.load(a.coef + (size_t)(i + k) * FAST_NCOEF);
                        FS::apply(a.P, x, row, row[0] * z[k]);
                    }
                    const int t = i + k + 1;
#pragma unroll
                    for (int j = 0; j < NX; ++j) {
                        const int s = t - hd[j];
                        if (s >= 0 && s < 16 * nbody) sts_f64(ring0 + (uint32_t)j * RING + prev + traj_slot(lane_row, l7, s & 15), x[j]);
                        else if (valid) grow[j][t] = x[j];
                    }
                }
        }
        if (issued < nbody) {                                    // the last 1..CH chunks (the box is clipped at nbody chunks)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < NX; ++j)
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                                 ::"l"(&maps.m[w * NX + j]), "r"(0), "r"((int)(grp * 32u)), "r"(issued),
                                   "r"(ring0 + (uint32_t)j * RING + (uint32_t)(issued % NSLOT) * CHUNK) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        if (valid) {
            if (a.XT) {
#pragma unroll
                for (int j = 0; j < NX; ++j) a.XT[p * NX + j] = x[j];
            }
            accumulate_terminal<NX>(a, x, acc, sh_hist);
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    flush_block_stats_dyn<NX>(a, acc, sh_hist, reinterpret_cast<double *>(sh_hist + nhist));
}

// ---- supplied-B kernel: Brownian paths in, trajectory and/or terminal state out ----------------
// Same streaming pattern as k_traj, on both sides of the step: each thread walks its own row of B
// (B[p*nt + i], time fastest) with aligned 256-bit loads — one 32-byte sector = four increments — and
// writes X through the sector writer.  No shared memory, no barrier: rows are private to a thread.
// dB_i = B[i+1] - B[i] is formed from the same doubles the reference subtracts (:437); with A = Strict
// every operation is individually rounded in the reference's order, so X is bit-identical to it.
#ifndef SDEGPU_STREAM_THREADS
#define SDEGPU_STREAM_THREADS 256
#endif
#ifndef SDEGPU_STREAM_BLOCKS_PER_SM
#define SDEGPU_STREAM_BLOCKS_PER_SM 0          // 0: as many as fit
#endif
#ifndef SDEGPU_STREAM_PREFETCH
#define SDEGPU_STREAM_PREFETCH 2               // sectors of B in flight per thread
#endif
constexpr int STREAM_THREADS = SDEGPU_STREAM_THREADS;
constexpr int STREAM_PREFETCH = SDEGPU_STREAM_PREFETCH;

__device__ __forceinline__ void ld_sector(const double *p, double (&s)[4])
{
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(s[0]), "=d"(s[1]), "=d"(s[2]), "=d"(s[3]) : "l"(p));
}

// Writes the sector that completes in a 4-value group: g = address of the group's first value, d = its
// position in the 32-byte grid (4 = element-wise only), pv = previous group's values.
__device__ __forceinline__ void store_group(double *g, int d, bool before_row, const double (&v)[4], double (&pv)[4])
{
    if (d == 4 || before_row) {                                  // element-wise: the sector would start before the row
        const int n = d == 4 ? 4 : 4 - d;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < n) g[k] = v[k];
    } else if (d == 0) st_sector(g, v[0], v[1], v[2], v[3]);
    else if (d == 1) st_sector(g - 1, pv[3], v[0], v[1], v[2]);
    else if (d == 2) st_sector(g - 2, pv[2], pv[3], v[0], v[1]);
    else st_sector(g - 3, pv[1], pv[2], pv[3], v[0]);
#pragma unroll
    for (int k = 0; k < 4; ++k) pv[k] = v[k];
}

template <class M, int SCHEME, int PROJ, class A>
__global__ void __launch_bounds__(STREAM_THREADS) k_stream(const RunArgs a, const int vec_in, const int vec_out)
{
    constexpr int NX = M::NX;
    extern __shared__ __align__(16) unsigned int sh_hist[];
    if (a.hist)
        for (int i = threadIdx.x; i < a.hist_nbins + 3; i += STREAM_THREADS) sh_hist[i] = 0u;
    __syncthreads();
    double acc[NX][5];
#pragma unroll
    for (int j = 0; j < NX; ++j)
#pragma unroll
        for (int m = 0; m < 5; ++m) acc[j][m] = 0.0;

    const int nsteps = a.nt - 1;
    const uint64_t stride = (uint64_t)gridDim.x * STREAM_THREADS;
    for (uint64_t p = (uint64_t)blockIdx.x * STREAM_THREADS + threadIdx.x; p < a.npaths; p += stride) {
        double x[NX], v[NX][4], pv[NX][4];
        const double *brow = a.B + p * (uint64_t)a.nt;           // B[i] of this path
        double *xrow[NX];                                        // &X[i = 1] of each component
        load_x0<NX>(a, p, x);
        // steps taken one by one until B[s+1] starts an aligned sector (all of them if B is unaligned)
        int sa = vec_in ? (int)((4 - ((p * (uint64_t)a.nt + 1) & 3)) & 3) : nsteps;
        if (sa > nsteps) sa = nsteps;
        const int ngroups = (nsteps - sa) >> 2;
        int d[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            const uint64_t e0 = (p * NX + j) * (uint64_t)a.nt;
            xrow[j] = a.X ? a.X + e0 + 1 : nullptr;
            d[j] = vec_out ? (int)((e0 + 1 + sa) & 3) : 4;       // grid position of the first grouped value
            if (a.X) a.X[e0] = x[j];                             // X[1,] <- x0 (:441)
#pragma unroll
            for (int k = 0; k < 4; ++k) pv[j][k] = x[j];
        }
        double prev = __ldg(brow);
        int s = 0;
        for (; s < sa; ++s) {                                    // head: at most three steps
            const double cur = __ldg(brow + s + 1);
            const double2 dt = __ldg(&a.dtsq[s]);
            sde_step<M, SCHEME, PROJ, A>(a.P, x, dt.x, A::sub(cur, prev));
            prev = cur;
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                if (a.X) xrow[j][s] = x[j];
#pragma unroll
                for (int k = 0; k < 3; ++k) pv[j][k] = pv[j][k + 1];
                pv[j][3] = x[j];
            }
        }
        // register pipeline STREAM_PREFETCH sectors deep: DRAM latency is ~1 us and a thread owns one row, so while
        // group m integrates, the sectors of the next groups are already in flight (loop unrolled by the depth)
        double bq[STREAM_PREFETCH][4];
#pragma unroll
        for (int k = 0; k < STREAM_PREFETCH; ++k)
            if (ngroups > k) ld_sector(brow + s + 1 + 4 * k, bq[k]);
        auto group = [&](double (&buf)[4], int m) {
            double b[4];
            double2 dts[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { b[k] = buf[k]; dts[k] = __ldg(&a.dtsq[s + k]); }
            if (m + STREAM_PREFETCH < ngroups) ld_sector(brow + s + 1 + 4 * STREAM_PREFETCH, buf);   // refill for group m+depth
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                sde_step<M, SCHEME, PROJ, A>(a.P, x, dts[k].x, A::sub(b[k], prev));      // dB = diff(B)  (:437)
                prev = b[k];
#pragma unroll
                for (int j = 0; j < NX; ++j) v[j][k] = x[j];
            }
            if (a.X) {
#pragma unroll
                for (int j = 0; j < NX; ++j) store_group(xrow[j] + s, d[j], m == 0 && d[j] > sa + 1, v[j], pv[j]);
            }
            s += 4;
        };
        int m = 0;
        for (; m + STREAM_PREFETCH <= ngroups; m += STREAM_PREFETCH) {
#pragma unroll
            for (int k = 0; k < STREAM_PREFETCH; ++k) group(bq[k], m + k);
        }
#pragma unroll
        for (int k = 0; k < STREAM_PREFETCH - 1; ++k)
            if (m + k < ngroups) group(bq[k], m + k);
        if (a.X && ngroups > 0) {                                // values that opened a sector no full group completes
#pragma unroll
            for (int j = 0; j < NX; ++j)
                if (d[j] >= 1 && d[j] <= 3) {
#pragma unroll
                    for (int k = 1; k < 4; ++k)
                        if (k >= 4 - d[j]) xrow[j][s - 4 + k] = pv[j][k];
                }
        }
        for (; s < nsteps; ++s) {                                // tail: at most three steps
            const double cur = __ldg(brow + s + 1);
            const double2 dt = __ldg(&a.dtsq[s]);
            sde_step<M, SCHEME, PROJ, A>(a.P, x, dt.x, A::sub(cur, prev));
            prev = cur;
            if (a.X) {
#pragma unroll
                for (int j = 0; j < NX; ++j) xrow[j][s] = x[j];
            }
        }
        if (a.XT) {
#pragma unroll
            for (int j = 0; j < NX; ++j) a.XT[p * NX + j] = x[j];
        }
        accumulate_terminal<NX>(a, x, acc, sh_hist);
    }
    flush_block_stats<NX, STREAM_THREADS>(a, acc, sh_hist);
}

// ---- running integrals along the path, nothing stored (SURVEY build-plan step 5) ------------------------
// For every path and component j the terminal entries of the reference's integral functions on that path:
//   QV    QuadraticVariation(X_j)[nt]      = sum (X_{i+1}-X_i)^2                    R/sdetools.R:134-137
//   ito   itointegral(X_j, B)[nt]          = sum X_i dB_i                           :155-158
//   strat stochint(X_j, B, "c")[nt]        = 0.5 (sum X_i dB_i + sum X_{i+1} dB_i)   :105-113
//   time  itointegral(X_j, times)[nt]      = sum X_i dt_i
// each sum in cumsum's 80-bit accumulator (float80.cuh) and each term formed with the same individually
// rounded operations k_integral uses, so with a supplied B they are bit-identical to running the
// integral kernels on the stored trajectory.  Without B, dB_i is the generator's increment sqrt(dt_i) Z_i.
#ifndef SDEGPU_FUNC_THREADS
#define SDEGPU_FUNC_THREADS 128
#endif
constexpr int FUNC_THREADS = SDEGPU_FUNC_THREADS;
constexpr int FUNC_PER_COMP = 4;

template <class M, int SCHEME, int PROJ, class A>
__global__ void __launch_bounds__(FUNC_THREADS) k_functionals(const RunArgs a)
{
    constexpr int NX = M::NX;
    extern __shared__ __align__(16) double sh_tab[];             // [quantile table (generator source only)][histogram]
    unsigned int *sh_hist = reinterpret_cast<unsigned int *>(sh_tab + (a.B ? 0 : ICDF_TABLE_DOUBLES));
    if (!a.B) icdf_stage(sh_tab, a.icdf, threadIdx.x, FUNC_THREADS);
    if (a.hist)
        for (int i = threadIdx.x; i < a.hist_nbins + 3; i += FUNC_THREADS) sh_hist[i] = 0u;
    __syncthreads();
    double acc[NX][5];
#pragma unroll
    for (int j = 0; j < NX; ++j)
#pragma unroll
        for (int m = 0; m < 5; ++m) acc[j][m] = 0.0;
    const int nsteps = a.nt - 1;
    const bool tdep = a.forcing || a.gscale;
    const uint64_t stride = (uint64_t)gridDim.x * FUNC_THREADS;
    for (uint64_t p = (uint64_t)blockIdx.x * FUNC_THREADS + threadIdx.x; p < a.npaths; p += stride) {
        double x[NX];
        load_x0<NX>(a, p, x);
        CumsumAcc qv[NX], il[NX], ir[NX], it[NX];
        F80 big[NX][FUNC_PER_COMP];                              // integer fallback states (local memory, rarely touched)
        double out[NX][FUNC_PER_COMP];
        CumsumAcc mf[NX], mg[NX];
        F80 mbig[NX][2];
        double mout[NX][2];
#pragma unroll
        for (int j = 0; j < NX; ++j) {
#pragma unroll
            for (int k = 0; k < FUNC_PER_COMP; ++k) out[j][k] = 0.0;
            mout[j][0] = mout[j][1] = 0.0;
        }
        const double *brow = a.B ? a.B + p * (uint64_t)a.nt : nullptr;
        double prev = brow ? __ldg(brow) : 0.0;
        double z[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < nsteps; ++i) {
            const double2 dt = __ldg(&a.dtsq[i]);
            double dB;
            if (brow) {
                const double cur = __ldg(brow + i + 1);
                dB = A::sub(cur, prev);                          // dB = diff(B)  (:437)
                prev = cur;
            } else {
                if ((i & 3) == 0) normal_quad(a.seed, a.path_offset + p, (uint32_t)(i >> 2), 0u, sh_tab, z);
                const int s4 = i & 3;
                dB = dt.y * (s4 == 0 ? z[0] : (s4 == 1 ? z[1] : (s4 == 2 ? z[2] : z[3])));
            }
            double xo[NX];
#pragma unroll
            for (int j = 0; j < NX; ++j) xo[j] = x[j];
            TDep<NX> td;
            if (tdep) {
                load_tdep<NX>(a, i, td);
                sde_step<M, SCHEME, PROJ, A, true>(a.P, x, dt.x, dB, &td);
            } else
                sde_step<M, SCHEME, PROJ, A>(a.P, x, dt.x, dB);
            if (a.func) {
#pragma unroll
                for (int j = 0; j < NX; ++j) {
                    const double dx = __dsub_rn(x[j], xo[j]);
                    out[j][0] = qv[j].add(__dmul_rn(dx, dx), &big[j][0]);
                    out[j][1] = il[j].add(__dmul_rn(xo[j], dB), &big[j][1]);
                    out[j][2] = ir[j].add(__dmul_rn(x[j], dB), &big[j][2]);
                    out[j][3] = it[j].add(__dmul_rn(xo[j], dt.x), &big[j][3]);
                }
            }
            if (a.mfunc) {                                       // itointegral(f(X), times), itointegral(g(X), B): integrands at (t_i, X_i)
                double fo[NX], go[NX];
                M::template f<A>(a.P, xo, fo);
                M::template g<A>(a.P, xo, go);
#pragma unroll
                for (int j = 0; j < NX; ++j) {
                    if (tdep && td.has_c) fo[j] = A::add(fo[j], td.cX[j]);
                    if (tdep && td.has_s) go[j] = A::mul(td.sX, go[j]);
                    mout[j][0] = mf[j].add(__dmul_rn(fo[j], dt.x), &mbig[j][0]);
                    mout[j][1] = mg[j].add(__dmul_rn(go[j], dB), &mbig[j][1]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            if (a.func) {
                double *f = a.func + (p * NX + j) * FUNC_PER_COMP;
                f[0] = out[j][0];
                f[1] = out[j][1];
                f[2] = __dmul_rn(0.5, __dadd_rn(out[j][1], out[j][2]));
                f[3] = out[j][3];
            }
            if (a.mfunc) {
                a.mfunc[(p * NX + j) * 2] = mout[j][0];
                a.mfunc[(p * NX + j) * 2 + 1] = mout[j][1];
            }
            if (a.XT) a.XT[p * NX + j] = x[j];
        }
        accumulate_terminal<NX>(a, x, acc, sh_hist);
    }
    flush_block_stats<NX, FUNC_THREADS>(a, acc, sh_hist);
}

// component `comp` of the state without indexing the register array by a run-time value
template <int NX>
__device__ __forceinline__ double pick_component(const double (&x)[NX], int comp)
{
    double v = x[0];
#pragma unroll
    for (int j = 1; j < NX; ++j) v = comp == j ? x[j] : v;
    return v;
}

// ---- mortality and stopping (R/sdetools.R:397-416,455-469; vignettes/Killing.Rmd:152-190) ----------
// Per step, in the reference's order: the new state is tested against h (stop: S is not updated and
// reads 0); otherwise S <- S * exp(-r(X_i) * dt_i) with the OLD state, and the path is killed when
// S < Stau.  tau = times[i+1] of the break, or the last grid time.  Paths leave the loop individually.
// Arithmetic is STRICT for both increment sources; Stau[p] defaults to a Philox uniform per path
// (the reference draws one runif() per euler() call).
constexpr int KILLED_THREADS = 256;

template <int NX>
__device__ __forceinline__ double kill_rate(const RunArgs &a, const double (&x)[NX])
{
    const double v = pick_component<NX>(x, a.kill_comp);
    if (a.kill_kind == SDEGPU_KILL_CONST) return a.kill_a;
    if (a.kill_kind == SDEGPU_KILL_EXP) return __dmul_rn(a.kill_a, exp(__dmul_rn(a.kill_b, v)));
    if (a.kill_kind == SDEGPU_KILL_QUAD) { const double t = __dsub_rn(v, a.kill_b); return __dmul_rn(a.kill_a, __dmul_rn(t, t)); }
    return 0.0;
}

template <int NX>
__device__ __forceinline__ double stop_value(const RunArgs &a, const double (&x)[NX])
{
    const double v = pick_component<NX>(x, a.stop_comp);
    if (a.stop_kind == SDEGPU_STOP_BELOW) return __dsub_rn(v, a.stop_lo);
    if (a.stop_kind == SDEGPU_STOP_ABOVE) return __dsub_rn(a.stop_hi, v);
    if (a.stop_kind == SDEGPU_STOP_OUTSIDE) return __dmul_rn(__dsub_rn(v, a.stop_lo), __dsub_rn(a.stop_hi, v));
    return 1.0;
}

template <class M, int SCHEME, int PROJ>
__global__ void __launch_bounds__(KILLED_THREADS) k_killed(const RunArgs a)
{
    constexpr int NX = M::NX;
    extern __shared__ __align__(16) double sh_tab[];             // [quantile table (if no B)][histogram]
    unsigned int *sh_hist = reinterpret_cast<unsigned int *>(sh_tab + (a.B ? 0 : ICDF_TABLE_DOUBLES));
    if (!a.B) icdf_stage(sh_tab, a.icdf, threadIdx.x, KILLED_THREADS);
    if (a.hist)
        for (int i = threadIdx.x; i < a.hist_nbins + 3; i += KILLED_THREADS) sh_hist[i] = 0u;
    __syncthreads();
    double acc[NX][5];
#pragma unroll
    for (int j = 0; j < NX; ++j)
#pragma unroll
        for (int m = 0; m < 5; ++m) acc[j][m] = 0.0;

    const int nsteps = a.nt - 1;
    const bool tdep = a.forcing || a.gscale;
    const uint64_t stride = (uint64_t)gridDim.x * KILLED_THREADS;
    for (uint64_t p = (uint64_t)blockIdx.x * KILLED_THREADS + threadIdx.x; p < a.npaths; p += stride) {
        double x[NX], xo[NX];
        load_x0<NX>(a, p, x);
        const uint64_t gp = a.path_offset + p;
        double Stau;
        if (a.Stau) Stau = a.Stau[p];
        else {                                                   // runif(1): 52-bit uniform in [0,1)
            uint32_t r0, r1, r2, r3;
            philox_block(a.seed, gp, 0u, 0xC0000000u, r0, r1, r2, r3);
            Stau = __hiloint2double((int)(0x3FF00000u | (r1 >> 12)), (int)((r1 << 20) | (r0 >> 12))) - 1.0;
        }
        if (a.X) {
#pragma unroll
            for (int j = 0; j < NX; ++j) a.X[(p * NX + j) * (uint64_t)a.nt] = x[j];
        }
        const double *brow = a.B ? a.B + p * (uint64_t)a.nt : nullptr;
        double S = a.S0, Sout = a.S0, zq[4] = {0.0, 0.0, 0.0, 0.0};
        double *srow = a.Spath ? a.Spath + p * (uint64_t)a.nt : nullptr;
        if (srow) srow[0] = a.S0;                                // S[1] <- S0  (:444)
        int last = nsteps;                                       // index of the last valid row of X
        for (int i = 0; i < nsteps; ++i) {
            const double2 dt = __ldg(&a.dtsq[i]);
            double dB;
            if (brow) dB = __dsub_rn(__ldg(brow + i + 1), __ldg(brow + i));
            else {
                if ((i & 3) == 0) normal_quad(a.seed, gp, (uint32_t)(i >> 2), 0u, sh_tab, zq);
                const int k = i & 3;
                dB = __dmul_rn(dt.y, k == 0 ? zq[0] : (k == 1 ? zq[1] : (k == 2 ? zq[2] : zq[3])));
            }
#pragma unroll
            for (int j = 0; j < NX; ++j) xo[j] = x[j];
            if (tdep) {
                TDep<NX> td;
                load_tdep<NX>(a, i, td);
                sde_step<M, SCHEME, PROJ, Strict, true>(a.P, x, dt.x, dB, &td);
            } else
                sde_step<M, SCHEME, PROJ, Strict>(a.P, x, dt.x, dB);
            if (a.X) {
#pragma unroll
                for (int j = 0; j < NX; ++j) a.X[(p * NX + j) * (uint64_t)a.nt + i + 1] = x[j];
            }
            if (a.stop_kind != SDEGPU_STOP_NONE && !(stop_value<NX>(a, x) > 0.0)) {      // h(...) <= 0  (:455)
                last = i + 1; Sout = 0.0;
                if (srow) srow[i + 1] = 0.0;                     // S <- numeric(nt) is not written at the stop (:443,:455)
                break;
            }
            if (a.kill_kind != SDEGPU_KILL_NONE) {
                S = __dmul_rn(S, exp(__dmul_rn(-kill_rate<NX>(a, xo), dt.x)));           // :456, old state
                Sout = S;
                if (srow) srow[i + 1] = S;
                if (S < Stau) { last = i + 1; break; }                                    // :457
            } else if (srow) srow[i + 1] = S;
        }
        const double nan = __longlong_as_double(0x7FF8000000000000ll);
        if (a.X) {                                               // rows after the break stay NA in the reference (:439)
            for (int i = last + 1; i <= nsteps; ++i)
#pragma unroll
                for (int j = 0; j < NX; ++j) a.X[(p * NX + j) * (uint64_t)a.nt + i] = nan;
        }
        if (srow)                                                // the reference returns S[1:(i+1)]: nothing after tau
            for (int i = last + 1; i <= nsteps; ++i) srow[i] = nan;
        if (a.tau) a.tau[p] = a.times[last];                     // tau <- times[i+1]  (:467)
        if (a.S_tau) a.S_tau[p] = Sout;
        if (a.XT) {
#pragma unroll
            for (int j = 0; j < NX; ++j) a.XT[p * NX + j] = x[j];
        }
        accumulate_terminal<NX>(a, x, acc, sh_hist);
    }
    flush_block_stats<NX, KILLED_THREADS>(a, acc, sh_hist);
}

// ---- mortality and stopping without a trajectory: lane-level refill ----------------------------------
// Lifetimes differ from path to path, so with one path per thread a warp runs until its longest-lived lane
// ends (measured: 19-29 % of the lanes' steps useful).  Here every warp owns a contiguous chunk of paths and
// a lane whose path has ended takes the chunk's next path; new paths start only on warp iterations that are
// multiples of four, so all lanes stay in the same phase of the four-normals-per-Philox-block stream.
// No atomics: which lane runs which path depends only on the lifetimes, so results — including the order of
// every floating-point accumulation — are reproducible.  Per-path results are those of k_killed bit for bit.
template <class M, int SCHEME, int PROJ>
__global__ void __launch_bounds__(KILLED_THREADS) k_killed_refill(const RunArgs a)
{
    constexpr int NX = M::NX;
    extern __shared__ __align__(16) double sh_tab[];             // [quantile table (if no B)][histogram]
    unsigned int *sh_hist = reinterpret_cast<unsigned int *>(sh_tab + (a.B ? 0 : ICDF_TABLE_DOUBLES));
    if (!a.B) icdf_stage(sh_tab, a.icdf, threadIdx.x, KILLED_THREADS);
    if (a.hist)
        for (int i = threadIdx.x; i < a.hist_nbins + 3; i += KILLED_THREADS) sh_hist[i] = 0u;
    __syncthreads();
    double acc[NX][5];
#pragma unroll
    for (int j = 0; j < NX; ++j)
#pragma unroll
        for (int m = 0; m < 5; ++m) acc[j][m] = 0.0;

    const int nsteps = a.nt - 1, lane = threadIdx.x & 31;
    const bool tdep = a.forcing || a.gscale;
    const uint64_t nwarps = (uint64_t)gridDim.x * (KILLED_THREADS / 32);
    const uint64_t warp = (uint64_t)blockIdx.x * (KILLED_THREADS / 32) + (threadIdx.x >> 5);
    const uint64_t chunk = (a.npaths + nwarps - 1) / nwarps;
    uint64_t next = warp * chunk;                                // warp-uniform: first unclaimed path of the chunk
    const uint64_t end = next + chunk < a.npaths ? next + chunk : a.npaths;

    bool active = false;
    uint64_t p = 0;
    int i = 0;
    double x[NX], xo[NX], S = 0.0, Sout = 0.0, Stau = 0.0, zq[4] = {0.0, 0.0, 0.0, 0.0};
    const double *brow = nullptr;
    for (uint32_t k = 0;; ++k) {
        if ((k & 3u) == 0u) {
            const unsigned idle = __ballot_sync(0xffffffffu, !active);
            if (idle) {
                if (!active) {
                    const uint64_t cand = next < end ? next + (uint64_t)__popc(idle & ((1u << lane) - 1u)) : end;
                    if (cand < end) {
                        p = cand; active = true; i = 0;
                        load_x0<NX>(a, p, x);
                        if (a.Stau) Stau = a.Stau[p];
                        else {                                   // runif(1): 52-bit uniform in [0,1)
                            uint32_t r0, r1, r2, r3;
                            philox_block(a.seed, a.path_offset + p, 0u, 0xC0000000u, r0, r1, r2, r3);
                            Stau = __hiloint2double((int)(0x3FF00000u | (r1 >> 12)), (int)((r1 << 20) | (r0 >> 12))) - 1.0;
                        }
                        brow = a.B ? a.B + p * (uint64_t)a.nt : nullptr;
                        S = a.S0; Sout = a.S0;
                    }
                }
                if (next < end) { const uint64_t left = end - next, want = (uint64_t)__popc(idle); next += want < left ? want : left; }
                if (!__any_sync(0xffffffffu, active)) break;     // chunk exhausted and every lane done
            }
            if (active && !brow) normal_quad(a.seed, a.path_offset + p, (uint32_t)(i >> 2), 0u, sh_tab, zq);
        }
        if (!active) continue;
        int last = -1;
        if (i < nsteps) {
            const double2 dt = __ldg(&a.dtsq[i]);
            double dB;
            if (brow) dB = __dsub_rn(__ldg(brow + i + 1), __ldg(brow + i));
            else {
                const int q = i & 3;
                dB = __dmul_rn(dt.y, q == 0 ? zq[0] : (q == 1 ? zq[1] : (q == 2 ? zq[2] : zq[3])));
            }
#pragma unroll
            for (int j = 0; j < NX; ++j) xo[j] = x[j];
            if (tdep) {
                TDep<NX> td;
                load_tdep<NX>(a, i, td);
                sde_step<M, SCHEME, PROJ, Strict, true>(a.P, x, dt.x, dB, &td);
            } else
                sde_step<M, SCHEME, PROJ, Strict>(a.P, x, dt.x, dB);
            if (a.stop_kind != SDEGPU_STOP_NONE && !(stop_value<NX>(a, x) > 0.0)) {      // h(...) <= 0  (:455)
                last = i + 1; Sout = 0.0;
            } else if (a.kill_kind != SDEGPU_KILL_NONE) {
                S = __dmul_rn(S, exp(__dmul_rn(-kill_rate<NX>(a, xo), dt.x)));           // :456, old state
                Sout = S;
                if (S < Stau) last = i + 1;                                               // :457
            }
            ++i;
            if (last < 0 && i == nsteps) last = nsteps;
        } else last = nsteps;                                    // nt = 1 never happens (nt >= 2), kept for safety
        if (last >= 0) {
            if (a.tau) a.tau[p] = a.times[last];                 // tau <- times[i+1]  (:467)
            if (a.S_tau) a.S_tau[p] = Sout;
            if (a.XT) {
#pragma unroll
                for (int j = 0; j < NX; ++j) a.XT[p * NX + j] = x[j];
            }
            accumulate_terminal<NX>(a, x, acc, sh_hist);
            active = false;
        }
    }
    flush_block_stats<NX, KILLED_THREADS>(a, acc, sh_hist);
}

// ---- host-side launch signature shared by the per-model translation units ---
struct LaunchCfg {
    int scheme, proj, arith;
    bool src_b, want_x, killed, want_func;
    bool uniform;        // the time grid is uniform to rounding: one coefficient row as a kernel parameter
    bool general;        // time-dependent closures: the general kernel (k_killed without a rule), STRICT arithmetic
    int grid;            // 0: choose from occupancy; every launcher clamps to 32 blocks per SM, the capacity of the partials buffer
    int sm_count;
    cudaStream_t stream;
};
// returns cudaError_t as int; *grid_used gets the grid size actually launched
typedef int (*model_launcher)(const RunArgs &, const LaunchCfg &, int *grid_used);
int launch_model(int model, const RunArgs &, const LaunchCfg &, int *grid_used);

}  // namespace sdegpu
