// Stochastic integrals on stored paths, batched over columns:
//   stochint(f,g,rule)      R/sdetools.R:105-113   l = c(0,cumsum(f[-n]*dg)), r = c(0,cumsum(f[-1]*dg)), c = 0.5*(l+r)
//   itointegral(G,B)        R/sdetools.R:155-158   == stochint left rule
//   QuadraticVariation(X)   R/sdetools.R:134-137   c(0,cumsum(diff(X)^2))
//   CrossVariation(X,Y)     R/sdetools.R:184-187   c(0,cumsum(diff(X)*diff(Y)))
// plus rvBM (R/sdetools.R:16-22,40-46) whose B0+cumsum(dB) uses the same accumulator.
//
// A prefix sum with R's rounding is a strict recurrence per column, so the
// parallel axis is the column (= path): one thread per column, the 80-bit
// accumulator in registers as a pair of doubles (float80.cuh).
#include <cuda_runtime.h>
#include <stdint.h>

#include "float80.cuh"
#include "integrals.cuh"
#include "philox.cuh"

namespace sdegpu {

#ifndef SDEGPU_INT_THREADS
#define SDEGPU_INT_THREADS 256
#endif
#ifndef SDEGPU_INT_MIN_BLOCKS
#define SDEGPU_INT_MIN_BLOCKS 1
#endif
#ifndef SDEGPU_INT_MAX_BLOCKS_PER_SM
#define SDEGPU_INT_MAX_BLOCKS_PER_SM 0         // 0: as many as fit
#endif
#ifndef SDEGPU_INT_DEPTH
#define SDEGPU_INT_DEPTH 1                     // sectors per input stream in flight ahead of the one being summed
#endif
constexpr int INT_THREADS = SDEGPU_INT_THREADS;


__device__ __forceinline__ void ld_sector(const double *p, double (&s)[4])
{
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(s[0]), "=d"(s[1]), "=d"(s[2]), "=d"(s[3]) : "l"(p));
}
__device__ __forceinline__ void st_sector(double *p, const double (&s)[4])
{
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(s[0]), "d"(s[1]), "d"(s[2]), "d"(s[3]) : "memory");
}

// One thread per column; the column is streamed as aligned 32-byte sectors (256-bit loads/stores, four
// elements each) straight between HBM and registers: columns are private to a thread, so there is no
// shared memory and no barrier.  All arrays have the same shape, hence the same sector phase.
template <int MODE>              // 0 stochint, 1 cross-variation
__global__ void __launch_bounds__(INT_THREADS, MODE == 1 ? 3 : SDEGPU_INT_MIN_BLOCKS) k_integral(const IntegralArgs q, const int vec_ok)
{
    const uint64_t nsteps = q.n - 1;
    const bool need_l = MODE == 1 || q.l || q.c, need_r = MODE == 0 && (q.r || q.c);
    const uint64_t stride = (uint64_t)gridDim.x * INT_THREADS;
    for (uint64_t col = (uint64_t)blockIdx.x * INT_THREADS + threadIdx.x; col < q.ncols; col += stride) {
        const uint64_t base = col * q.n;
        const double *pa = q.a + base, *pb = q.b + base;
        CumsumAcc accl, accr;
        F80 bigl, bigr;                                          // integer fallback state (local memory, rarely touched)
        if (q.l) q.l[base] = 0.0;                                // c(0, ...)
        if (MODE == 0 && q.r) q.r[base] = 0.0;
        if (MODE == 0 && q.c) q.c[base] = 0.0;
        double a0 = __ldg(pa), b0 = __ldg(pb);
        uint64_t sa = vec_ok ? ((4 - ((base + 1) & 3)) & 3) : nsteps;
        if (sa > nsteps) sa = nsteps;
        const uint64_t ngroups = (nsteps - sa) >> 2;
        double ol, orr, oc;
        auto element = [&](double a1, double b1) {               // consumes (a0,b0)->(a1,b1), produces out[s+1]
            if (MODE == 0) {
                const double dg = __dsub_rn(b1, b0);
                ol = need_l ? accl.add(__dmul_rn(a0, dg), &bigl) : 0.0;
                orr = need_r ? accr.add(__dmul_rn(a1, dg), &bigr) : 0.0;
                oc = __dmul_rn(0.5, __dadd_rn(ol, orr));
            } else {
                ol = accl.add(__dmul_rn(__dsub_rn(a1, a0), __dsub_rn(b1, b0)), &bigl);
            }
            a0 = a1; b0 = b1;
        };
        uint64_t s = 0;
        for (; s < sa; ++s) {
            element(__ldg(pa + s + 1), __ldg(pb + s + 1));
            if (q.l) q.l[base + s + 1] = ol;
            if (MODE == 0 && q.r) q.r[base + s + 1] = orr;
            if (MODE == 0 && q.c) q.c[base + s + 1] = oc;
        }
        double na[4], nb[4];                                     // next sectors, in flight while this group is summed
#if SDEGPU_INT_DEPTH == 2
        double na2[4], nb2[4];                                   // and the ones after: two sectors per stream in flight
#endif
        if (ngroups > 0) {
            ld_sector(pa + s + 1, na);
            if (pb != pa) ld_sector(pb + s + 1, nb);
#if SDEGPU_INT_DEPTH == 2
            if (ngroups > 1) {
                ld_sector(pa + s + 5, na2);
                if (pb != pa) ld_sector(pb + s + 5, nb2);
            }
#endif
        }
        for (uint64_t m = 0; m < ngroups; ++m, s += 4) {
            double va[4], vb[4], vl[4], vr[4], vc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { va[k] = na[k]; vb[k] = pb != pa ? nb[k] : na[k]; }
#if SDEGPU_INT_DEPTH == 2
#pragma unroll
            for (int k = 0; k < 4; ++k) { na[k] = na2[k]; nb[k] = nb2[k]; }
            if (m + 2 < ngroups) {
                ld_sector(pa + s + 9, na2);
                if (pb != pa) ld_sector(pb + s + 9, nb2);
            }
#else
            if (m + 1 < ngroups) {
                ld_sector(pa + s + 5, na);
                if (pb != pa) ld_sector(pb + s + 5, nb);
            }
#endif
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                element(va[k], vb[k]);
                vl[k] = ol; vr[k] = orr; vc[k] = oc;
            }
            if (q.l) st_sector(q.l + base + s + 1, vl);
            if (MODE == 0 && q.r) st_sector(q.r + base + s + 1, vr);
            if (MODE == 0 && q.c) st_sector(q.c + base + s + 1, vc);
        }
        for (; s < nsteps; ++s) {
            element(__ldg(pa + s + 1), __ldg(pb + s + 1));
            if (q.l) q.l[base + s + 1] = ol;
            if (MODE == 0 && q.r) q.r[base + s + 1] = orr;
            if (MODE == 0 && q.c) q.c[base + s + 1] = oc;
        }
    }
}

int launch_integral(int mode, const IntegralArgs &q, int sm_count, cudaStream_t stream)
{
    if (q.n == 0 || q.ncols == 0) return 0;
    const uint64_t nblk = (q.ncols + INT_THREADS - 1) / INT_THREADS;
    auto k0 = k_integral<0>;
    auto k1 = k_integral<1>;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mode == 0 ? k0 : k1, INT_THREADS, 0);
    if (per_sm < 1) per_sm = 1;
    if (SDEGPU_INT_MAX_BLOCKS_PER_SM > 0 && per_sm > SDEGPU_INT_MAX_BLOCKS_PER_SM) per_sm = SDEGPU_INT_MAX_BLOCKS_PER_SM;
    uint64_t grid = (uint64_t)per_sm * sm_count;
    if (grid > nblk) grid = nblk;
    uintptr_t bits = reinterpret_cast<uintptr_t>(q.a) | reinterpret_cast<uintptr_t>(q.b) | reinterpret_cast<uintptr_t>(q.l) |
                     reinterpret_cast<uintptr_t>(q.r) | reinterpret_cast<uintptr_t>(q.c);
    const int vec_ok = (bits & 31) == 0;
    if (mode == 0) k0<<<(int)grid, INT_THREADS, 0, stream>>>(q, vec_ok);
    else k1<<<(int)grid, INT_THREADS, 0, stream>>>(q, vec_ok);
    return (int)cudaGetLastError();
}

// ---- rvBM ensembles ---------------------------------------------------------

// One thread per column, the column written as aligned 32-byte sectors straight from registers (the k_traj
// store pattern): columns are dealt to lanes with stride Q (Q*nt = 0 mod 4) so that the sector phase — and with
// it the position of the four-normal Philox blocks inside a sector — is warp-uniform and nothing diverges.
// (Round 1 staged tiles through shared memory for coalesced stores: 2.3x slower than this, the barriers and the
// transposition cost more than the scattered sectors.)
constexpr int BM_THREADS = 256;

__global__ void __launch_bounds__(BM_THREADS) k_rvbm(const BMArgs q, const int vec_ok, const int Q)
{
    extern __shared__ __align__(16) double sh_tab[];
    icdf_stage(sh_tab, q.icdf, threadIdx.x, BM_THREADS);
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * BM_THREADS, W = 32u * (unsigned)Q;
    const uint64_t Tmax = (q.ncols + W - 1) / W * W;
    for (uint64_t T = (uint64_t)blockIdx.x * BM_THREADS + threadIdx.x; T < Tmax; T += stride) {
        const uint64_t within = T % W, col = (T - within) + (within >> 5) + (uint64_t)Q * (within & 31);
        if (col >= q.ncols) continue;
        double *row = q.B + col * (uint64_t)q.nt;
        CumsumAcc acc;
        F80 big;
        double zq[4] = {0.0, 0.0, 0.0, 0.0};
        auto element = [&](int i) -> double {                    // B[i] of this column
            if ((i & 3) == 0) normal_quad(q.seed, q.path_offset + col, (uint32_t)(i >> 2), 0u, sh_tab, zq);
            const double2 d = __ldg(&q.dtsq[i]);
            const int s4 = i & 3;
            const double z = s4 == 0 ? zq[0] : (s4 == 1 ? zq[1] : (s4 == 2 ? zq[2] : zq[3]));
            // rnorm(mean=u*dt, sd=sigma*sqrt(dt)) = mean + sd*Z
            const double dB = __dadd_rn(__dmul_rn(q.u, d.x), __dmul_rn(__dmul_rn(q.sigma, d.y), z));
            return __dadd_rn(q.B0, acc.add(dB, &big));
        };
        int head = vec_ok ? (int)((4 - ((col * (uint64_t)q.nt) & 3)) & 3) : q.nt;   // elements before the first sector boundary
        if (head > q.nt) head = q.nt;
        int i = 0;
        for (; i < head; ++i) row[i] = element(i);
        for (; i + 3 < q.nt; i += 4) {
            double v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = element(i + k);
            st_sector(row + i, v);
        }
        for (; i < q.nt; ++i) row[i] = element(i);
    }
}

int launch_rvbm(const BMArgs &q, int sm_count, cudaStream_t stream)
{
    if (q.ncols == 0 || q.nt == 0) return 0;
    const size_t smem = ICDF_TABLE_BYTES;
    const uint64_t nblk = (q.ncols + BM_THREADS - 1) / BM_THREADS;
    int per_sm = 1;
    cudaFuncSetAttribute(k_rvbm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rvbm, BM_THREADS, smem);
    uint64_t grid = (uint64_t)(per_sm < 1 ? 1 : per_sm) * sm_count;
    if (grid > nblk) grid = nblk;
    const int vec_ok = (reinterpret_cast<uintptr_t>(q.B) & 31) == 0;
    const int Q = (q.nt % 4 == 0) ? 1 : ((q.nt % 2 == 0) ? 2 : 4);
    k_rvbm<<<(int)grid, BM_THREADS, smem, stream>>>(q, vec_ok, Q);
    return (int)cudaGetLastError();
}

// ---- path functionals of Brownian ensembles -----------------------------
// apply(rvBM(times, n), 2, max) and friends (vignettes/BrownianMotion.Rmd:101-144) without storing the paths:
// column c is regenerated exactly as k_rvbm builds it (same normals, same 80-bit cumsum, same B0 + ...), so
// every functional is bit-identical to the one computed from the stored rvBM matrix.  One thread per column.

__global__ void __launch_bounds__(INT_THREADS) k_bm_functionals(const BMFuncArgs q)
{
    extern __shared__ __align__(16) double sh_tab[];
    icdf_stage(sh_tab, q.bm.icdf, threadIdx.x, INT_THREADS);
    __syncthreads();
    const bool up = q.level >= q.bm.B0;
    const uint64_t stride = (uint64_t)gridDim.x * INT_THREADS;
    for (uint64_t col = (uint64_t)blockIdx.x * INT_THREADS + threadIdx.x; col < q.bm.ncols; col += stride) {
        CumsumAcc acc;
        F80 big;
        double zq[4] = {0.0, 0.0, 0.0, 0.0};
        double vmax = 0.0, vmin = 0.0;
        int imax = 0, imin = 0, ihit = -1;
        for (int i = 0; i < q.bm.nt; ++i) {
            if ((i & 3) == 0) normal_quad(q.bm.seed, q.bm.path_offset + col, (uint32_t)(i >> 2), 0u, sh_tab, zq);
            const double2 d = __ldg(&q.bm.dtsq[i]);
            const int s4 = i & 3;
            const double z = s4 == 0 ? zq[0] : (s4 == 1 ? zq[1] : (s4 == 2 ? zq[2] : zq[3]));
            const double dB = __dadd_rn(__dmul_rn(q.bm.u, d.x), __dmul_rn(__dmul_rn(q.bm.sigma, d.y), z));
            const double b = __dadd_rn(q.bm.B0, acc.add(dB, &big));
            if (i == 0 || b > vmax) { vmax = b; imax = i; }      // which.max / which.min: first occurrence
            if (i == 0 || b < vmin) { vmin = b; imin = i; }
            if (ihit < 0 && (up ? b >= q.level : b <= q.level)) ihit = i;
        }
        if (q.vmax) q.vmax[col] = vmax;
        if (q.vmin) q.vmin[col] = vmin;
        if (q.tmax) q.tmax[col] = __ldg(&q.times[imax]);
        if (q.tmin) q.tmin[col] = __ldg(&q.times[imin]);
        if (q.thit) q.thit[col] = ihit >= 0 ? __ldg(&q.times[ihit]) : __longlong_as_double(0x7FF8000000000000ll);
    }
}

int launch_bm_functionals(const BMFuncArgs &q, int sm_count, cudaStream_t stream)
{
    if (q.bm.ncols == 0 || q.bm.nt == 0) return 0;
    const size_t smem = ICDF_TABLE_BYTES;
    cudaFuncSetAttribute(k_bm_functionals, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bm_functionals, INT_THREADS, smem);
    uint64_t grid = (uint64_t)(per_sm < 1 ? 1 : per_sm) * sm_count;
    const uint64_t nblk = (q.bm.ncols + INT_THREADS - 1) / INT_THREADS;
    if (grid > nblk) grid = nblk;
    k_bm_functionals<<<(int)grid, INT_THREADS, smem, stream>>>(q);
    return (int)cudaGetLastError();
}

}  // namespace sdegpu
