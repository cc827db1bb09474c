// Exact transition-law samplers: the bias-free ensembles Euler/Heun are judged against.
//   rOU(n,x0,lambda,xi,sigma,t)              R/sdetools.R:784-791   rnorm(n, mean, sd) with ou.parameters :706-712
//   rCIR(n,x0,lambda,xi,gamma,t,Stratonovich) R/sdetools.R:695-702   rchisq(n, df, ncp)/2/c with cir.parameters :606-617
// and their recursive use X_{k+1} ~ law(. | X_k, dt) (vignettes/CoxIngersollRoss.Rmd:76-84).
//
// One thread per draw; all randomness is a pure function of (seed, global draw id, transition index):
//   OU   one normal from the same quad stream the integrators use (channel 0);
//   CIR  R's rchisq builds the non-central chi-square as a Poisson mixture of central ones
//        (nmath/rchisq.c -> rnchisq.c: r = rpois(ncp/2); rchisq(2r) + rgamma(df/2, 2)); the two gammas have the same
//        scale, so here it is ONE gamma: 2*Gamma(df/2 + N), N ~ Poisson(ncp/2).  Poisson: inversion by
//        sequential search below mean 10, Hoermann's transformed rejection (PTRS) above; Gamma: Marsaglia-Tsang
//        squeeze (shape < 1 boosted with U^(1/shape)).  Each rejection trial consumes its own Philox block
//        (channel = 0xA0000000 | trial for Poisson, 0xB0000000 | trial for gamma), so a draw never depends on
//        how many trials another draw needed.
// The draws are not R's (different generator), the laws are: tests compare with pOU/pCIR.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "laws.cuh"
#include "philox.cuh"

namespace sdegpu {

constexpr int LAW_THREADS = 256;
constexpr uint32_t CH_POISSON = 0xA0000000u, CH_GAMMA = 0xB0000000u, CH_BOOST = 0xB8000000u;

__device__ __forceinline__ double u52(uint32_t lo, uint32_t hi)   // uniform on (0,1): (k + 1/2) 2^-52
{
    const uint64_t m = (((uint64_t)hi << 32) | lo) >> 12;
    return ((double)m + 0.5) * 0x1p-52;
}

// N ~ Poisson(mu)
__device__ double poisson_draw(const PhiloxKey &key, uint64_t id, uint32_t step, double mu)
{
    if (!(mu > 0.0)) return 0.0;
    if (mu < 10.0) {                                             // inversion, one uniform
        uint32_t r0, r1, r2, r3;
        philox_block(key, id, step, CH_POISSON, r0, r1, r2, r3);
        const double u = u52(r0, r1);
        double p = exp(-mu), F = p, k = 0.0;
        while (u > F && k < 200.0) { k += 1.0; p *= mu / k; F += p; }
        return k;
    }
    // PTRS (W. Hoermann 1993, "The transformed rejection method for generating Poisson random variables")
    const double smu = sqrt(mu), b = 0.931 + 2.53 * smu, a = -0.059 + 0.02483 * b;
    const double inv_alpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0), lmu = log(mu);
    for (uint32_t trial = 0; trial < 256u; ++trial) {
        uint32_t r0, r1, r2, r3;
        philox_block(key, id, step, CH_POISSON | trial, r0, r1, r2, r3);
        const double U = u52(r0, r1) - 0.5, V = u52(r2, r3);
        const double us = 0.5 - fabs(U);
        const double k = floor((2.0 * a / us + b) * U + mu + 0.43);
        if (us >= 0.07 && V <= vr) return k;
        if (k < 0.0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(inv_alpha) - log(a / (us * us) + b) <= -mu + k * lmu - lgamma(k + 1.0)) return k;
    }
    return floor(mu);                                            // unreachable in practice (acceptance > 0.8 per trial)
}

// G ~ Gamma(shape, scale 1)
__device__ double gamma_draw(const PhiloxKey &key, uint64_t id, uint32_t step, double shape, const double *sh_tab)
{
    if (!(shape > 0.0)) return 0.0;
    const double a = shape < 1.0 ? shape + 1.0 : shape;
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    double g = d;
    for (uint32_t trial = 0; trial < 256u; ++trial) {            // Marsaglia & Tsang 2000
        uint32_t r0, r1, r2, r3;
        philox_block(key, id, step, CH_GAMMA | trial, r0, r1, r2, r3);
        const double x = icdf_normal<false>(r0, r1, sh_tab), u = u52(r2, r3);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        const double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2 || log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) { g = d * v; break; }
    }
    if (shape < 1.0) {
        uint32_t r0, r1, r2, r3;
        philox_block(key, id, step, CH_BOOST, r0, r1, r2, r3);
        g *= pow(u52(r0, r1), 1.0 / shape);
    }
    return g;
}

__global__ void __launch_bounds__(LAW_THREADS) k_rlaw(const LawArgs q)
{
    extern __shared__ __align__(16) double sh_tab[];
    icdf_stage(sh_tab, q.icdf, threadIdx.x, LAW_THREADS);
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * LAW_THREADS;
    for (uint64_t p = (uint64_t)blockIdx.x * LAW_THREADS + threadIdx.x; p < q.n; p += stride) {
        const uint64_t id = q.path_offset + p;
        double x = q.x0 ? q.x0[p * q.x0_stride] : q.x0s;
        double *row = q.path ? q.path + p * (uint64_t)(q.nsteps + 1) : nullptr;
        if (row) row[0] = x;
        double z[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < q.nsteps; ++i) {
            if (q.law == 0) {                                    // OU: mean + sd Z, mean = xi + (x - xi) e^{-lambda t}
                if ((i & 3) == 0) normal_quad(q.seed, id, (uint32_t)(i >> 2), 0u, sh_tab, z);
                const int s = i & 3;
                const double zi = s == 0 ? z[0] : (s == 1 ? z[1] : (s == 2 ? z[2] : z[3]));
                x = fma(q.sd, zi, fma(x - q.xi, q.decay, q.xi));
            } else {                                             // CIR: ncchisq(df, 2 c x e^{-lambda t}) / (2c)
                const double ncp = q.two_c * x * q.decay;
                const double N = poisson_draw(q.seed, id, (uint32_t)i, 0.5 * ncp);
                x = 2.0 * gamma_draw(q.seed, id, (uint32_t)i, 0.5 * q.df + N, sh_tab) / q.two_c;
            }
            if (row) row[i + 1] = x;
        }
        q.out[p] = x;
    }
}

// ---- evaluation of the laws: dOU/pOU (R/sdetools.R:745-765), dCIR/pCIR (:652-676) --------------------------
// The reference calls R's dnorm/pnorm and dchisq/pchisq(ncp); here the non-central chi-square is summed as the
// Poisson mixture it is, outwards from the heaviest Poisson term j0 = floor(ncp/2):
//   P(y; k, ncp) = sum_j Pois(j; ncp/2) P_gamma(k/2 + j, y/2),   f = sum_j Pois(j; ncp/2) g(k/2 + j)
// with one incomplete-gamma evaluation at j0 and the recurrences P(a+1,x) = P(a,x) - t(a), t(a) = x^a e^-x/Gamma(a+1),
// t(a+1) = t(a) x/(a+1), g(a) = t(a-1)/2.  Meant for goodness-of-fit on device-resident terminal states
// (probability integral transform u = F(X_T)) without a round trip through the host.
// log( x^a e^-x / Gamma(a) ).  For large a the three terms are ~a log a each and cancel to O(1): written with
// Stirling's series for lgamma the cancellation happens analytically (a log(x/a) - (x - a) + ...).
__device__ double log_lead(double a, double x)
{
    if (a < 30.0) return a * log(x) - x - lgamma(a);
    const double r = 1.0 / a, r2 = r * r;
    const double corr = r * (1.0 / 12.0 - r2 * (1.0 / 360.0 - r2 * (1.0 / 1260.0)));
    return a * log1p((x - a) / a) - (x - a) + 0.5 * log(a) - 0.91893853320467274178 - corr;
}

__device__ double gamma_p(double a, double x)                    // regularised lower incomplete gamma P(a, x)
{
    if (!(x > 0.0)) return 0.0;
    if (!(a > 0.0)) return 1.0;
    const double lead = exp(log_lead(a, x));                     // x^a e^-x / Gamma(a)
    if (x < a + 1.0) {                                           // series: lead * sum_n x^n / (a (a+1) ... (a+n))
        double term = 1.0 / a, sum = term, ap = a;
        for (int n = 0; n < 100000; ++n) {
            ap += 1.0;
            term *= x / ap;
            sum += term;
            if (term < sum * 1e-17) break;
        }
        return fmin(1.0, lead * sum);
    }
    // continued fraction for Q(a, x) (modified Lentz)
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int n = 1; n < 100000; ++n) {
        const double an = -n * (n - a);
        b += 2.0;
        d = an * d + b; if (fabs(d) < tiny) d = tiny;
        c = b + an / c; if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return fmax(0.0, 1.0 - lead * h);
}

// what = 0: density, 1: distribution function of chi'^2_k(ncp) at y
__device__ double ncx2_eval(int what, double y, double k, double ncp)
{
    if (!(y > 0.0)) return 0.0;
    const double x = 0.5 * y, half = 0.5 * ncp;
    if (!(half > 0.0)) {
        if (what) return gamma_p(0.5 * k, x);
        return 0.5 * exp((0.5 * k - 1.0) * log(x) - x - lgamma(0.5 * k));
    }
    const double j0 = floor(half), a0 = 0.5 * k + j0;
    const double w0 = j0 >= 1.0 ? exp(log_lead(j0, half)) / j0 : exp(-half);  // Pois(j0; ncp/2) = half^j0 e^-half / j0!
    const double t0 = exp(log_lead(a0, x)) / a0;                              // t(a0) = x^a0 e^-x / Gamma(a0 + 1)
    const double P0 = what ? gamma_p(a0, x) : 0.0;
    // value(a) summed: P(a) for the distribution, g(a) = t(a-1)/2 for the density; t(a0-1) = t(a0) a0/x
    double sum = w0 * (what ? P0 : 0.5 * t0 * a0 / x);
    {                                                            // upwards from j0
        double w = w0, t = t0, P = P0, a = a0;
        for (int n = 1; n < 200000; ++n) {
            w *= half / (j0 + n);
            double v;
            if (what) { P -= t; if (P < 0.0) P = 0.0; v = P; } else v = 0.5 * t;        // g(a+1) = t(a)/2
            a += 1.0;
            t *= x / (a + 0.0);                                  // t(a) from t(a-1): x/a ... a already incremented
            const double add = w * v;
            sum += add;
            if (add <= sum * 1e-17 && j0 + n > half) break;
        }
    }
    {                                                            // downwards to j = 0
        double w = w0, P = P0, a = a0, tm = t0 * a0 / x;         // tm = t(a-1)
        for (double j = j0; j >= 1.0; j -= 1.0) {
            w *= j / half;
            double v;
            if (what) { P += tm; if (P > 1.0) P = 1.0; v = P; }  // P(a-1) = P(a) + t(a-1)
            a -= 1.0;
            tm *= a / x;                                         // t(a-2) = t(a-1) (a-1)/x
            if (!what) v = 0.5 * tm;                             // g(a-1) = t(a-2)/2
            const double add = w * v;
            sum += add;
            if (add <= sum * 1e-17) break;
        }
    }
    return what ? fmin(1.0, sum) : sum;
}

__global__ void __launch_bounds__(LAW_THREADS) k_law_eval(const LawArgs q, const int what, const double *xs, double *out)
{
    const uint64_t stride = (uint64_t)gridDim.x * LAW_THREADS;
    for (uint64_t p = (uint64_t)blockIdx.x * LAW_THREADS + threadIdx.x; p < q.n; p += stride) {
        const double x0 = q.x0 ? q.x0[p * q.x0_stride] : q.x0s, x = xs[p];
        double r;
        if (q.law == 0) {
            const double z = (x - fma(x0 - q.xi, q.decay, q.xi)) / q.sd;
            r = what ? 0.5 * erfc(-z * 0.70710678118654752440) : exp(-0.5 * z * z) / (q.sd * 2.50662827463100050242);
        } else {
            const double v = ncx2_eval(what, q.two_c * x, q.df, q.two_c * x0 * q.decay);
            r = what ? v : q.two_c * v;                          // dCIR = 2c dchisq(2c x)  (:660)
        }
        out[p] = r;
    }
}

int launch_law_eval(const LawArgs &q, int what, const double *xs, double *out, int sm_count, cudaStream_t stream)
{
    if (q.n == 0) return 0;
    uint64_t grid = (uint64_t)sm_count * 4;
    const uint64_t nblk = (q.n + LAW_THREADS - 1) / LAW_THREADS;
    if (grid > nblk) grid = nblk;
    k_law_eval<<<(int)grid, LAW_THREADS, 0, stream>>>(q, what, xs, out);
    return (int)cudaGetLastError();
}

int launch_rlaw(const LawArgs &q, int sm_count, cudaStream_t stream)
{
    if (q.n == 0) return 0;
    const size_t smem = ICDF_TABLE_BYTES;
    cudaFuncSetAttribute(k_rlaw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rlaw, LAW_THREADS, smem);
    uint64_t grid = (uint64_t)(per_sm < 1 ? 1 : per_sm) * sm_count;
    const uint64_t nblk = (q.n + LAW_THREADS - 1) / LAW_THREADS;
    if (grid > nblk) grid = nblk;
    k_rlaw<<<(int)grid, LAW_THREADS, smem, stream>>>(q);
    return (int)cudaGetLastError();
}

}  // namespace sdegpu
