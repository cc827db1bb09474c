/* Plain-C restatement of the SDEtools sample-path hot path for ensembles.
 * TEST INFRASTRUCTURE ONLY: the product (sdetools_b200/, libsdegpu.so) never
 * links or calls this; tests, __graft_entry__.smoke() and bench.py's CPU legs do.
 *
 * PARITY: the reference (/root/reference/R/sdetools.R) has no numeric golden
 * vector for euler/heun and GNU R is not installed; this file is pinned bit for
 * bit against the numpy restatement (oracle/sde_oracle.py), and both against the
 * reference's own source text executed by oracle/minir — config C1 at full size
 * (1,000 paths x 1,000 steps) by SHA-256 (tests/test_minir.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off + SSE2 doubles = one correctly rounded IEEE op per R operator
 * (SURVEY.md App. A).  `long double` is x87 80-bit = R's cumsum accumulator.
 *
 * Reference lines followed:
 *   euler step      R/sdetools.R:446,437,450-453
 *   heun step       R/sdetools.R:316-323   (0.5*(gX+gY) %*% dB == 0.5*((gX+gY)%*%dB))
 *   itointegral     R/sdetools.R:155-158   stochint :105-113
 *   QuadraticVariation :134-137            CrossVariation :184-187
 *   rBM             R/sdetools.R:16-22
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { M_OU = 0, M_CIR = 1, M_VDP = 2, M_LOGISTIC = 3, M_GBM = 4, M_DOUBLEWELL = 5, M_LINEAR = 6, M_DWSQRT = 7, M_LOGLOGISTIC = 8 };
enum { P_ID = 0, P_ABS = 1, P_RELU = 2 };
enum { S_EULER = 0, S_HEUN = 1 };

/* ---- catalogue: drift f (nx) and diffusion g (nx x nB col-major) -------- */
static void drift(int model, const double *P, int nx, const double *x, double *f)
{
    switch (model) {
    case M_OU:
    case M_CIR: f[0] = P[0] * (P[1] - x[0]); break;                 /* lambda*(xi-x) */
    case M_VDP: f[0] = x[1]; f[1] = x[1] * (P[0] - x[0] * x[0]) - x[0]; break;
    case M_LOGISTIC: f[0] = x[0] * (1.0 - x[0]) - P[1] * (x[0] * x[0]); break;
    case M_GBM: f[0] = P[0] * x[0]; break;
    case M_DOUBLEWELL:
    case M_DWSQRT: f[0] = x[0] - x[0] * x[0] * x[0]; break;
    case M_LOGLOGISTIC: f[0] = (1.0 - exp(x[0])) - 0.5 * (P[0] * P[0]); break;
    case M_LINEAR: {                                                /* A %*% x, A = P[0..d*d) col-major */
        for (int j = 0; j < nx; ++j) f[j] = P[j] * x[0];
        for (int k = 1; k < nx; ++k)
            for (int j = 0; j < nx; ++j) f[j] = f[j] + P[(size_t)k * nx + j] * x[k];
    } break;
    }
}

static void diffusion(int model, const double *P, int nx, int nB, const double *x, double *g)
{
    switch (model) {
    case M_OU: g[0] = P[2]; break;
    case M_CIR: g[0] = P[2] * sqrt(fabs(x[0])); break;
    case M_VDP: g[0] = 0.0; g[1] = P[1]; break;
    case M_LOGISTIC: g[0] = P[0] * x[0]; break;
    case M_GBM: g[0] = P[1] * x[0]; break;
    case M_DOUBLEWELL: g[0] = P[0]; break;
    case M_DWSQRT: g[0] = sqrt(1.0 + x[0] * x[0]); break;
    case M_LOGLOGISTIC: g[0] = P[0]; break;
    case M_LINEAR: memcpy(g, P + (size_t)nx * nx, sizeof(double) * nx * nB); break;
    }
}

static inline double project(int proj, double v)
{
    if (proj == P_ABS) return fabs(v);
    if (proj == P_RELU) return v < 0.0 ? 0.0 : v;     /* pmax(x,0); NaN stays NaN */
    return v;
}

/* gX %*% dB: column-axpy order, no FMA; one multiply when nB == 1 */
static inline double matvec_row(const double *g, int nx, int nB, int j, const double *dB)
{
    double acc = g[j] * dB[0];
    for (int k = 1; k < nB; ++k) acc = acc + g[(size_t)k * nx + j] * dB[k];
    return acc;
}

/* ---- Philox4x32-10 + inversion (same spec as oracle/philox_ref.py) ------- */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

/* qnorm by Wichura's AS241 (PPND16), the algorithm behind R's qnorm and hence behind R's
 * default rnorm (inversion), R/sdetools.R:19.  Relative accuracy about 1e-16. */
static double as241_qnorm(double p)
{
    double q = p - 0.5, r, val;
    if (fabs(q) <= 0.425) {
        r = 0.180625 - q * q;
        return q * (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r
                         + 45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r
                       + 133.14166789178437745) * r + 3.387132872796366608)
                 / (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r
                        + 21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r
                      + 42.313330701600911252) * r + 1.0);
    }
    r = q < 0 ? p : 1.0 - p;
    r = sqrt(-log(r));
    if (r <= 5.0) {
        r -= 1.6;
        val = (((((((r * 7.7454501427834140764e-4 + 0.0227238449892691845833) * r + 0.24178072517745061177) * r
                   + 1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r
                + 4.6303378461565452959) * r + 1.42343711074968357734)
              / (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + 0.0151986665636164571966) * r
                     + 0.14810397642748007459) * r + 0.68976733498510000455) * r + 1.6763848301838038494) * r
                  + 2.05319162663775882187) * r + 1.0);
    } else {
        r -= 5.0;
        val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + 0.0012426609473880784386) * r
                   + 0.026532189526576123093) * r + 0.29656057182850489123) * r + 1.7848265399172913358) * r
                + 5.4637849111641143699) * r + 6.6579046435011037772)
              / (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r
                     + 7.868691311456132591e-4) * r + 0.0148753612908506148525) * r + 0.13692988092273580531) * r
                  + 0.59983220655588793769) * r + 1.0);
    }
    return q < 0.0 ? -val : val;
}

double oracle_qnorm(double p) { return as241_qnorm(p); }

/* one normal from 64 random bits: u from the top 52 bits, sign from bit 11 of lo, |z| = -qnorm(u/2) */
static double bits_to_normal(uint32_t lo, uint32_t hi)
{
    uint64_t m = (((uint64_t)hi << 32) | lo) >> 12;
    double u = (double)m * 0x1p-52;
    if (m == 0) u = 0x1p-52;
    double a = -as241_qnorm(0.5 * u);
    return (lo & 0x800u) ? -a : a;
}

/* one Brownian channel, "stream v2" (sdetools_b200/csrc/philox.cuh): four normals per block from 32-bit words.
 * w -> odd integer s = (w|1) - 2^31, sign(z) = sign(s), v = |s|, u = v 2^-31; draws beyond 2^-12 in the tail
 * (v < 2^19) refine their cell with the top 22 bits of the extension block's word (channel | 0x40000000):
 * u = ((v-1) 2^21 + e22 + 1) 2^-52, so the tail keeps 52-bit resolution */
static void normal_quad(uint64_t seed, uint64_t path, uint32_t block, uint32_t channel, double z[4])
{
    uint32_t c[4] = { (uint32_t)path, (uint32_t)(path >> 32), block, channel };
    uint32_t e[4] = { (uint32_t)path, (uint32_t)(path >> 32), block, channel | 0x40000000u };
    int have_ext = 0;
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    for (int s = 0; s < 4; ++s) {
        const int64_t si = (int64_t)(c[s] | 1u) - ((int64_t)1 << 31);
        const uint64_t v = (uint64_t)(si < 0 ? -si : si);
        uint64_t m = v << 21;
        if (v < (1u << 19)) {
            if (!have_ext) { philox4x32_10(e, (uint32_t)seed, (uint32_t)(seed >> 32)); have_ext = 1; }
            m = ((v - 1) << 21) + (e[s] >> 10) + 1;
        }
        double u = (double)m * 0x1p-52;
        double a = -as241_qnorm(0.5 * u);
        z[s] = si < 0 ? -a : a;
    }
}

/* matrix-valued g: two normals per block (one channel pair), 52-bit uniforms */
static void normal_pair(uint64_t seed, uint64_t path, uint32_t block, uint32_t channel, double z[2])
{
    uint32_t c[4] = { (uint32_t)path, (uint32_t)(path >> 32), block, channel };
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    z[0] = bits_to_normal(c[0], c[1]);
    z[1] = bits_to_normal(c[2], c[3]);
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = { ctr[0], ctr[1], ctr[2], ctr[3] };
    philox4x32_10(c, key[0], key[1]);
    memcpy(out, c, sizeof c);
}

/* ---- the ensemble: P independent euler()/heun() calls ------------------- */
#define MAXD 512
int oracle_ensemble(int scheme, int model, const double *params, int nx, int nB, int proj,
                    int nt, const double *times, const double *x0, int x0_stride,
                    const double *B, uint64_t seed, uint64_t path_offset, uint64_t npaths,
                    double *X, double *XT, int nthreads)
{
    if (nx > MAXD || nB > MAXD || nt < 2) return -1;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (uint64_t p = 0; p < npaths; ++p) {
        double x[MAXD], xn[MAXD], Y[MAXD], fX[MAXD], fY[MAXD], dB[MAXD];
        double *gX = (double *)malloc(sizeof(double) * 3 * (size_t)nx * nB), *gY = gX + (size_t)nx * nB,
               *gS = gY + (size_t)nx * nB;
        double zbuf[MAXD][2], zquad[4];
        for (int j = 0; j < nx; ++j) x[j] = x0[(size_t)p * x0_stride + j];
        if (X) for (int j = 0; j < nx; ++j) X[((size_t)p * nx + j) * nt] = x[j];         /* X[1,] <- x0 :441 */
        for (int i = 0; i < nt - 1; ++i) {
            double dt = times[i + 1] - times[i];                                         /* :446 */
            if (B) {
                for (int k = 0; k < nB; ++k) {                                           /* :437 */
                    const double *b = B + ((size_t)p * nB + k) * nt;
                    dB[k] = b[i + 1] - b[i];
                }
            } else if (nB == 1) {                                                        /* fused generator: one block = 4 steps */
                if ((i & 3) == 0) normal_quad(seed, path_offset + p, (uint32_t)(i >> 2), 0u, zquad);
                dB[0] = sqrt(dt) * zquad[i & 3];
            } else {                                                   /* matrix-valued g: one block = one step, 2 channels */
                double sq = sqrt(dt);
                for (int k = 0; k < nB; ++k) {
                    if ((k & 1) == 0) normal_pair(seed, path_offset + p, (uint32_t)i, 0x80000000u + (uint32_t)(k >> 1), zbuf[0]);
                    dB[k] = sq * zbuf[0][k & 1];
                }
            }
            drift(model, params, nx, x, fX);                                             /* :450 */
            diffusion(model, params, nx, nB, x, gX);                                     /* :451 */
            if (scheme == S_EULER) {
                for (int j = 0; j < nx; ++j)                                             /* :453 */
                    xn[j] = project(proj, (x[j] + fX[j] * dt) + matvec_row(gX, nx, nB, j, dB));
            } else {
                for (int j = 0; j < nx; ++j)                                             /* :318 */
                    Y[j] = project(proj, (x[j] + fX[j] * dt) + matvec_row(gX, nx, nB, j, dB));
                drift(model, params, nx, Y, fY);                                         /* :321 */
                diffusion(model, params, nx, nB, Y, gY);                                 /* :322 */
                for (size_t q = 0; q < (size_t)nx * nB; ++q) gS[q] = gX[q] + gY[q];
                for (int j = 0; j < nx; ++j)                                             /* :323 */
                    xn[j] = project(proj, (x[j] + (0.5 * (fX[j] + fY[j])) * dt) + 0.5 * matvec_row(gS, nx, nB, j, dB));
            }
            for (int j = 0; j < nx; ++j) x[j] = xn[j];
            if (X) for (int j = 0; j < nx; ++j) X[((size_t)p * nx + j) * nt + i + 1] = x[j];
        }
        if (XT) for (int j = 0; j < nx; ++j) XT[(size_t)p * nx + j] = x[j];
        free(gX);
    }
    return 0;
}

/* normals the fused generator hands to steps 0..nsteps-1 of one path/channel */
void oracle_normals(uint64_t seed, uint64_t path, uint32_t channel, int nsteps, double *z)
{
    double zz[4];
    for (int i = 0; i < nsteps; ++i) {
        if ((i & 3) == 0) normal_quad(seed, path, (uint32_t)(i >> 2), channel, zz);
        z[i] = zz[i & 3];
    }
}

/* ---- integrals on stored paths, batched over columns (n x ncols col-major) */
/* which: 0 itointegral(G,B)  1 stochint right  2 stochint centre  3 QV(X=a)  4 CrossVariation(a,b) */
void oracle_integral(int which, const double *a, const double *b, size_t n, size_t ncols, double *out)
{
    for (size_t c = 0; c < ncols; ++c) {
        const double *A = a + c * n, *Bc = b ? b + c * n : NULL;
        double *o = out + c * n;
        long double accl = 0.0L, accr = 0.0L;          /* R cumsum: LDOUBLE accumulator, cum.c */
        if (n) o[0] = 0.0;
        for (size_t i = 0; i + 1 < n; ++i) {
            double l, r;
            switch (which) {
            case 0: accl += (long double)(double)(A[i] * (Bc[i + 1] - Bc[i])); o[i + 1] = (double)accl; break;
            case 1: accr += (long double)(double)(A[i + 1] * (Bc[i + 1] - Bc[i])); o[i + 1] = (double)accr; break;
            case 2: {
                double dg = Bc[i + 1] - Bc[i];
                accl += (long double)(double)(A[i] * dg); accr += (long double)(double)(A[i + 1] * dg);
                l = (double)accl; r = (double)accr; o[i + 1] = 0.5 * (l + r);
            } break;
            case 3: { double d = A[i + 1] - A[i]; accl += (long double)(double)(d * d); o[i + 1] = (double)accl; } break;
            case 4: { double d = A[i + 1] - A[i], e = Bc[i + 1] - Bc[i];
                      accl += (long double)(double)(d * e); o[i + 1] = (double)accl; } break;
            }
        }
    }
}

/* x87 extended add of a double onto an 80-bit accumulator held as (mantissa, sign|exp):
 * exposes the real FPU result so tests can check the device's integer emulation. */
void oracle_ld_add(uint64_t *mant, uint16_t *sexp, double d)
{
    long double acc;
    unsigned char raw[16] = {0};
    memcpy(raw, mant, 8); memcpy(raw + 8, sexp, 2);
    memcpy(&acc, raw, sizeof(long double) < 16 ? sizeof(long double) : 16);
    acc += (long double)d;
    memset(raw, 0, 16); memcpy(raw, &acc, 10);
    memcpy(mant, raw, 8); memcpy(sexp, raw + 8, 2);
}

/* rBM :16-22 with supplied standard normals z[nt] */
void oracle_rBM(const double *times, int nt, double sigma, double B0, double u, const double *z, double *B)
{
    long double acc = 0.0L;
    for (int i = 0; i < nt; ++i) {
        double dt = i == 0 ? times[0] : times[i] - times[i - 1];
        double dB = u * dt + (sigma * sqrt(dt)) * z[i];
        acc += (long double)dB;
        B[i] = B0 + (double)acc;
    }
}
