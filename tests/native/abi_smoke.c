/* The C ABI used the way the R shim uses it: plain C, host pointers, no Python in between.
 * OU euler() with a supplied B on 3 paths x 6 time points against the reference recursion
 * X[i+1] = X[i] + lambda*(xi - X[i])*dt + sigma*(B[i+1]-B[i])  (R/sdetools.R:450-453), one rounded operation
 * per R operator; then itointegral() and the error path.  Prints "abi smoke ok" and returns 0 on success. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "sdegpu.h"

int main(void)
{
    sdegpu_ctx *ctx = NULL;
    if (sdegpu_abi_version() != SDEGPU_ABI_VERSION) { printf("ABI mismatch\n"); return 1; }
    if (sdegpu_create(0, &ctx) != SDEGPU_OK) { printf("create: %s\n", sdegpu_last_error()); return 1; }
    enum { NT = 6, P = 3 };
    const double lambda = 1.3, xi = 0.7, sigma = 0.9, x0 = 2.0;
    double times[NT], B[P * NT], X[P * NT], want[P * NT], params[3] = {lambda, xi, sigma};
    for (int i = 0; i < NT; ++i) times[i] = 0.1 * i * (1.0 + 0.01 * i);            /* non-uniform grid */
    for (int p = 0; p < P; ++p)
        for (int i = 0; i < NT; ++i) B[p * NT + i] = i == 0 ? 0.0 : B[p * NT + i - 1] + 0.3 * sin(1.0 + 3.0 * p + 0.7 * i);
    for (int p = 0; p < P; ++p) {
        volatile double x = x0;                                                      /* volatile: no contraction, no x87 */
        want[p * NT] = x;
        for (int i = 0; i + 1 < NT; ++i) {
            volatile double dt = times[i + 1] - times[i], dB = B[p * NT + i + 1] - B[p * NT + i];
            volatile double a = xi - x, f = lambda * a, fdt = f * dt, gdB = sigma * dB, s = x + fdt;
            x = s + gdB;
            want[p * NT + i + 1] = x;
        }
    }
    sdegpu_problem pr;
    sdegpu_output out;
    memset(&pr, 0, sizeof pr);
    memset(&out, 0, sizeof out);
    pr.struct_size = sizeof pr; out.struct_size = sizeof out;
    pr.model = SDEGPU_MODEL_OU; pr.scheme = SDEGPU_EULER; pr.arith = SDEGPU_ARITH_STRICT; pr.projection = SDEGPU_PROJ_IDENTITY;
    pr.nx = 1; pr.nB = 1; pr.nparams = 3; pr.params = params; pr.nt = NT; pr.times = times; pr.x0 = &x0; pr.x0_stride = 0;
    pr.B = B; pr.npaths = P; pr.S0 = 1.0;
    out.X = X;
    if (sdegpu_run(ctx, &pr, &out) != SDEGPU_OK) { printf("run: %s\n", sdegpu_last_error()); return 1; }
    if (memcmp(X, want, sizeof X) != 0) { printf("euler differs from the reference recursion\n"); return 1; }
    double I[P * NT];
    if (sdegpu_itointegral(ctx, X, B, NT, P, I, 0) != SDEGPU_OK) { printf("itointegral: %s\n", sdegpu_last_error()); return 1; }
    for (int p = 0; p < P; ++p) {
        long double acc = 0.0L;                                                      /* R's cumsum accumulator */
        if (I[p * NT] != 0.0) { printf("itointegral[0] != 0\n"); return 1; }
        for (int i = 0; i + 1 < NT; ++i) {
            volatile double dB = B[p * NT + i + 1] - B[p * NT + i], term = X[p * NT + i] * dB;
            acc += (long double)term;
            if (I[p * NT + i + 1] != (double)acc) { printf("itointegral differs at %d,%d\n", p, i); return 1; }
        }
    }
    pr.nt = 1;                                                                       /* error path: status + message, no abort */
    if (sdegpu_run(ctx, &pr, &out) == SDEGPU_OK || strstr(sdegpu_last_error(), "at least 2 time points") == NULL) {
        printf("error path: %s\n", sdegpu_last_error());
        return 1;
    }
    sdegpu_destroy(ctx);
    printf("abi smoke ok\n");
    return 0;
}
