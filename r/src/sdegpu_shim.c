/* sdegpu_shim.c — the .Call shim a SDEtools maintainer adds under src/ to run euler()/heun() ensembles
 * and the cumsum-based integrals on libsdegpu.so (include/sdegpu.h).
 *
 * The reference package has no native code (NAMESPACE:1-34); with this file NAMESPACE gains
 *     useDynLib(SDEtools, .registration = TRUE)
 * and R/sdegpu.R routes catalogue models here, everything else to the original R loops
 * (R/sdetools.R:375-470, :240-338), which therefore stay the fallback and the live oracle.
 *
 * Rules followed: plain C (Rf_error longjmps: no C++ frames above it); R objects are only touched from
 * the calling thread; inputs are read-only and used synchronously; outputs are allocated by R and
 * PROTECTed; every libsdegpu error becomes Rf_error(sdegpu_last_error()) after local cleanup; every R vector
 * is checked for type and length before its pointer is handed to the library; long runs poll
 * R_CheckUserInterrupt() between chunks through sdegpu_set_interrupt (under R_ToplevelExec, so the longjmp of
 * an interrupt never crosses the library's C++ frames).
 * R itself is not available in the build image: the file is compiled against tests/mock_r/Rinternals.h and
 * executed on tests/mock_r/mock_runtime.c by tests/test_r_shim.py.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "sdegpu.h"

#define MAX_DEVICES 64
static sdegpu_ctx *g_ctx = NULL;
static int g_devices[MAX_DEVICES], g_ndev = 0;

/* R_CheckUserInterrupt longjmps when the user pressed Ctrl-C; R_ToplevelExec catches that jump and returns FALSE */
static void check_interrupt(void *unused) { (void)unused; R_CheckUserInterrupt(); }
static int interrupt_pending(void *unused) { (void)unused; return R_ToplevelExec(check_interrupt, NULL) == FALSE; }

/* the context for opts$devices (NULL: the current one, or GPU 0); re-created when the device list changes */
static sdegpu_ctx *ctx_for(SEXP devices)
{
    int want[MAX_DEVICES], n = 0;
    if (!Rf_isNull(devices)) {
        if (XLENGTH(devices) < 1 || XLENGTH(devices) > MAX_DEVICES) Rf_error("sdegpu: devices must list 1..%d CUDA ordinals", MAX_DEVICES);
        n = (int)XLENGTH(devices);
        for (int k = 0; k < n; ++k) {
            const double v = TYPEOF(devices) == REALSXP ? REAL(devices)[k] : (double)INTEGER(devices)[k];
            if (!R_finite(v) || v < 0 || v > 1023 || v != floor(v)) Rf_error("sdegpu: devices must be CUDA ordinals (0-based integers)");
            want[k] = (int)v;
        }
    } else if (g_ctx) return g_ctx;
    else { n = 1; want[0] = 0; }
    if (g_ctx && n == g_ndev && memcmp(want, g_devices, sizeof(int) * (size_t)n) == 0) return g_ctx;
    if (g_ctx) { sdegpu_destroy(g_ctx); g_ctx = NULL; g_ndev = 0; }
    if (sdegpu_create_multi(want, n, &g_ctx) != SDEGPU_OK) Rf_error("sdegpu: %s", sdegpu_last_error());
    memcpy(g_devices, want, sizeof(int) * (size_t)n);
    g_ndev = n;
    sdegpu_set_interrupt(g_ctx, interrupt_pending, NULL);
    return g_ctx;
}
static sdegpu_ctx *ctx(void) { return ctx_for(R_NilValue); }

static SEXP list_get(SEXP list, const char *name)
{
    if (Rf_isNull(list)) return R_NilValue;
    SEXP names = Rf_getAttrib(list, R_NamesSymbol);
    if (Rf_isNull(names)) return R_NilValue;
    for (R_xlen_t i = 0; i < XLENGTH(list); ++i)
        if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
    return R_NilValue;
}

static double opt_real(SEXP opts, const char *name, double dflt)
{
    SEXP v = list_get(opts, name);
    return Rf_isNull(v) ? dflt : Rf_asReal(v);
}

/* an integer option in [lo, hi]; NA / non-finite / fractional values are R errors, not casts */
static int opt_int(SEXP opts, const char *name, int dflt, int lo, int hi)
{
    SEXP v = list_get(opts, name);
    if (Rf_isNull(v)) return dflt;
    const double d = Rf_asReal(v);
    if (!R_finite(d) || d != floor(d) || d < lo || d > hi) Rf_error("sdegpu: option '%s' must be an integer in [%d, %d]", name, lo, hi);
    return (int)d;
}

/* a double vector of exactly `n` elements (what = its name in messages), or NULL when absent and allowed */
static const double *real_arg(SEXP x, R_xlen_t n, const char *what, int may_be_null)
{
    if (Rf_isNull(x)) {
        if (may_be_null) return NULL;
        Rf_error("sdegpu: %s is missing", what);
    }
    if (TYPEOF(x) != REALSXP) Rf_error("sdegpu: %s must be a double vector (use as.numeric)", what);
    if (XLENGTH(x) != n) Rf_error("sdegpu: %s has length %ld, expected %ld", what, (long)XLENGTH(x), (long)n);
    return REAL(x);
}

/* seed = NULL: take 53 bits from R's own stream, so that set.seed() governs reproducibility AND the seed survives the
 * round trip through an R double (res$seed can be passed back in).  A user seed must be a whole number in [0, 2^53]. */
static uint64_t seed_from(SEXP opts)
{
    SEXP v = list_get(opts, "seed");
    if (!Rf_isNull(v)) {
        const double d = Rf_asReal(v);
        if (!R_finite(d) || d < 0 || d > 9007199254740992.0 || d != floor(d)) Rf_error("sdegpu: seed must be a whole number in [0, 2^53]");
        return (uint64_t)d;
    }
    GetRNGstate();
    const uint64_t hi = (uint64_t)(unif_rand() * 2097152.0), lo = (uint64_t)(unif_rand() * 4294967296.0);   /* 21 + 32 bits */
    PutRNGstate();
    return ((hi & 0x1FFFFFull) << 32) | (lo & 0xFFFFFFFFull);
}

static void fail_run(int np, const char *who, int rc)
{
    UNPROTECT(np);
    if (rc == SDEGPU_ERR_INTERRUPT) {
        R_CheckUserInterrupt();                 /* the pending interrupt, if still pending, is delivered here */
        Rf_error("%s: interrupted", who);
    }
    Rf_error("%s: %s", who, sdegpu_last_error());
}

/* .Call("C_sdegpu_run", scheme, model, params, times, x0, B, opts)
 *   scheme  0 euler (R/sdetools.R:375) / 1 heun (:240);  model, params: attributes of the catalogue closures
 *   x0      length nx (shared) or nx x npaths;  B: NULL, or nt x nB x npaths (the `B` argument, :431-437)
 *   opts    list(npaths, nx, nB, projection, arith, seed, output = 0 none/1 terminal/2 path, moments, center,
 *                hist = c(lo, hi, nbins, component), integrals, model_integrals, devices, path_offset,
 *                kill_kind/kill_comp/kill_a/kill_b, stop_kind/stop_comp/stop_lo/stop_hi, S0, Stau (npaths, or absent: one
 *                Philox uniform per path), forcing (nt x nx: c(times)), gscale (nt: s(times)), Spath)
 * returns list(X = nt x nx x npaths | NULL, XT = nx x npaths | NULL, power_sums = 5 x nx | NULL, hist = counts | NULL, seed,
 *              tau, S = survival at tau, integrals = 4 x nx x npaths | NULL, Spath = nt x npaths | NULL,
 *              model_integrals = 2 x nx x npaths | NULL)
 *   with a kill/stop rule X holds NA after each path's tau and XT is the state at tau */
SEXP C_sdegpu_run(SEXP scheme, SEXP model, SEXP params, SEXP times, SEXP x0, SEXP B, SEXP opts)
{
    sdegpu_problem pr;
    sdegpu_output out;
    memset(&pr, 0, sizeof pr);
    memset(&out, 0, sizeof out);
    pr.struct_size = sizeof pr;
    out.struct_size = sizeof out;
    const int nx = opt_int(opts, "nx", 1, 1, 1 << 20), nB = opt_int(opts, "nB", 1, 1, 1 << 20);
    if (TYPEOF(times) != REALSXP || XLENGTH(times) < 2 || XLENGTH(times) > INT_MAX) Rf_error("sdegpu: times must be a double vector of length >= 2");
    const int nt = (int)XLENGTH(times);
    const int npaths = opt_int(opts, "npaths", 1, 0, INT_MAX);      /* R arrays are indexed by int extents */
    const int output = opt_int(opts, "output", 2, 0, 2);
    const R_xlen_t np_ = (R_xlen_t)npaths;
    if (TYPEOF(params) != REALSXP) Rf_error("sdegpu: params must be a double vector");
    if (TYPEOF(x0) != REALSXP || (XLENGTH(x0) != nx && XLENGTH(x0) != (R_xlen_t)nx * np_))
        Rf_error("sdegpu: x0 has length %ld, expected nx = %d or nx*npaths = %ld", (long)XLENGTH(x0), nx, (long)((R_xlen_t)nx * np_));
    pr.scheme = Rf_asInteger(scheme);
    pr.model = Rf_asInteger(model);
    pr.arith = opt_int(opts, "arith", Rf_isNull(B) ? SDEGPU_ARITH_FAST : SDEGPU_ARITH_STRICT, 0, 1);
    pr.projection = opt_int(opts, "projection", SDEGPU_PROJ_IDENTITY, 0, 2);
    pr.nx = nx; pr.nB = nB;
    pr.params = REAL(params); pr.nparams = (int)XLENGTH(params);
    pr.times = REAL(times); pr.nt = nt;
    pr.x0 = REAL(x0); pr.x0_stride = XLENGTH(x0) == nx ? 0 : nx;
    pr.B = real_arg(B, (R_xlen_t)nt * nB * np_, "B (nt x nB x npaths)", 1);
    pr.npaths = (uint64_t)npaths;
    const double poff = opt_real(opts, "path_offset", 0);
    if (!R_finite(poff) || poff < 0 || poff > 9007199254740992.0) Rf_error("sdegpu: path_offset must be in [0, 2^53]");
    pr.path_offset = (uint64_t)poff;
    pr.seed = seed_from(opts);
    /* mortality / stopping: catalogue r(x), h(t,x) (R/sdetools.R:397-416,455-457) */
    pr.kill_kind = opt_int(opts, "kill_kind", SDEGPU_KILL_NONE, 0, 3);
    pr.kill_comp = opt_int(opts, "kill_comp", 0, 0, nx - 1);
    pr.kill_a = opt_real(opts, "kill_a", 0); pr.kill_b = opt_real(opts, "kill_b", 0);
    pr.stop_kind = opt_int(opts, "stop_kind", SDEGPU_STOP_NONE, 0, 3);
    pr.stop_comp = opt_int(opts, "stop_comp", 0, 0, nx - 1);
    pr.stop_lo = opt_real(opts, "stop_lo", 0); pr.stop_hi = opt_real(opts, "stop_hi", 0);
    pr.S0 = opt_real(opts, "S0", 1);
    pr.Stau = real_arg(list_get(opts, "Stau"), np_, "Stau (one threshold per path)", 1);
    /* time-dependent closures tabulated on the grid by the wrapper: f0(x) + c(t), s(t) * g0(x)  (:381-394) */
    pr.drift_forcing = real_arg(list_get(opts, "forcing"), (R_xlen_t)nt * nx, "forcing (nt x nx)", 1);
    pr.diffusion_scale = real_arg(list_get(opts, "gscale"), nt, "gscale (nt)", 1);
    const int killed = pr.kill_kind != SDEGPU_KILL_NONE || pr.stop_kind != SDEGPU_STOP_NONE;

    int np = 0;
    enum { NOUT = 10 };
    SEXP res = PROTECT(Rf_allocVector(VECSXP, NOUT)); ++np;
    SEXP nms = PROTECT(Rf_allocVector(STRSXP, NOUT)); ++np;
    const char *names[NOUT] = {"X", "XT", "power_sums", "hist", "seed", "tau", "S", "integrals", "Spath", "model_integrals"};
    for (int i = 0; i < NOUT; ++i) SET_STRING_ELT(nms, i, Rf_mkChar(names[i]));
    Rf_setAttrib(res, R_NamesSymbol, nms);
    if (output == 2) {
        if ((double)nt * nx * (double)npaths > 4503599627370496.0) Rf_error("sdegpu: trajectory array too large");
        SEXP X = PROTECT(Rf_alloc3DArray(REALSXP, nt, nx, npaths)); ++np;      /* nt x nx x npaths, column-major */
        SET_VECTOR_ELT(res, 0, X);
        out.X = REAL(X);
    }
    if (output >= 1) {
        SEXP XT = PROTECT(Rf_allocMatrix(REALSXP, nx, npaths)); ++np;
        SET_VECTOR_ELT(res, 1, XT);
        out.XT = REAL(XT);
    }
    if (opt_real(opts, "moments", 0) != 0) {
        SEXP ps = PROTECT(Rf_allocMatrix(REALSXP, 5, nx)); ++np;
        SET_VECTOR_ELT(res, 2, ps);
        out.power_sums = REAL(ps);
        out.center = real_arg(list_get(opts, "center"), nx, "center (one shift per component)", 1);
    }
    if (killed) {                                                   /* tau <- times[i+1], S at tau  (:467-468) */
        SEXP tau = PROTECT(Rf_allocVector(REALSXP, np_)); ++np;
        SEXP S = PROTECT(Rf_allocVector(REALSXP, np_)); ++np;
        SET_VECTOR_ELT(res, 5, tau);
        SET_VECTOR_ELT(res, 6, S);
        out.tau = REAL(tau);
        out.S_tau = REAL(S);
        if (opt_real(opts, "Spath", 0) != 0) {                      /* the S vector of list(times, X, S, tau)  (:443-444,456,468) */
            SEXP SP = PROTECT(Rf_allocMatrix(REALSXP, nt, npaths)); ++np;
            SET_VECTOR_ELT(res, 8, SP);
            out.S = REAL(SP);
        }
    }
    if (opt_real(opts, "integrals", 0) != 0) {          /* 4 x nx x npaths: QV, ito, strat, time — terminal entries */
        SEXP PI = PROTECT(Rf_alloc3DArray(REALSXP, 4, nx, npaths)); ++np;
        SET_VECTOR_ELT(res, 7, PI);
        out.path_integrals = REAL(PI);
    }
    if (opt_real(opts, "model_integrals", 0) != 0) {    /* 2 x nx x npaths: itointegral(f(X),t), itointegral(g(X),B) */
        SEXP MI = PROTECT(Rf_alloc3DArray(REALSXP, 2, nx, npaths)); ++np;
        SET_VECTOR_ELT(res, 9, MI);
        out.model_integrals = REAL(MI);
    }
    uint64_t *counts = NULL;
    SEXP h = list_get(opts, "hist");
    if (!Rf_isNull(h)) {
        if (TYPEOF(h) != REALSXP || XLENGTH(h) < 3) Rf_error("sdegpu: hist must be c(lo, hi, nbins[, component])");
        const double nb = REAL(h)[2], comp = XLENGTH(h) > 3 ? REAL(h)[3] : 0;
        if (!R_finite(nb) || nb < 1 || nb > 8192 || nb != floor(nb)) Rf_error("sdegpu: hist nbins must be an integer in [1, 8192]");
        if (!R_finite(comp) || comp < 0 || comp >= nx || comp != floor(comp)) Rf_error("sdegpu: hist component out of range");
        out.hist_lo = REAL(h)[0]; out.hist_hi = REAL(h)[1]; out.hist_nbins = (int)nb;
        out.hist_component = (int)comp;
        counts = (uint64_t *)R_alloc((size_t)out.hist_nbins + 3, sizeof(uint64_t));   /* freed by R at .Call exit */
        out.hist_counts = counts;
    }
    const int rc = sdegpu_run(ctx_for(list_get(opts, "devices")), &pr, &out);
    if (rc != SDEGPU_OK) fail_run(np, "sdegpu_run", rc);
    if (counts) {                                   /* R has no 64-bit integers: counts < 2^53 are exact as doubles */
        SEXP hc = PROTECT(Rf_allocVector(REALSXP, out.hist_nbins + 3)); ++np;
        for (int i = 0; i < out.hist_nbins + 3; ++i) REAL(hc)[i] = (double)counts[i];
        SET_VECTOR_ELT(res, 3, hc);
    }
    SEXP sd = PROTECT(Rf_ScalarReal((double)pr.seed)); ++np;       /* < 2^53: exact, can be passed back as `seed` */
    SET_VECTOR_ELT(res, 4, sd);
    UNPROTECT(np);
    return res;
}

static void dims(SEXP x, uint64_t *n, uint64_t *ncols)
{
    SEXP d = Rf_getAttrib(x, R_DimSymbol);
    if (Rf_isNull(d)) { *n = (uint64_t)XLENGTH(x); *ncols = 1; }
    else { *n = (uint64_t)INTEGER(d)[0]; *ncols = (uint64_t)(XLENGTH(x) / (*n ? *n : 1)); }
}

static void need_real_pair(SEXP a, SEXP b, const char *who)
{
    if (TYPEOF(a) != REALSXP || TYPEOF(b) != REALSXP) Rf_error("%s: arguments must be double vectors", who);
    if (XLENGTH(a) != XLENGTH(b)) Rf_error("%s: arguments must have the same length", who);
}

/* stochint(f, g, rule): rules is a 3-vector of 0/1 flags for l, r, c (R/sdetools.R:105-113) */
SEXP C_sdegpu_stochint(SEXP f, SEXP g, SEXP rules)
{
    uint64_t n, nc;
    need_real_pair(f, g, "stochint");
    if (TYPEOF(rules) != INTSXP || XLENGTH(rules) != 3) Rf_error("stochint: rules must be three integer flags");
    dims(f, &n, &nc);
    SEXP res = PROTECT(Rf_allocVector(VECSXP, 3));
    double *o[3] = {NULL, NULL, NULL};
    for (int k = 0; k < 3; ++k)
        if (INTEGER(rules)[k]) {
            SEXP v = PROTECT(Rf_duplicate(f));
            SET_VECTOR_ELT(res, k, v);
            UNPROTECT(1);
            o[k] = REAL(v);
        }
    const int rc = sdegpu_stochint(ctx(), REAL(f), REAL(g), n, nc, o[0], o[1], o[2], 0);
    if (rc != SDEGPU_OK) fail_run(1, "sdegpu_stochint", rc);
    UNPROTECT(1);
    return res;
}

/* itointegral(G, B)  R/sdetools.R:155-158 */
SEXP C_sdegpu_itointegral(SEXP G, SEXP B)
{
    uint64_t n, nc;
    need_real_pair(G, B, "itointegral");
    dims(G, &n, &nc);
    SEXP res = PROTECT(Rf_duplicate(G));
    const int rc = sdegpu_itointegral(ctx(), REAL(G), REAL(B), n, nc, REAL(res), 0);
    if (rc != SDEGPU_OK) fail_run(1, "sdegpu_itointegral", rc);
    UNPROTECT(1);
    return res;
}

/* QuadraticVariation(X)  R/sdetools.R:134-137 */
SEXP C_sdegpu_qv(SEXP X)
{
    uint64_t n, nc;
    if (TYPEOF(X) != REALSXP) Rf_error("QuadraticVariation: X must be a double vector");
    dims(X, &n, &nc);
    SEXP res = PROTECT(Rf_duplicate(X));
    const int rc = sdegpu_qv(ctx(), REAL(X), n, nc, REAL(res), 0);
    if (rc != SDEGPU_OK) fail_run(1, "sdegpu_qv", rc);
    UNPROTECT(1);
    return res;
}

/* CrossVariation(X, Y)  R/sdetools.R:184-187 */
SEXP C_sdegpu_crossvar(SEXP X, SEXP Y)
{
    uint64_t n, nc;
    need_real_pair(X, Y, "CrossVariation");
    dims(X, &n, &nc);
    SEXP res = PROTECT(Rf_duplicate(X));
    const int rc = sdegpu_crossvar(ctx(), REAL(X), REAL(Y), n, nc, REAL(res), 0);
    if (rc != SDEGPU_OK) fail_run(1, "sdegpu_crossvar", rc);
    UNPROTECT(1);
    return res;
}

static int count_arg(SEXP n, const char *who)
{
    const double d = Rf_asReal(n);
    if (!R_finite(d) || d < 0 || d > INT_MAX || d != floor(d)) Rf_error("%s: n must be a whole number in [0, %d]", who, INT_MAX);
    return (int)d;
}

/* rvBM(times, n, sigma, B0, u) for scalar sigma/B0/u  R/sdetools.R:40-46 */
SEXP C_sdegpu_rvbm(SEXP times, SEXP n, SEXP sigma, SEXP B0, SEXP u, SEXP opts)
{
    if (TYPEOF(times) != REALSXP || XLENGTH(times) < 1 || XLENGTH(times) > INT_MAX) Rf_error("rvBM: times must be a non-empty double vector");
    const int nt = (int)XLENGTH(times), nc = count_arg(n, "rvBM");
    SEXP res = PROTECT(Rf_allocMatrix(REALSXP, nt, nc));
    const int rc = sdegpu_rvbm(ctx(), REAL(times), nt, (uint64_t)nc, Rf_asReal(sigma), Rf_asReal(B0), Rf_asReal(u), seed_from(opts),
                               (uint64_t)opt_real(opts, "path_offset", 0), REAL(res), 0);
    if (rc != SDEGPU_OK) fail_run(1, "sdegpu_rvbm", rc);
    UNPROTECT(1);
    return res;
}

/* apply(rvBM(times, n, sigma, B0, u), 2, max) and friends without the n x nt matrix (vignettes/BrownianMotion.Rmd:101-144):
 * list(max, min, tmax, tmin, thit); opts$level (optional) is the first-passage level */
SEXP C_sdegpu_bm_functionals(SEXP times, SEXP n, SEXP sigma, SEXP B0, SEXP u, SEXP opts)
{
    if (TYPEOF(times) != REALSXP || XLENGTH(times) < 1 || XLENGTH(times) > INT_MAX) Rf_error("rvBM: times must be a non-empty double vector");
    const int nt = (int)XLENGTH(times), nc = count_arg(n, "rvBM.functionals");
    SEXP lvl = list_get(opts, "level");
    SEXP res = PROTECT(Rf_allocVector(VECSXP, 5));
    SEXP nms = PROTECT(Rf_allocVector(STRSXP, 5));
    const char *names[5] = {"max", "min", "tmax", "tmin", "thit"};
    double *out[5];
    for (int i = 0; i < 5; ++i) {
        SET_STRING_ELT(nms, i, Rf_mkChar(names[i]));
        SEXP v = PROTECT(Rf_allocVector(REALSXP, nc));
        SET_VECTOR_ELT(res, i, v);
        UNPROTECT(1);                                   /* held by res from here on */
        out[i] = REAL(v);
    }
    Rf_setAttrib(res, R_NamesSymbol, nms);
    const int rc = sdegpu_bm_functionals(ctx(), REAL(times), nt, (uint64_t)nc, Rf_asReal(sigma), Rf_asReal(B0), Rf_asReal(u),
                                         Rf_isNull(lvl) ? 0.0 : Rf_asReal(lvl), seed_from(opts), (uint64_t)opt_real(opts, "path_offset", 0),
                                         out[0], out[1], out[2], out[3], Rf_isNull(lvl) ? NULL : out[4], 0);
    if (rc != SDEGPU_OK) fail_run(2, "sdegpu_bm_functionals", rc);
    if (Rf_isNull(lvl)) SET_VECTOR_ELT(res, 4, R_NilValue);
    UNPROTECT(2);
    return res;
}

/* rOU(n,x0,lambda,xi,sigma,t) R/sdetools.R:784-791 and rCIR(n,x0,lambda,xi,gamma,t,Stratonovich) :695-702;
 * x0 is recycled to length n as R's rnorm/rchisq recycle their parameters; opts$steps > 1 applies the
 * transition repeatedly (vignettes/CoxIngersollRoss.Rmd:76-84). */
SEXP C_sdegpu_rlaw(SEXP law, SEXP n, SEXP x0, SEXP params, SEXP t, SEXP opts)
{
    const int nn = count_arg(n, "sdegpu_rlaw");
    if (TYPEOF(x0) != REALSXP || TYPEOF(params) != REALSXP) Rf_error("sdegpu_rlaw: x0 and params must be double vectors");
    const R_xlen_t nx0 = XLENGTH(x0);
    if (nx0 < 1 || XLENGTH(params) != 3) Rf_error("sdegpu_rlaw: bad arguments");
    SEXP res = PROTECT(Rf_allocVector(REALSXP, nn));
    SEXP start = PROTECT(Rf_allocVector(REALSXP, nn > 0 ? nn : 1));
    for (int i = 0; i < nn; ++i) REAL(start)[i] = REAL(x0)[i % nx0];
    const int rc = sdegpu_rlaw(ctx(), Rf_asInteger(law), REAL(params), Rf_asReal(t), opt_int(opts, "stratonovich", 0, 0, 1), REAL(start),
                               nx0 == 1 ? 0 : 1, (uint64_t)nn, opt_int(opts, "steps", 1, 0, INT_MAX), seed_from(opts),
                               (uint64_t)opt_real(opts, "path_offset", 0), REAL(res), NULL, 0);
    if (rc != SDEGPU_OK) fail_run(2, "sdegpu_rlaw", rc);
    UNPROTECT(2);
    return res;
}

/* dOU/pOU (R/sdetools.R:745-765) and dCIR/pCIR (:652-676) at the points x; x0 is recycled to length(x) */
SEXP C_sdegpu_law_eval(SEXP law, SEXP what, SEXP x, SEXP x0, SEXP params, SEXP t, SEXP stratonovich)
{
    if (TYPEOF(x) != REALSXP || TYPEOF(x0) != REALSXP || TYPEOF(params) != REALSXP) Rf_error("sdegpu_law_eval: x, x0 and params must be double vectors");
    const R_xlen_t n = XLENGTH(x), nx0 = XLENGTH(x0);
    if (nx0 < 1 || XLENGTH(params) != 3) Rf_error("sdegpu_law_eval: bad arguments");
    SEXP res = PROTECT(Rf_allocVector(REALSXP, n));
    SEXP start = PROTECT(Rf_allocVector(REALSXP, n > 0 ? n : 1));
    for (R_xlen_t i = 0; i < n; ++i) REAL(start)[i] = REAL(x0)[i % nx0];
    const int rc = sdegpu_law_eval(ctx(), Rf_asInteger(law), Rf_asInteger(what), REAL(params), Rf_asReal(t), Rf_asInteger(stratonovich),
                                   REAL(start), nx0 == 1 ? 0 : 1, REAL(x), (uint64_t)n, REAL(res), 0);
    if (rc != SDEGPU_OK) fail_run(2, "sdegpu_law_eval", rc);
    UNPROTECT(2);
    return res;
}

/* lyap(A, Q)  R/sdetools.R:593-600: the symmetric X with A X + X t(A) + Q = 0, d x d */
SEXP C_sdegpu_lyap(SEXP A, SEXP Q)
{
    need_real_pair(A, Q, "lyap");
    const int d = (int)floor(sqrt((double)XLENGTH(A)) + 0.5);
    if ((R_xlen_t)d * d != XLENGTH(A) || d < 1) Rf_error("lyap: A must be a square matrix");
    SEXP res = PROTECT(Rf_allocMatrix(REALSXP, d, d));
    const int rc = sdegpu_lyap(ctx(), REAL(A), REAL(Q), d, REAL(res));
    if (rc != SDEGPU_OK) fail_run(1, "sdegpu_lyap", rc);
    UNPROTECT(1);
    return res;
}

/* dLinSDE(A, G, t, x0, u, S0)  R/sdetools.R:508-572: list(eAt | EX, St); u is NULL, length nx (zero-order hold)
 * or nx x 2 (first-order hold); S0 NULL means 0*A */
SEXP C_sdegpu_dlinsde(SEXP A, SEXP G, SEXP t, SEXP x0, SEXP u, SEXP S0)
{
    if (TYPEOF(A) != REALSXP || TYPEOF(G) != REALSXP) Rf_error("dLinSDE: A and G must be double matrices");
    const int d = (int)floor(sqrt((double)XLENGTH(A)) + 0.5);
    if ((R_xlen_t)d * d != XLENGTH(A) || d < 1) Rf_error("dLinSDE: A must be a square matrix");
    if (XLENGTH(G) % d != 0 || XLENGTH(G) < d) Rf_error("dLinSDE: G must have as many rows as A");
    const int nB = (int)(XLENGTH(G) / d);
    const double *px0 = real_arg(x0, d, "x0", 1), *pS0 = real_arg(S0, (R_xlen_t)d * d, "S0", 1), *pu = NULL;
    int ucols = 0;
    if (!Rf_isNull(u)) {
        if (TYPEOF(u) != REALSXP || (XLENGTH(u) != d && XLENGTH(u) != 2 * (R_xlen_t)d)) Rf_error("dLinSDE: u must have length nx or be nx x 2");
        pu = REAL(u);
        ucols = (int)(XLENGTH(u) / d);
    }
    SEXP res = PROTECT(Rf_allocVector(VECSXP, 2));
    SEXP nms = PROTECT(Rf_allocVector(STRSXP, 2));
    SEXP first = PROTECT(px0 ? Rf_allocVector(REALSXP, d) : Rf_allocMatrix(REALSXP, d, d));
    SEXP St = PROTECT(Rf_allocMatrix(REALSXP, d, d));
    SET_STRING_ELT(nms, 0, Rf_mkChar(px0 ? "EX" : "eAt"));
    SET_STRING_ELT(nms, 1, Rf_mkChar("St"));
    SET_VECTOR_ELT(res, 0, first);
    SET_VECTOR_ELT(res, 1, St);
    Rf_setAttrib(res, R_NamesSymbol, nms);
    const int rc = sdegpu_dlinsde(ctx(), REAL(A), REAL(G), d, nB, Rf_asReal(t), px0, pu, ucols, pS0, px0 ? NULL : REAL(first),
                                  px0 ? REAL(first) : NULL, REAL(St));
    if (rc != SDEGPU_OK) fail_run(4, "sdegpu_dlinsde", rc);
    UNPROTECT(4);
    return res;
}

/* list(devices = <ordinals in use>, reduction = "nccl" | "host" | "none") of the shim's context */
SEXP C_sdegpu_info(void)
{
    sdegpu_ctx *c = ctx();
    SEXP res = PROTECT(Rf_allocVector(VECSXP, 2));
    SEXP nms = PROTECT(Rf_allocVector(STRSXP, 2));
    SEXP dev = PROTECT(Rf_allocVector(INTSXP, g_ndev));
    for (int k = 0; k < g_ndev; ++k) INTEGER(dev)[k] = g_devices[k];
    SEXP red = PROTECT(Rf_mkString(sdegpu_last_reduction(c)));
    SET_STRING_ELT(nms, 0, Rf_mkChar("devices"));
    SET_STRING_ELT(nms, 1, Rf_mkChar("reduction"));
    SET_VECTOR_ELT(res, 0, dev);
    SET_VECTOR_ELT(res, 1, red);
    Rf_setAttrib(res, R_NamesSymbol, nms);
    UNPROTECT(4);
    return res;
}

static const R_CallMethodDef call_methods[] = {
    {"C_sdegpu_run", (DL_FUNC)&C_sdegpu_run, 7},
    {"C_sdegpu_stochint", (DL_FUNC)&C_sdegpu_stochint, 3},
    {"C_sdegpu_itointegral", (DL_FUNC)&C_sdegpu_itointegral, 2},
    {"C_sdegpu_qv", (DL_FUNC)&C_sdegpu_qv, 1},
    {"C_sdegpu_crossvar", (DL_FUNC)&C_sdegpu_crossvar, 2},
    {"C_sdegpu_rvbm", (DL_FUNC)&C_sdegpu_rvbm, 6},
    {"C_sdegpu_bm_functionals", (DL_FUNC)&C_sdegpu_bm_functionals, 6},
    {"C_sdegpu_rlaw", (DL_FUNC)&C_sdegpu_rlaw, 6},
    {"C_sdegpu_law_eval", (DL_FUNC)&C_sdegpu_law_eval, 7},
    {"C_sdegpu_lyap", (DL_FUNC)&C_sdegpu_lyap, 2},
    {"C_sdegpu_dlinsde", (DL_FUNC)&C_sdegpu_dlinsde, 6},
    {"C_sdegpu_info", (DL_FUNC)&C_sdegpu_info, 0},
    {NULL, NULL, 0}};

/* R calls R_init_<package>: `sdegpu` as the stand-alone package under r/ (DESCRIPTION, NAMESPACE, src/Makevars), `SDEtools`
 * when the two files are merged into the reference package. */
void R_init_sdegpu(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
void R_init_SDEtools(DllInfo *dll) { R_init_sdegpu(dll); }

void R_unload_sdegpu(DllInfo *dll)
{
    (void)dll;
    if (g_ctx) sdegpu_destroy(g_ctx);
    g_ctx = NULL;
    g_ndev = 0;
}
void R_unload_SDEtools(DllInfo *dll) { R_unload_sdegpu(dll); }
