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
 * PROTECTed; every libsdegpu error becomes Rf_error(sdegpu_last_error()) after local cleanup.
 * Compile-checked against tests/mock_r/Rinternals.h (R itself is not available in the build image).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>

#include "sdegpu.h"

static sdegpu_ctx *g_ctx = NULL;

static sdegpu_ctx *ctx(void)
{
    if (!g_ctx && sdegpu_create(0, &g_ctx) != SDEGPU_OK) Rf_error("sdegpu: %s", sdegpu_last_error());
    return g_ctx;
}

static SEXP list_get(SEXP list, const char *name)
{
    SEXP names = Rf_getAttrib(list, R_NamesSymbol);
    for (R_xlen_t i = 0; i < XLENGTH(list); ++i)
        if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
    return R_NilValue;
}

static double opt_real(SEXP opts, const char *name, double dflt)
{
    SEXP v = list_get(opts, name);
    return Rf_isNull(v) ? dflt : Rf_asReal(v);
}

/* seed = NULL: take 64 bits from R's own stream so set.seed() governs reproducibility */
static uint64_t seed_from(SEXP opts)
{
    SEXP v = list_get(opts, "seed");
    if (!Rf_isNull(v)) return (uint64_t)Rf_asReal(v);
    GetRNGstate();
    uint64_t hi = (uint64_t)(unif_rand() * 4294967296.0), lo = (uint64_t)(unif_rand() * 4294967296.0);
    PutRNGstate();
    return (hi << 32) | lo;
}

/* .Call("C_sdegpu_run", scheme, model, params, times, x0, B, opts)
 *   scheme  0 euler (R/sdetools.R:375) / 1 heun (:240);  model, params: attributes of the catalogue closures
 *   x0      length nx (shared) or nx x npaths;  B: NULL, or nt x nB x npaths (the `B` argument, :431-437)
 *   opts    list(npaths, nx, nB, projection, arith, seed, output = 0 none/1 terminal/2 path, moments, center,
 *                hist = c(lo, hi, nbins, component))
 *                kill_kind/kill_comp/kill_a/kill_b, stop_kind/stop_comp/stop_lo/stop_hi, S0, Stau)
 * returns list(X = nt x nx x npaths | NULL, XT = nx x npaths | NULL, power_sums = 5 x nx | NULL, hist = counts | NULL, seed,
 *              tau, S, integrals = 4 x nx x npaths | NULL)   — with a kill/stop rule X holds NA after each path's tau and XT is the state at tau */
SEXP C_sdegpu_run(SEXP scheme, SEXP model, SEXP params, SEXP times, SEXP x0, SEXP B, SEXP opts)
{
    sdegpu_problem pr;
    sdegpu_output out;
    memset(&pr, 0, sizeof pr);
    memset(&out, 0, sizeof out);
    pr.struct_size = sizeof pr;
    out.struct_size = sizeof out;
    const int nx = (int)opt_real(opts, "nx", 1), nB = (int)opt_real(opts, "nB", 1);
    const int nt = (int)XLENGTH(times);
    const double npaths = opt_real(opts, "npaths", 1);
    const int output = (int)opt_real(opts, "output", 2);
    pr.scheme = Rf_asInteger(scheme);
    pr.model = Rf_asInteger(model);
    pr.arith = (int)opt_real(opts, "arith", Rf_isNull(B) ? SDEGPU_ARITH_FAST : SDEGPU_ARITH_STRICT);
    pr.projection = (int)opt_real(opts, "projection", SDEGPU_PROJ_IDENTITY);
    pr.nx = nx; pr.nB = nB;
    pr.params = REAL(params); pr.nparams = (int)XLENGTH(params);
    pr.times = REAL(times); pr.nt = nt;
    pr.x0 = REAL(x0); pr.x0_stride = XLENGTH(x0) == nx ? 0 : nx;
    pr.B = Rf_isNull(B) ? NULL : REAL(B);
    pr.npaths = (uint64_t)npaths;
    pr.path_offset = (uint64_t)opt_real(opts, "path_offset", 0);
    pr.seed = seed_from(opts);
    /* mortality / stopping: catalogue r(x), h(t,x) (R/sdetools.R:397-416,455-457) */
    pr.kill_kind = (int)opt_real(opts, "kill_kind", SDEGPU_KILL_NONE);
    pr.kill_comp = (int)opt_real(opts, "kill_comp", 0);
    pr.kill_a = opt_real(opts, "kill_a", 0); pr.kill_b = opt_real(opts, "kill_b", 0);
    pr.stop_kind = (int)opt_real(opts, "stop_kind", SDEGPU_STOP_NONE);
    pr.stop_comp = (int)opt_real(opts, "stop_comp", 0);
    pr.stop_lo = opt_real(opts, "stop_lo", 0); pr.stop_hi = opt_real(opts, "stop_hi", 0);
    pr.S0 = opt_real(opts, "S0", 1);
    SEXP stau = list_get(opts, "Stau");
    pr.Stau = Rf_isNull(stau) ? NULL : REAL(stau);              /* one threshold per path */
    const int killed = pr.kill_kind != SDEGPU_KILL_NONE || pr.stop_kind != SDEGPU_STOP_NONE;

    int np = 0;
    SEXP res = PROTECT(Rf_allocVector(VECSXP, 8)); ++np;
    SEXP nms = PROTECT(Rf_allocVector(STRSXP, 8)); ++np;
    const char *names[8] = {"X", "XT", "power_sums", "hist", "seed", "tau", "S", "integrals"};
    for (int i = 0; i < 8; ++i) SET_STRING_ELT(nms, i, Rf_mkChar(names[i]));
    Rf_setAttrib(res, R_NamesSymbol, nms);
    if (output == 2) {
        SEXP X = PROTECT(Rf_alloc3DArray(REALSXP, nt, nx, (int)npaths)); ++np;      /* nt x nx x npaths, column-major */
        SET_VECTOR_ELT(res, 0, X);
        out.X = REAL(X);
    }
    if (output >= 1) {
        SEXP XT = PROTECT(Rf_allocMatrix(REALSXP, nx, (int)npaths)); ++np;
        SET_VECTOR_ELT(res, 1, XT);
        out.XT = REAL(XT);
    }
    if (opt_real(opts, "moments", 0) != 0) {
        SEXP ps = PROTECT(Rf_allocMatrix(REALSXP, 5, nx)); ++np;
        SET_VECTOR_ELT(res, 2, ps);
        out.power_sums = REAL(ps);
        SEXP cen = list_get(opts, "center");
        out.center = Rf_isNull(cen) ? NULL : REAL(cen);
    }
    if (killed) {                                                   /* tau <- times[i+1], S at tau  (:467-468) */
        SEXP tau = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)npaths)); ++np;
        SEXP S = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)npaths)); ++np;
        SET_VECTOR_ELT(res, 5, tau);
        SET_VECTOR_ELT(res, 6, S);
        out.tau = REAL(tau);
        out.S_tau = REAL(S);
    }
    if (opt_real(opts, "integrals", 0) != 0) {          /* 4 x nx x npaths: QV, ito, strat, time — terminal entries */
        SEXP PI = PROTECT(Rf_alloc3DArray(REALSXP, 4, nx, (int)npaths)); ++np;
        SET_VECTOR_ELT(res, 7, PI);
        out.path_integrals = REAL(PI);
    }
    uint64_t *counts = NULL;
    SEXP h = list_get(opts, "hist");
    if (!Rf_isNull(h)) {
        out.hist_lo = REAL(h)[0]; out.hist_hi = REAL(h)[1]; out.hist_nbins = (int)REAL(h)[2];
        out.hist_component = XLENGTH(h) > 3 ? (int)REAL(h)[3] : 0;
        counts = (uint64_t *)R_alloc((size_t)out.hist_nbins + 3, sizeof(uint64_t));   /* freed by R at .Call exit */
        out.hist_counts = counts;
    }
    const int rc = sdegpu_run(ctx(), &pr, &out);
    if (rc != SDEGPU_OK) {
        UNPROTECT(np);
        Rf_error("sdegpu_run: %s", sdegpu_last_error());
    }
    if (counts) {                                   /* R has no 64-bit integers: counts < 2^53 are exact as doubles */
        SEXP hc = PROTECT(Rf_allocVector(REALSXP, out.hist_nbins + 3)); ++np;
        for (int i = 0; i < out.hist_nbins + 3; ++i) REAL(hc)[i] = (double)counts[i];
        SET_VECTOR_ELT(res, 3, hc);
    }
    SEXP sd = PROTECT(Rf_ScalarReal((double)pr.seed)); ++np;
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

/* stochint(f, g, rule): rules is a 3-vector of 0/1 flags for l, r, c (R/sdetools.R:105-113) */
SEXP C_sdegpu_stochint(SEXP f, SEXP g, SEXP rules)
{
    uint64_t n, nc;
    dims(f, &n, &nc);
    if (XLENGTH(g) != XLENGTH(f)) Rf_error("stochint: f and g must have the same length");
    SEXP res = PROTECT(Rf_allocVector(VECSXP, 3));
    double *o[3] = {NULL, NULL, NULL};
    for (int k = 0; k < 3; ++k)
        if (INTEGER(rules)[k]) {
            SEXP v = PROTECT(Rf_duplicate(f));
            SET_VECTOR_ELT(res, k, v);
            UNPROTECT(1);
            o[k] = REAL(v);
        }
    if (sdegpu_stochint(ctx(), REAL(f), REAL(g), n, nc, o[0], o[1], o[2], 0) != SDEGPU_OK) {
        UNPROTECT(1);
        Rf_error("sdegpu_stochint: %s", sdegpu_last_error());
    }
    UNPROTECT(1);
    return res;
}

/* itointegral(G, B)  R/sdetools.R:155-158 */
SEXP C_sdegpu_itointegral(SEXP G, SEXP B)
{
    uint64_t n, nc;
    dims(G, &n, &nc);
    if (XLENGTH(G) != XLENGTH(B)) Rf_error("itointegral: G and B must have the same length");
    SEXP res = PROTECT(Rf_duplicate(G));
    if (sdegpu_itointegral(ctx(), REAL(G), REAL(B), n, nc, REAL(res), 0) != SDEGPU_OK) {
        UNPROTECT(1);
        Rf_error("sdegpu_itointegral: %s", sdegpu_last_error());
    }
    UNPROTECT(1);
    return res;
}

/* QuadraticVariation(X)  R/sdetools.R:134-137 */
SEXP C_sdegpu_qv(SEXP X)
{
    uint64_t n, nc;
    dims(X, &n, &nc);
    SEXP res = PROTECT(Rf_duplicate(X));
    if (sdegpu_qv(ctx(), REAL(X), n, nc, REAL(res), 0) != SDEGPU_OK) {
        UNPROTECT(1);
        Rf_error("sdegpu_qv: %s", sdegpu_last_error());
    }
    UNPROTECT(1);
    return res;
}

/* CrossVariation(X, Y)  R/sdetools.R:184-187 */
SEXP C_sdegpu_crossvar(SEXP X, SEXP Y)
{
    uint64_t n, nc;
    dims(X, &n, &nc);
    if (XLENGTH(X) != XLENGTH(Y)) Rf_error("CrossVariation: X and Y must have the same length");
    SEXP res = PROTECT(Rf_duplicate(X));
    if (sdegpu_crossvar(ctx(), REAL(X), REAL(Y), n, nc, REAL(res), 0) != SDEGPU_OK) {
        UNPROTECT(1);
        Rf_error("sdegpu_crossvar: %s", sdegpu_last_error());
    }
    UNPROTECT(1);
    return res;
}

/* rvBM(times, n, sigma, B0, u) for scalar sigma/B0/u  R/sdetools.R:40-46 */
SEXP C_sdegpu_rvbm(SEXP times, SEXP n, SEXP sigma, SEXP B0, SEXP u, SEXP opts)
{
    const int nt = (int)XLENGTH(times), nc = Rf_asInteger(n);
    SEXP res = PROTECT(Rf_allocMatrix(REALSXP, nt, nc));
    if (sdegpu_rvbm(ctx(), REAL(times), nt, (uint64_t)nc, Rf_asReal(sigma), Rf_asReal(B0), Rf_asReal(u), seed_from(opts),
                    (uint64_t)opt_real(opts, "path_offset", 0), REAL(res), 0) != SDEGPU_OK) {
        UNPROTECT(1);
        Rf_error("sdegpu_rvbm: %s", sdegpu_last_error());
    }
    UNPROTECT(1);
    return res;
}

/* apply(rvBM(times, n, sigma, B0, u), 2, max) and friends without the n x nt matrix (vignettes/BrownianMotion.Rmd:101-144):
 * list(max, min, tmax, tmin, thit); opts$level (optional) is the first-passage level */
SEXP C_sdegpu_bm_functionals(SEXP times, SEXP n, SEXP sigma, SEXP B0, SEXP u, SEXP opts)
{
    const int nt = (int)XLENGTH(times), nc = Rf_asInteger(n);
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
    if (sdegpu_bm_functionals(ctx(), REAL(times), nt, (uint64_t)nc, Rf_asReal(sigma), Rf_asReal(B0), Rf_asReal(u),
                              Rf_isNull(lvl) ? 0.0 : Rf_asReal(lvl), seed_from(opts), (uint64_t)opt_real(opts, "path_offset", 0),
                              out[0], out[1], out[2], out[3], Rf_isNull(lvl) ? NULL : out[4], 0) != SDEGPU_OK) {
        UNPROTECT(2);
        Rf_error("sdegpu_bm_functionals: %s", sdegpu_last_error());
    }
    if (Rf_isNull(lvl)) SET_VECTOR_ELT(res, 4, R_NilValue);
    UNPROTECT(2);
    return res;
}

/* rOU(n,x0,lambda,xi,sigma,t) R/sdetools.R:784-791 and rCIR(n,x0,lambda,xi,gamma,t,Stratonovich) :695-702;
 * x0 is recycled to length n as R's rnorm/rchisq recycle their parameters; opts$steps > 1 applies the
 * transition repeatedly (vignettes/CoxIngersollRoss.Rmd:76-84). */
SEXP C_sdegpu_rlaw(SEXP law, SEXP n, SEXP x0, SEXP params, SEXP t, SEXP opts)
{
    const int nn = Rf_asInteger(n);
    const R_xlen_t nx0 = XLENGTH(x0);
    if (nn < 0 || nx0 < 1 || XLENGTH(params) != 3) Rf_error("sdegpu_rlaw: bad arguments");
    SEXP res = PROTECT(Rf_allocVector(REALSXP, nn));
    SEXP start = PROTECT(Rf_allocVector(REALSXP, nn > 0 ? nn : 1));
    for (int i = 0; i < nn; ++i) REAL(start)[i] = REAL(x0)[i % nx0];
    if (sdegpu_rlaw(ctx(), Rf_asInteger(law), REAL(params), Rf_asReal(t), (int)opt_real(opts, "stratonovich", 0), REAL(start),
                    nx0 == 1 ? 0 : 1, (uint64_t)nn, (int)opt_real(opts, "steps", 1), seed_from(opts),
                    (uint64_t)opt_real(opts, "path_offset", 0), REAL(res), NULL, 0) != SDEGPU_OK) {
        UNPROTECT(2);
        Rf_error("sdegpu_rlaw: %s", sdegpu_last_error());
    }
    UNPROTECT(2);
    return res;
}

/* dOU/pOU (R/sdetools.R:745-765) and dCIR/pCIR (:652-676) at the points x; x0 is recycled to length(x) */
SEXP C_sdegpu_law_eval(SEXP law, SEXP what, SEXP x, SEXP x0, SEXP params, SEXP t, SEXP stratonovich)
{
    const R_xlen_t n = XLENGTH(x), nx0 = XLENGTH(x0);
    if (nx0 < 1 || XLENGTH(params) != 3) Rf_error("sdegpu_law_eval: bad arguments");
    SEXP res = PROTECT(Rf_allocVector(REALSXP, n));
    SEXP start = PROTECT(Rf_allocVector(REALSXP, n > 0 ? n : 1));
    for (R_xlen_t i = 0; i < n; ++i) REAL(start)[i] = REAL(x0)[i % nx0];
    if (sdegpu_law_eval(ctx(), Rf_asInteger(law), Rf_asInteger(what), REAL(params), Rf_asReal(t), Rf_asInteger(stratonovich),
                        REAL(start), nx0 == 1 ? 0 : 1, REAL(x), (uint64_t)n, REAL(res), 0) != SDEGPU_OK) {
        UNPROTECT(2);
        Rf_error("sdegpu_law_eval: %s", sdegpu_last_error());
    }
    UNPROTECT(2);
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
    {NULL, NULL, 0}};

void R_init_SDEtools(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}

void R_unload_SDEtools(DllInfo *dll)
{
    (void)dll;
    if (g_ctx) sdegpu_destroy(g_ctx);
    g_ctx = NULL;
}
