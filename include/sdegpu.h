/* sdegpu.h — C ABI of libsdegpu.so, the B200 (sm_100a) ensemble SDE integrator
 * behind the SDEtools sample-path API.
 *
 * The reference (Uffe-H-Thygesen/SDEtools) is interpreted R with no FFI
 * (NAMESPACE:1-34 has no useDynLib), so this boundary is new surface.  Each
 * entry point names the reference interface it replaces; the R-side `.Call`
 * shim that binds them is r/src/sdegpu_shim.c.
 *
 * Conventions: extern "C", plain pointers and sizes, no C++/torch/R types, no
 * exceptions across the boundary.  Every function returns 0 on success or a
 * negative sdegpu_status; sdegpu_last_error() gives the thread-local message.
 * All matrices use the reference's layout: column-major, time fastest —
 *   X[(p*nx + j)*nt + i]   path p, state component j, time index i
 *   B[(p*nB + k)*nt + i]   path p, Brownian channel k, time index i
 * i.e. for one path X is exactly R's `nt x nx` matrix (R/sdetools.R:439-441,464)
 * and an ensemble is the `nt x nx x npaths` array sapply/replicate would build.
 * Pointers are host pointers unless SDEGPU_DEVICE_PTRS is set in `flags`.
 */
#ifndef SDEGPU_H
#define SDEGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDEGPU_ABI_VERSION 5

#if defined(__GNUC__)
#define SDEGPU_API __attribute__((visibility("default")))
#else
#define SDEGPU_API
#endif

typedef struct sdegpu_ctx sdegpu_ctx;

typedef enum {
    SDEGPU_OK = 0,
    SDEGPU_ERR_ARG = -1,      /* invalid argument / unsupported combination */
    SDEGPU_ERR_CUDA = -2,     /* CUDA runtime failure (message has the cudaError string) */
    SDEGPU_ERR_NOMEM = -3,    /* device or host allocation failed */
    SDEGPU_ERR_NODEVICE = -4, /* no sm_100 device / wrong architecture */
    SDEGPU_ERR_INTERRUPT = -5 /* the interrupt callback asked to stop (sdegpu_set_interrupt); outputs are undefined */
} sdegpu_status;

/* Model catalogue: drift/diffusion closures cannot be arbitrary R functions on
 * the device (reference: `f`,`g` arguments of euler/heun, R/sdetools.R:375,240),
 * so they are selected by id with a parameter vector.  Operation order is fixed
 * as the R closure would evaluate it.  All have nB = 1
 * (vector-valued g, R/sdetools.R:422-424) except LINEAR. */
typedef enum {
    SDEGPU_MODEL_OU = 0,         /* f = lambda*(xi-x)          g = sigma               params {lambda, xi, sigma} */
    SDEGPU_MODEL_CIR = 1,        /* f = lambda*(xi-x)          g = gamma*sqrt(abs(x))  params {lambda, xi, gamma} */
    SDEGPU_MODEL_VDP = 2,        /* f = (v, v*(mu-x*x)-x)      g = (0, sigma)          params {mu, sigma}; nx = 2 */
    SDEGPU_MODEL_LOGISTIC = 3,   /* f = x*(1-x) - c*(x*x)      g = sigma*x             params {sigma, c} */
    SDEGPU_MODEL_GBM = 4,        /* f = mu*x                   g = sigma*x             params {mu, sigma} */
    SDEGPU_MODEL_DOUBLEWELL = 5, /* f = x - x*x*x              g = sigma               params {sigma} */
    SDEGPU_MODEL_LINEAR = 6,     /* f = A %*% x                g = G (nx x nB)         params {A col-major, G col-major} */
    SDEGPU_MODEL_DWSQRT = 7,     /* f = x - x*x*x              g = sqrt(1 + x*x)       no params   (vignettes/ItoIntegrals.Rmd:66-69) */
    SDEGPU_MODEL_LOGLOGISTIC = 8 /* f = (1 - exp(y)) - 0.5*sigma*sigma   g = sigma     params {sigma}  (vignettes/ItosFormula.Rmd:113-116) */
} sdegpu_model;

typedef enum { SDEGPU_EULER = 0, SDEGPU_HEUN = 1 } sdegpu_scheme;            /* euler :375-470 / heun :240-338 */
typedef enum { SDEGPU_PROJ_IDENTITY = 0, SDEGPU_PROJ_ABS = 1, SDEGPU_PROJ_RELU = 2 } sdegpu_projection; /* `p` argument */

/* STRICT: one correctly rounded IEEE op per R operator, reference order, no FMA
 *         -> bit-identical to the reference arithmetic for nB = 1 and closures made of + - * / sqrt abs.
 *         Closures that call exp() (SDEGPU_MODEL_LOGLOGISTIC, SDEGPU_KILL_EXP, the survival update exp(-r dt)) agree to
 *         exp()'s last ulp only: CUDA's exp and the C library's are each within 1 ulp but not identical.
 * FAST:   FMA contraction and algebraically equivalent regrouping allowed (1e-12 relative); the linear model then
 *         runs on the FP64 tensor cores where its shape allows (5 <= nx <= 512). */
typedef enum { SDEGPU_ARITH_STRICT = 0, SDEGPU_ARITH_FAST = 1 } sdegpu_arith;

/* Killing rate r(x) and stopping function h(t,x) of euler()/heun() (R/sdetools.R:397-416,455-457;
 * vignettes/Killing.Rmd:152-190), again from a catalogue; `comp` selects the state component.
 *   survival  S_{i+1} = S_i * exp(-r(X_i) * dt_i), evaluated at the OLD state; the path is killed when S < Stau
 *   stopping  the path stops at the first i with h(t_{i+1}, X_{i+1}) <= 0 (S is then not updated: it reads 0) */
typedef enum {
    SDEGPU_KILL_NONE = 0,
    SDEGPU_KILL_CONST = 1,     /* r = a */
    SDEGPU_KILL_EXP = 2,       /* r = a * exp(b * x)        (Killing.Rmd: 0.5*exp(2*x)) */
    SDEGPU_KILL_QUAD = 3       /* r = a * (x - b) * (x - b) */
} sdegpu_kill;
typedef enum {
    SDEGPU_STOP_NONE = 0,
    SDEGPU_STOP_BELOW = 1,     /* h = x - lo          stop when x <= lo */
    SDEGPU_STOP_ABOVE = 2,     /* h = hi - x          stop when x >= hi */
    SDEGPU_STOP_OUTSIDE = 3    /* h = (x - lo) * (hi - x)   stop on leaving (lo, hi) */
} sdegpu_stop;

enum {
    SDEGPU_DEVICE_PTRS = 1u,  /* B, per-path x0 and every output buffer are device pointers */
    SDEGPU_ACCUMULATE = 2u,   /* add into power_sums / hist_counts instead of overwriting */
    SDEGPU_EXACT_GRID = 4u    /* fused-generator runs (no B, no kill/stop rule) treat a grid whose increments agree to the
                                 rounding of its time stamps (|dt_i - mean dt| <= 4 eps max|t|, e.g. seq(0, T, by = h)) as
                                 the uniform grid dt = (t_n - t_0)/n; this flag keeps every dt_i = times[i+1]-times[i] */
};

/* One euler()/heun() ensemble: npaths independent calls of the reference function. */
typedef struct {
    uint32_t struct_size;     /* = sizeof(sdegpu_problem), for forward compatibility */
    int32_t model;            /* sdegpu_model */
    int32_t scheme;           /* sdegpu_scheme */
    int32_t arith;            /* sdegpu_arith */
    int32_t projection;       /* sdegpu_projection */
    int32_t nx, nB;           /* length(x0); Brownian channels (R/sdetools.R:377,420-428) */
    int32_t nparams;
    const double *params;     /* host, nparams doubles */
    int32_t nt;               /* length(times) >= 2 */
    int32_t x0_stride;        /* 0: one x0 shared by all paths (host ptr); nx: x0[p*nx+j] per path */
    const double *times;      /* host, nt doubles; dt_i = times[i+1]-times[i] is never assumed constant (:446) */
    const double *x0;
    const double *B;          /* Brownian paths (`B` argument, :431-437) or NULL: increments are then drawn in-kernel
                                 from Philox4x32-10, dB = sqrt(dt_i)*Z, keyed by (seed; path_offset+p, step, channel) */
    uint64_t npaths;
    uint64_t path_offset;     /* global id of path 0: shards of one ensemble use disjoint id ranges */
    uint64_t seed;
    uint32_t flags;
    uint32_t reserved;
    /* --- ABI 2: mortality and stopping (`r`, `h`, `S0`, `Stau` arguments) --- */
    int32_t kill_kind, kill_comp;   /* sdegpu_kill */
    double kill_a, kill_b;
    int32_t stop_kind, stop_comp;   /* sdegpu_stop */
    double stop_lo, stop_hi;
    double S0;                /* initial survival (reference default 1) */
    const double *Stau;       /* per-path thresholds Stau[p] (the reference draws runif(1) per call), or NULL: drawn
                                 in-kernel from Philox, counter (path, 0, 0xC0000000) */
    /* --- ABI 5: time-dependent drift and diffusion.  The reference takes f(t,x) and g(t,x) (R/sdetools.R:381-394) and
     * evaluates them on the grid only: f(times[i], X_i), and in heun() also f(times[i+1], Y) (:316-323).  A closure of the
     * form  f(t,x) = f0(x) + c(t),  g(t,x) = s(t) * g0(x)  with catalogue f0, g0 is therefore fully described by its
     * values on the grid, which the caller tabulates (the R wrapper calls the closures c, s at `times`):
     *   drift_forcing   host, nt x nx column-major, c[j*nt + i] = c_j(times[i])   e.g. F*sin(2*t) of the example at :211,:363
     *   diffusion_scale host, nt, s[i] = s(times[i])
     * Either may be NULL.  Operation order: fX = f0(X) + c_i (one rounded add after the catalogue drift), gX = s_i * g0(X). */
    const double *drift_forcing;
    const double *diffusion_scale;
} sdegpu_problem;

/* Caller-allocated outputs; any pointer may be NULL. */
typedef struct {
    uint32_t struct_size;     /* = sizeof(sdegpu_output) */
    int32_t hist_component;   /* state component the histogram is taken on */
    double *X;                /* nt*nx*npaths: the `X` of list(times=,X=) for every path (:464) */
    double *XT;               /* nx*npaths: terminal state X[nt,] of every path, XT[p*nx+j] */
    double *power_sums;       /* 5*nx: per component {n, S1, S2, S3, S4}, S_k = sum_p (X_T - center[j])^k over finite X_T */
    const double *center;     /* host, nx doubles (NULL = 0): shift for the power sums */
    uint64_t *hist_counts;    /* hist_nbins+3: {under, bins..., over, nan}; equal-width breaks on [lo,hi] with R's
                                 hist(right=TRUE, include.lowest=TRUE) convention: bin k = (lo+k*w, lo+(k+1)*w], lo in bin 0 */
    double hist_lo, hist_hi;
    int32_t hist_nbins;
    int32_t reserved;
    float *elapsed_ms;        /* host: CUDA-event time of the kernels of this call (no copies) */
    /* --- ABI 2: with kill_kind/stop_kind set, XT is the state at tau and X holds NaN after tau --- */
    double *tau;              /* npaths: times[i+1] at the stop/kill step, or times[nt-1] (R/sdetools.R:467) */
    double *S_tau;            /* npaths: survival S at tau (0 after an h-stop, like the reference's S vector) */
    /* --- ABI 3: integrals along each path, accumulated in the step loop (no trajectory stored; X must be NULL) --- */
    double *path_integrals;   /* 4*nx*npaths, [(p*nx+j)*4 + k]: the terminal entries of
                                 k=0 QuadraticVariation(X_j) (R/sdetools.R:134-137), k=1 itointegral(X_j,B) (:155-158),
                                 k=2 stochint(X_j,B,"c") (:105-113), k=3 itointegral(X_j,times); 80-bit cumsum each.
                                 With a supplied B bit-identical to those functions run on the stored path. */
    /* --- ABI 4: the survival path of a mortality run, the `S` of list(times, X, S, tau) (R/sdetools.R:443-444,456,468; heun :336) --- */
    double *S;                /* nt*npaths, S[p*nt + i]: S[0] = S0, S_{i+1} = S_i exp(-r(X_i) dt_i) up to tau; after an h-stop the
                                 entry at tau reads 0 (the reference's S <- numeric(nt) is never written there); NaN after tau.
                                 For one path S[0 .. k-1] with times[k-1] = tau is exactly the reference's S[1:(i+1)]. */
    /* --- ABI 5: integrals of the model's own drift and diffusion along each path, accumulated in the step loop like
     * path_integrals (X must be NULL):  X_t - X_0 = itointegral(f(X),t) + itointegral(g(X),B) for an euler() path, the
     * machine-precision identity of vignettes/ItoIntegrals.Rmd:82-93 --- */
    double *model_integrals;  /* 2*nx*npaths, [(p*nx+j)*2 + k]: terminal entries of k=0 itointegral(f_j(X), times),
                                 k=1 itointegral(g_j(X), B)  (R/sdetools.R:155-158), f and g evaluated at (times[i], X_i) */
} sdegpu_output;

typedef struct {
    char name[64];
    int32_t sm_count, cc_major, cc_minor, clock_khz;
    uint64_t global_mem_bytes;
} sdegpu_devinfo;

SDEGPU_API int sdegpu_abi_version(void);
SDEGPU_API const char *sdegpu_last_error(void);

/* One context per process per GPU.  `device` is the CUDA ordinal. */
SDEGPU_API int sdegpu_create(int device, sdegpu_ctx **ctx);
/* One context over several GPUs of this process (sdegpu_create(devices, ndev); 8(e)).  devices = NULL
 * means ordinals 0..ndev-1.  sdegpu_run on such a context splits the ensemble into ndev contiguous path ranges, one
 * per device, each keyed by its global path ids (so every per-path output, and every histogram count, is identical to a
 * one-device run); each range runs on its own host thread and stream; power sums and histogram counts are combined by
 * ONE all-reduce over NVLink (NCCL, single-process ncclCommInitAll; loaded at run time) or, when libnccl cannot be
 * loaded, summed on the host in device order.  Host pointers only (SDEGPU_DEVICE_PTRS is refused for ndev > 1); the other
 * entry points run on devices[0]. */
SDEGPU_API int sdegpu_create_multi(const int *devices, int ndev, sdegpu_ctx **ctx);
SDEGPU_API int sdegpu_device_count(sdegpu_ctx *ctx);            /* ndev of the context */
/* "nccl" / "host" / "none": how the last multi-device sdegpu_run combined its accumulators */
SDEGPU_API const char *sdegpu_last_reduction(sdegpu_ctx *ctx);
/* Long host-pointer runs are cut into path ranges (chunks) that go through the device one after another: inputs and
 * outputs move through pinned staging buffers on their own copy streams while the next chunk computes, and between
 * chunks `fn(user)` is called ON THE CALLING THREAD (R's API is single-threaded: the shim's callback wraps
 * R_CheckUserInterrupt).  A non-zero return stops the run with SDEGPU_ERR_INTERRUPT.  fn = NULL removes it.
 * Chunking never changes a per-path result or a histogram count: the generator is keyed by global path id. */
typedef int (*sdegpu_interrupt_fn)(void *user);
SDEGPU_API int sdegpu_set_interrupt(sdegpu_ctx *ctx, sdegpu_interrupt_fn fn, void *user);
/* Paths per chunk of host-pointer runs: 0 = automatic (runs with more than 512 MB of per-path I/O: a quarter of the run
 * per chunk, between 256 MB and 4 GB; runs without bulk I/O: about 2e10 path-steps per chunk when an interrupt callback
 * is set, otherwise one launch). */
SDEGPU_API int sdegpu_set_chunk_paths(sdegpu_ctx *ctx, uint64_t paths);
SDEGPU_API int sdegpu_destroy(sdegpu_ctx *ctx);
SDEGPU_API int sdegpu_device_info(sdegpu_ctx *ctx, sdegpu_devinfo *info);
/* Launch on a caller-owned cudaStream_t (e.g. torch's current stream; NULL = CUDA's default stream);
 * sdegpu_reset_stream goes back to the context's own non-blocking stream. */
SDEGPU_API int sdegpu_set_stream(sdegpu_ctx *ctx, void *cuda_stream);
SDEGPU_API int sdegpu_reset_stream(sdegpu_ctx *ctx);
SDEGPU_API int sdegpu_synchronize(sdegpu_ctx *ctx);

/* Replaces sapply/replicate around euler() (R/sdetools.R:375-470) or heun() (:240-338).
 * Host pointers: blocking; large runs move through the chunk pipeline (pinned staging, copy streams) and poll the
 * interrupt callback between chunks.  SDEGPU_DEVICE_PTRS: the per-run tables (times, parameters) are uploaded with one
 * stream synchronisation, then the kernels are only enqueued on the stream (unless elapsed_ms is requested). */
SDEGPU_API int sdegpu_run(sdegpu_ctx *ctx, const sdegpu_problem *prob, sdegpu_output *out);

/* Stochastic integrals on stored paths, batched over ncols columns of length n
 * (column c at a + c*n).  Partial sums follow R's cumsum exactly: an x87 80-bit
 * accumulator rounded to double per element (emulated in integer arithmetic on
 * the device), so results are bit-identical to the reference on x86-64.
 *   sdegpu_stochint   stochint(f,g,rule)        R/sdetools.R:105-113; l, r, c may each be NULL
 *   sdegpu_itointegral itointegral(G,B)         R/sdetools.R:155-158
 *   sdegpu_qv         QuadraticVariation(X)     R/sdetools.R:134-137
 *   sdegpu_crossvar   CrossVariation(X,Y)       R/sdetools.R:184-187 */
SDEGPU_API int sdegpu_stochint(sdegpu_ctx *ctx, const double *f, const double *g, uint64_t n, uint64_t ncols,
                    double *l, double *r, double *c, uint32_t flags);
SDEGPU_API int sdegpu_itointegral(sdegpu_ctx *ctx, const double *G, const double *B, uint64_t n, uint64_t ncols,
                       double *out, uint32_t flags);
SDEGPU_API int sdegpu_qv(sdegpu_ctx *ctx, const double *X, uint64_t n, uint64_t ncols, double *out, uint32_t flags);
SDEGPU_API int sdegpu_crossvar(sdegpu_ctx *ctx, const double *X, const double *Y, uint64_t n, uint64_t ncols,
                    double *out, uint32_t flags);

/* rvBM(times, n) for an ensemble (R/sdetools.R:16-22,40-46): column c is
 * B0 + cumsum(N(u*dt, sigma^2 dt)) with dt = c(times[1], diff(times)), 80-bit
 * cumsum, normals from the fused Philox generator keyed by (seed; path_offset+c). */
SDEGPU_API int sdegpu_rvbm(sdegpu_ctx *ctx, const double *times, int32_t nt, uint64_t ncols, double sigma, double B0,
                double u, uint64_t seed, uint64_t path_offset, double *B, uint32_t flags);

/* Path functionals of the rvBM ensemble without storing it (apply(rvBM(0:Nt,Np),2,max),
 * vignettes/BrownianMotion.Rmd:101-144).  Column c is regenerated exactly as sdegpu_rvbm builds it (same
 * arguments, same seed => same doubles), so each output equals the functional of the stored matrix bit for bit:
 * vmax/vmin = max/min over the grid, tmax/tmin = times[] of the first maximum/minimum (which.max), thit = first
 * times[i] with B_i >= level (level >= B0) or B_i <= level (level < B0), NaN if the level is never reached.
 * Any output may be NULL. */
SDEGPU_API int sdegpu_bm_functionals(sdegpu_ctx *ctx, const double *times, int32_t nt, uint64_t ncols, double sigma,
                double B0, double u, double level, uint64_t seed, uint64_t path_offset, double *vmax, double *vmin,
                double *tmax, double *tmin, double *thit, uint32_t flags);

/* Exact transition-law samplers and their recursive use:
 *   law = SDEGPU_LAW_OU   rOU(n,x0,lambda,xi,sigma,t)                R/sdetools.R:784-791 (ou.parameters :706-712)
 *   law = SDEGPU_LAW_CIR  rCIR(n,x0,lambda,xi,gamma,t,Stratonovich)  R/sdetools.R:695-702 (cir.parameters :606-617)
 * params = {lambda, xi, sigma|gamma}.  Draw p starts at x0[p*x0_stride] (x0_stride 0: one shared start, as the
 * reference's scalar x0) and takes `nsteps` transitions of length t, X_{k+1} ~ law(. | X_k, t)
 * (vignettes/CoxIngersollRoss.Rmd:76-84; nsteps = 1 is the reference call).  out[p] = final state;
 * path (optional) [p*(nsteps+1)+i] = state after i transitions.  Randomness is a pure function of
 * (seed, path_offset+p, transition index); the OU normals are those euler()/heun() hand to the same path
 * and step under the same seed.  x0, out, path are host pointers unless SDEGPU_DEVICE_PTRS. */
typedef enum { SDEGPU_LAW_OU = 0, SDEGPU_LAW_CIR = 1 } sdegpu_law;
SDEGPU_API int sdegpu_rlaw(sdegpu_ctx *ctx, int32_t law, const double params[3], double t, int32_t stratonovich,
                const double *x0, int64_t x0_stride, uint64_t n, int32_t nsteps, uint64_t seed, uint64_t path_offset,
                double *out, double *path, uint32_t flags);

/* Evaluation of the same laws at n points: what = 0 density dOU/dCIR (R/sdetools.R:745-752, :652-663), what = 1
 * distribution function pOU/pCIR (:758-765, :669-676; lower tail, not log).  out[p] = law(x[p] | x0[p*x0_stride]).
 * On device pointers this is the probability integral transform of an ensemble's terminal states without a
 * round trip through the host. */
SDEGPU_API int sdegpu_law_eval(sdegpu_ctx *ctx, int32_t law, int32_t what, const double params[3], double t,
                int32_t stratonovich, const double *x0, int64_t x0_stride, const double *x, uint64_t n, double *out,
                uint32_t flags);

/* Linear-SDE laws at the dimension the device integrates.  The reference's lyap() solves the
 * d^2 x d^2 Kronecker system (O(d^6), R/sdetools.R:593-600) and cannot validate a d = 256 run; these are O(d^3) and
 * GEMM-shaped, with the matrix products on the GPU.  All matrices are HOST pointers, column-major (R's layout).
 *   sdegpu_lyap     lyap(A, Q): X with A X + X t(A) + Q = 0, d x d.  d <= 40: the reference's Kronecker system itself;
 *                   larger: bilinear transform + doubling, which needs the spectrum of A strictly on one side of the
 *                   imaginary axis (SDEGPU_ERR_ARG otherwise, like the reference's solve() error for a singular system).
 *   sdegpu_dlinsde  dLinSDE(A, G, t, x0, u, S0) (:508-572).  G is d x nB.  S0 NULL = 0*A.  x0 NULL: eAt (d x d) and St are
 *                   written (list(eAt, St), :536-538); else EX (d) and St.  u NULL / u_cols = 1 (zero-order hold, :545-553)
 *                   / u_cols = 2 (first-order hold from u[,1] to u[,2], :559-571).  St follows the reference: Sinf from the
 *                   Lyapunov equation when it has a solution, Van Loan's block exponential otherwise (:520-534). */
SDEGPU_API int sdegpu_lyap(sdegpu_ctx *ctx, const double *A, const double *Q, int32_t d, double *X);
SDEGPU_API int sdegpu_dlinsde(sdegpu_ctx *ctx, const double *A, const double *G, int32_t d, int32_t nB, double t, const double *x0,
                   const double *u, int32_t u_cols, const double *S0, double *eAt, double *EX, double *St);

/* Generator introspection (tests / known-answer vectors). */
SDEGPU_API int sdegpu_philox4x32_10(sdegpu_ctx *ctx, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* z[i*npaths + p] = standard normal handed to step i, channel `channel`, of path path_offset+p (host ptr). */
SDEGPU_API int sdegpu_normals(sdegpu_ctx *ctx, uint64_t seed, uint64_t path_offset, uint64_t npaths, uint32_t channel,
                   int32_t nsteps, double *z);

/* Roofline denominators measured on this device: a dependent-free DFMA stream
 * (TFLOP/s, 2 flops per FMA) and a device-to-device copy (GB/s, read+write). */
SDEGPU_API int sdegpu_measure_fp64_peak(sdegpu_ctx *ctx, double *tflops);
SDEGPU_API int sdegpu_measure_hbm_peak(sdegpu_ctx *ctx, double *gbs);

#ifdef __cplusplus
}
#endif
#endif /* SDEGPU_H */
