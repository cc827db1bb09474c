// C ABI of libsdegpu.so (include/sdegpu.h).  No R / torch / C++ types cross this file's exports.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <vector>

// the half-normal quantile table of the fused generator lives in this translation unit
#define ICDF_TABLE_DEFINE __device__ const unsigned long long g_icdf_table
#define ICDF32_TABLE_DEFINE __device__ const unsigned long long g_icdf32_table
#include "../../include/sdegpu.h"
#include "ensemble.cuh"
#include "integrals.cuh"
#include "laws.cuh"
#include "linear.cuh"

namespace sdegpu {
int launch_model_0(const RunArgs &, const LaunchCfg &, int *);
int launch_model_1(const RunArgs &, const LaunchCfg &, int *);
int launch_model_2(const RunArgs &, const LaunchCfg &, int *);
int launch_model_3(const RunArgs &, const LaunchCfg &, int *);
int launch_model_4(const RunArgs &, const LaunchCfg &, int *);
int launch_model_5(const RunArgs &, const LaunchCfg &, int *);
int launch_model_7(const RunArgs &, const LaunchCfg &, int *);
int launch_model_8(const RunArgs &, const LaunchCfg &, int *);

}  // namespace sdegpu

using namespace sdegpu;

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(e_ == cudaErrorMemoryAllocation ? SDEGPU_ERR_NOMEM : SDEGPU_ERR_CUDA, "%s: %s", \
                        #call, cudaGetErrorString(e_));                                                 \
    } while (0)

// grow-only device scratch
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct sdegpu_ctx {
    int device = 0;
    cudaDeviceProp prop;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    DevBuf dtsq, partials, stats, in0, in1, out0, out1, out2, params, scratch, times, kill, coef;
    const double *icdf = nullptr, *icdf32 = nullptr;
    int max_grid = 0;
};

static const int MODEL_NX[9] = {1, 1, 2, 1, 1, 1, 0, 1, 1};      // [6] linear: nx is free
static const int MODEL_NP[9] = {3, 3, 2, 2, 2, 1, 0, 0, 1};

extern "C" {

int sdegpu_abi_version(void) { return SDEGPU_ABI_VERSION; }
const char *sdegpu_last_error(void) { return g_err; }

int sdegpu_create(int device, sdegpu_ctx **out)
{
    if (!out) return fail(SDEGPU_ERR_ARG, "sdegpu_create: ctx is NULL");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(SDEGPU_ERR_NODEVICE, "sdegpu_create: no CUDA device visible");
    }
    if (device < 0 || device >= ndev) return fail(SDEGPU_ERR_ARG, "sdegpu_create: device %d out of range [0,%d)", device, ndev);
    sdegpu_ctx *c = new (std::nothrow) sdegpu_ctx();
    if (!c) return fail(SDEGPU_ERR_NOMEM, "sdegpu_create: host allocation failed");
    c->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&c->prop, device);
    if (e == cudaSuccess && c->prop.major != 10) {
        int maj = c->prop.major, min = c->prop.minor;
        delete c;
        return fail(SDEGPU_ERR_NODEVICE, "sdegpu_create: device is sm_%d%d; libsdegpu.so is built for sm_100a only", maj, min);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    c->max_grid = c->prop.multiProcessorCount * 32;
    if (e == cudaSuccess) e = c->partials.reserve(sizeof(double) * (size_t)c->max_grid * 5 * MAX_NX);
    if (e == cudaSuccess) e = cudaGetSymbolAddress((void **)&c->icdf, g_icdf_table);
    if (e == cudaSuccess) e = cudaGetSymbolAddress((void **)&c->icdf32, g_icdf32_table);
    if (e != cudaSuccess) {
        delete c;
        return fail(SDEGPU_ERR_CUDA, "sdegpu_create: %s", cudaGetErrorString(e));
    }
    c->stream = c->own_stream;
    *out = c;
    return SDEGPU_OK;
}

int sdegpu_destroy(sdegpu_ctx *c)
{
    if (!c) return SDEGPU_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf *bufs[] = {&c->dtsq, &c->partials, &c->stats, &c->in0, &c->in1, &c->out0, &c->out1, &c->out2, &c->params, &c->scratch, &c->times, &c->kill, &c->coef};
    for (DevBuf *b : bufs) b->release();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return SDEGPU_OK;
}

int sdegpu_device_info(sdegpu_ctx *c, sdegpu_devinfo *info)
{
    if (!c || !info) return fail(SDEGPU_ERR_ARG, "sdegpu_device_info: NULL argument");
    memset(info, 0, sizeof *info);
    strncpy(info->name, c->prop.name, sizeof info->name - 1);
    info->sm_count = c->prop.multiProcessorCount;
    info->cc_major = c->prop.major;
    info->cc_minor = c->prop.minor;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
    info->clock_khz = khz;
    info->global_mem_bytes = c->prop.totalGlobalMem;
    return SDEGPU_OK;
}

int sdegpu_set_stream(sdegpu_ctx *c, void *s)
{
    if (!c) return fail(SDEGPU_ERR_ARG, "sdegpu_set_stream: ctx is NULL");
    c->stream = (cudaStream_t)s;          // NULL is CUDA's default stream, a valid choice
    return SDEGPU_OK;
}

int sdegpu_reset_stream(sdegpu_ctx *c)
{
    if (!c) return fail(SDEGPU_ERR_ARG, "sdegpu_reset_stream: ctx is NULL");
    c->stream = c->own_stream;
    return SDEGPU_OK;
}

int sdegpu_synchronize(sdegpu_ctx *c)
{
    if (!c) return fail(SDEGPU_ERR_ARG, "sdegpu_synchronize: ctx is NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return SDEGPU_OK;
}

}  // extern "C"

// sum block partials in block order (deterministic), optionally accumulating
__global__ void k_reduce_partials(const double *partials, int nblocks, int nvals, double *out, int accumulate)
{
    const int v = threadIdx.x;
    if (v >= nvals) return;
    double s = accumulate ? out[v] : 0.0;
    for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * nvals + v];
    out[v] = s;
}

// {dt_i, sqrt(dt_i)} table; first = times[0] prepends the rBM convention dt_0 = times[1] (R/sdetools.R:18)
static int upload_dtsq(sdegpu_ctx *c, const double *times, int nt, bool rbm_convention, const double2 **dev)
{
    const int n = rbm_convention ? nt : nt - 1;
    std::vector<double2> h((size_t)n);
    for (int i = 0; i < n; ++i) {
        const double dt = rbm_convention ? (i == 0 ? times[0] : times[i] - times[i - 1]) : times[i + 1] - times[i];
        h[i] = make_double2(dt, sqrt(dt));
    }
    CU(c->dtsq.reserve(sizeof(double2) * (size_t)n));
    CU(cudaMemcpyAsync(c->dtsq.p, h.data(), sizeof(double2) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));   // h goes out of scope
    *dev = (const double2 *)c->dtsq.p;
    return SDEGPU_OK;
}

extern "C" int sdegpu_run(sdegpu_ctx *c, const sdegpu_problem *pr, sdegpu_output *out)
{
    if (!c || !pr || !out) return fail(SDEGPU_ERR_ARG, "sdegpu_run: NULL argument");
    if (pr->struct_size != sizeof(sdegpu_problem) || out->struct_size != sizeof(sdegpu_output))
        return fail(SDEGPU_ERR_ARG, "sdegpu_run: struct_size mismatch (ABI %d)", SDEGPU_ABI_VERSION);
    if (pr->model < 0 || pr->model > SDEGPU_MODEL_LOGLOGISTIC) return fail(SDEGPU_ERR_ARG, "sdegpu_run: unknown model %d", pr->model);
    if (pr->scheme != SDEGPU_EULER && pr->scheme != SDEGPU_HEUN) return fail(SDEGPU_ERR_ARG, "sdegpu_run: unknown scheme %d", pr->scheme);
    if (pr->projection < 0 || pr->projection > SDEGPU_PROJ_RELU) return fail(SDEGPU_ERR_ARG, "sdegpu_run: unknown projection %d", pr->projection);
    if (pr->arith != SDEGPU_ARITH_STRICT && pr->arith != SDEGPU_ARITH_FAST) return fail(SDEGPU_ERR_ARG, "sdegpu_run: unknown arith %d", pr->arith);
    // dB <- apply(B,2,diff) collapses for nt = 2 in the reference; nt >= 2 is well defined here
    if (pr->nt < 2) return fail(SDEGPU_ERR_ARG, "sdegpu_run: need at least 2 time points, got %d", pr->nt);
    if (!pr->times || !pr->x0 || (!pr->params && pr->nparams > 0) || pr->nparams < 0)
        return fail(SDEGPU_ERR_ARG, "sdegpu_run: times/x0/params must not be NULL");
    if (pr->x0_stride != 0 && pr->x0_stride != pr->nx) return fail(SDEGPU_ERR_ARG, "sdegpu_run: x0_stride must be 0 or nx");
    const bool killed = pr->kill_kind != SDEGPU_KILL_NONE || pr->stop_kind != SDEGPU_STOP_NONE;
    if (pr->kill_kind < 0 || pr->kill_kind > SDEGPU_KILL_QUAD || pr->stop_kind < 0 || pr->stop_kind > SDEGPU_STOP_OUTSIDE)
        return fail(SDEGPU_ERR_ARG, "sdegpu_run: unknown kill/stop kind (%d, %d)", pr->kill_kind, pr->stop_kind);
    if (killed && (pr->kill_comp < 0 || pr->kill_comp >= pr->nx || pr->stop_comp < 0 || pr->stop_comp >= pr->nx))
        return fail(SDEGPU_ERR_ARG, "sdegpu_run: kill/stop component out of range");
    if (killed && pr->model == SDEGPU_MODEL_LINEAR)
        return fail(SDEGPU_ERR_ARG, "sdegpu_run: killing/stopping is not available for the linear model");
    if (out->path_integrals && (killed || out->X || pr->model == SDEGPU_MODEL_LINEAR))
        return fail(SDEGPU_ERR_ARG, "sdegpu_run: path_integrals needs a catalogue model without kill/stop rule and X = NULL");
    if ((out->tau || out->S_tau || out->S) && !killed) return fail(SDEGPU_ERR_ARG, "sdegpu_run: tau/S_tau/S outputs need a kill or stop rule");
    if (!pr->B && pr->arith == SDEGPU_ARITH_STRICT && !killed)
        return fail(SDEGPU_ERR_ARG, "sdegpu_run: STRICT arithmetic is defined for a supplied B; fused-generator runs are FAST");
    for (int i = 0; i + 1 < pr->nt; ++i)
        if (!(pr->times[i + 1] >= pr->times[i])) return fail(SDEGPU_ERR_ARG, "sdegpu_run: times must be non-decreasing (index %d)", i);
    if (out->hist_counts && (out->hist_nbins < 1 || out->hist_nbins > 8192 || !(out->hist_hi > out->hist_lo) ||
                             out->hist_component < 0 || out->hist_component >= pr->nx))
        return fail(SDEGPU_ERR_ARG, "sdegpu_run: bad histogram spec (nbins in [1,8192], hi > lo, component < nx)");
    const bool dev = (pr->flags & SDEGPU_DEVICE_PTRS) != 0;
    const bool accum = (pr->flags & SDEGPU_ACCUMULATE) != 0;
    CU(cudaSetDevice(c->device));
    if (pr->npaths == 0) {
        if (out->elapsed_ms) *out->elapsed_ms = 0.f;
        return SDEGPU_OK;
    }
    const int nx = pr->nx, nt = pr->nt;
    const bool linear = pr->model == SDEGPU_MODEL_LINEAR;
    if (!linear) {
        if (nx != MODEL_NX[pr->model] || pr->nB != 1)
            return fail(SDEGPU_ERR_ARG, "sdegpu_run: model %d has nx=%d, nB=1 (got nx=%d, nB=%d)", pr->model, MODEL_NX[pr->model], nx, pr->nB);
        if (pr->nparams != MODEL_NP[pr->model])
            return fail(SDEGPU_ERR_ARG, "sdegpu_run: model %d takes %d parameters, got %d", pr->model, MODEL_NP[pr->model], pr->nparams);
    } else {
        if (nx < 1 || pr->nB < 1 || pr->nparams != nx * nx + nx * pr->nB)
            return fail(SDEGPU_ERR_ARG, "sdegpu_run: linear model needs nparams = nx*nx + nx*nB");
    }

    const double2 *dtsq = nullptr;
    int rc = upload_dtsq(c, pr->times, nt, false, &dtsq);
    if (rc) return rc;

    const size_t nB = (size_t)pr->nB;
    const size_t bytesB = sizeof(double) * pr->npaths * nB * nt, bytesX = sizeof(double) * pr->npaths * nx * nt,
                 bytesXT = sizeof(double) * pr->npaths * nx;
    const double *dB = nullptr, *dx0 = nullptr;
    double *dX = nullptr, *dXT = nullptr;
    if (pr->B) {
        if (dev) dB = pr->B;
        else {
            CU(c->in0.reserve(bytesB));
            CU(cudaMemcpyAsync(c->in0.p, pr->B, bytesB, cudaMemcpyHostToDevice, c->stream));
            dB = (const double *)c->in0.p;
        }
    }
    if (pr->x0_stride) {
        if (dev) dx0 = pr->x0;
        else {
            CU(c->in1.reserve(bytesXT));
            CU(cudaMemcpyAsync(c->in1.p, pr->x0, bytesXT, cudaMemcpyHostToDevice, c->stream));
            dx0 = (const double *)c->in1.p;
        }
    }
    if (out->X) {
        if (dev) dX = out->X;
        else { CU(c->out0.reserve(bytesX)); dX = (double *)c->out0.p; }
    }
    if (out->XT) {
        if (dev) dXT = out->XT;
        else { CU(c->out1.reserve(bytesXT)); dXT = (double *)c->out1.p; }
    }
    // statistics scratch: [5*nx power sums][nbins+3 counts]
    const int nps = 5 * nx, nh = out->hist_counts ? out->hist_nbins + 3 : 0;
    double *dps = nullptr;
    unsigned long long *dh = nullptr;
    if (out->power_sums || out->hist_counts) {
        CU(c->stats.reserve(sizeof(double) * nps + sizeof(unsigned long long) * (size_t)(nh + 1)));
        if (out->power_sums) dps = dev ? out->power_sums : (double *)c->stats.p;
        if (out->hist_counts) dh = dev ? (unsigned long long *)out->hist_counts : (unsigned long long *)((double *)c->stats.p + nps);
        if (dh) {
            if (accum && !dev) CU(cudaMemcpyAsync(dh, out->hist_counts, sizeof(unsigned long long) * nh, cudaMemcpyHostToDevice, c->stream));
            else if (!accum) CU(cudaMemsetAsync(dh, 0, sizeof(unsigned long long) * nh, c->stream));
        }
        if (dps && accum && !dev) CU(cudaMemcpyAsync(dps, out->power_sums, sizeof(double) * nps, cudaMemcpyHostToDevice, c->stream));
    }

    int grid_used = 0, cuerr = 0;
    CU(cudaEventRecord(c->ev0, c->stream));
    if (!linear) {
        RunArgs a;
        memset(&a, 0, sizeof a);
        for (int i = 0; i < pr->nparams; ++i) a.P.p[i] = pr->params[i];
        a.dtsq = dtsq;
        a.icdf = c->icdf;
        a.icdf32 = c->icdf32;
        bool uniform = false;
        if (!pr->B && !killed && !out->path_integrals) {   // k_terminal, k_traj: folded step coefficients (FAST only)
            // A grid whose increments agree to the rounding of its own time stamps (seq(0, T, by = h) gives
            // t_i = i*h rounded, so diff(times) scatters by ~ulp(T)) is integrated as the uniform grid it
            // stands for, dt = (t_n - t_0)/n: one coefficient row as a kernel parameter instead of a table.
            const double tmax = fmax(fabs(pr->times[0]), fabs(pr->times[nt - 1]));
            const double dtbar = (pr->times[nt - 1] - pr->times[0]) / (double)(nt - 1), tol = 4.0 * 2.220446049250313e-16 * tmax;
            uniform = !(pr->flags & SDEGPU_EXACT_GRID) && dtbar > 0.0;
            for (int i = 0; i + 1 < nt && uniform; ++i) uniform = fabs((pr->times[i + 1] - pr->times[i]) - dtbar) <= tol;
            if (uniform) fast_step_coefficients(pr->model, pr->params, dtbar, sqrt(dtbar), a.ucoef);
            else {
                std::vector<double> hc((size_t)(nt - 1) * FAST_NCOEF);
                for (int i = 0; i + 1 < nt; ++i) {
                    const double dt = pr->times[i + 1] - pr->times[i];
                    fast_step_coefficients(pr->model, pr->params, dt, sqrt(dt), &hc[(size_t)i * FAST_NCOEF]);
                }
                CU(c->coef.reserve(sizeof(double) * hc.size()));
                CU(cudaMemcpyAsync(c->coef.p, hc.data(), sizeof(double) * hc.size(), cudaMemcpyHostToDevice, c->stream));
                CU(cudaStreamSynchronize(c->stream));
                a.coef = (const double *)c->coef.p;
            }
        }
        a.x0 = dx0;
        if (!pr->x0_stride) for (int j = 0; j < nx; ++j) a.x0s[j] = pr->x0[j];
        a.B = dB; a.X = dX; a.XT = dXT;
        a.partials = dps ? (double *)c->partials.p : nullptr;
        a.hist = dh;
        for (int j = 0; j < nx; ++j) a.center[j] = out->center ? out->center[j] : 0.0;
        a.hist_lo = out->hist_lo; a.hist_hi = out->hist_hi;
        a.hist_invw = dh ? (double)out->hist_nbins / (out->hist_hi - out->hist_lo) : 0.0;
        a.npaths = pr->npaths; a.path_offset = pr->path_offset; a.seed = make_philox_key(pr->seed);
        a.nt = nt; a.hist_nbins = out->hist_nbins; a.hist_comp = out->hist_component;
        if (killed) {
            a.kill_kind = pr->kill_kind; a.kill_comp = pr->kill_comp; a.kill_a = pr->kill_a; a.kill_b = pr->kill_b;
            a.stop_kind = pr->stop_kind; a.stop_comp = pr->stop_comp; a.stop_lo = pr->stop_lo; a.stop_hi = pr->stop_hi;
            a.S0 = pr->S0;
            CU(c->times.reserve(sizeof(double) * nt));
            CU(cudaMemcpyAsync(c->times.p, pr->times, sizeof(double) * nt, cudaMemcpyHostToDevice, c->stream));
            a.times = (const double *)c->times.p;
            if (dev) { a.Stau = pr->Stau; a.tau = out->tau; a.S_tau = out->S_tau; a.Spath = out->S; }
            else {
                CU(c->kill.reserve(sizeof(double) * (3 + (out->S ? (size_t)nt : 0)) * pr->npaths));
                double *kb = (double *)c->kill.p;
                if (pr->Stau) {
                    CU(cudaMemcpyAsync(kb, pr->Stau, sizeof(double) * pr->npaths, cudaMemcpyHostToDevice, c->stream));
                    a.Stau = kb;
                }
                if (out->tau) a.tau = kb + pr->npaths;
                if (out->S_tau) a.S_tau = kb + 2 * pr->npaths;
                if (out->S) a.Spath = kb + 3 * pr->npaths;
            }
        }
        const size_t bytesF = sizeof(double) * 4 * (size_t)nx * pr->npaths;
        if (out->path_integrals) {
            if (dev) a.func = out->path_integrals;
            else { CU(c->out2.reserve(bytesF)); a.func = (double *)c->out2.p; }
        }
        LaunchCfg cfg;
        cfg.killed = killed; cfg.want_func = out->path_integrals != nullptr; cfg.uniform = uniform;
        cfg.scheme = pr->scheme; cfg.proj = pr->projection; cfg.arith = pr->B ? pr->arith : SDEGPU_ARITH_FAST;
        cfg.src_b = pr->B != nullptr; cfg.want_x = out->X != nullptr;
        cfg.grid = 0; cfg.sm_count = c->prop.multiProcessorCount; cfg.stream = c->stream;
        switch (pr->model) {
        case 0: cuerr = launch_model_0(a, cfg, &grid_used); break;
        case 1: cuerr = launch_model_1(a, cfg, &grid_used); break;
        case 2: cuerr = launch_model_2(a, cfg, &grid_used); break;
        case 3: cuerr = launch_model_3(a, cfg, &grid_used); break;
        case 4: cuerr = launch_model_4(a, cfg, &grid_used); break;
        case 5: cuerr = launch_model_5(a, cfg, &grid_used); break;
        case 7: cuerr = launch_model_7(a, cfg, &grid_used); break;
        default: cuerr = launch_model_8(a, cfg, &grid_used); break;
        }
        if (cuerr) return fail(SDEGPU_ERR_CUDA, "sdegpu_run: kernel launch: %s", cudaGetErrorString((cudaError_t)cuerr));
        if (grid_used > c->max_grid) return fail(SDEGPU_ERR_CUDA, "sdegpu_run: grid %d exceeds partials capacity", grid_used);
        if (dps) {
            k_reduce_partials<<<1, 32, 0, c->stream>>>((const double *)c->partials.p, grid_used, nps, dps, accum ? 1 : 0);
            CU(cudaGetLastError());
        }
    } else {
        LinearArgs la;
        memset(&la, 0, sizeof la);
        CU(c->params.reserve(sizeof(double) * (size_t)pr->nparams));
        CU(cudaMemcpyAsync(c->params.p, pr->params, sizeof(double) * (size_t)pr->nparams, cudaMemcpyHostToDevice, c->stream));
        la.A = (const double *)c->params.p;
        la.G = la.A + (size_t)nx * nx;
        la.dtsq = dtsq; la.icdf = c->icdf; la.x0 = pr->x0_stride ? dx0 : nullptr;
        if (!pr->x0_stride) {
            CU(c->in1.reserve(sizeof(double) * nx));
            CU(cudaMemcpyAsync(c->in1.p, pr->x0, sizeof(double) * nx, cudaMemcpyHostToDevice, c->stream));
            la.x0_shared = (const double *)c->in1.p;
        }
        la.B = dB; la.X = dX; la.XT = dXT;
        la.npaths = pr->npaths; la.path_offset = pr->path_offset; la.seed = make_philox_key(pr->seed);
        la.nt = nt; la.d = nx; la.nB = pr->nB; la.scheme = pr->scheme; la.strict = pr->B && pr->arith == SDEGPU_ARITH_STRICT;
        if (out->hist_counts) return fail(SDEGPU_ERR_ARG, "sdegpu_run: histogram output is not available for the linear model");
        la.moments = dps; la.accumulate = accum ? 1 : 0;
        CU(c->out2.reserve(sizeof(double) * nx));
        std::vector<double> hc((size_t)nx, 0.0);
        if (out->center) for (int j = 0; j < nx; ++j) hc[j] = out->center[j];
        CU(cudaMemcpyAsync(c->out2.p, hc.data(), sizeof(double) * nx, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        la.center = (const double *)c->out2.p;
        const int mode = dmma_mode(la, pr->params + (size_t)nx * nx);
        const size_t dp = ((size_t)nx + 31) / 32 * 32;              // the tensor kernels pad to a multiple of 32 rows
        if (mode == 1) {
            CU(c->scratch.reserve(sizeof(double) * (dp * dp + dp)));
            cuerr = launch_linear_dmma(la, (double *)c->scratch.p, c->prop.multiProcessorCount, c->stream);
        } else if (mode == 2) {
            CU(c->scratch.reserve(sizeof(double) * (dp * dp + dp * 4 * (((size_t)pr->nB + 3) / 4))));
            cuerr = launch_linear_dmma_dense(la, (double *)c->scratch.p, c->prop.multiProcessorCount, c->stream);
        } else
            cuerr = launch_linear(la, c->prop.multiProcessorCount, c->stream);
        if (cuerr) return fail(SDEGPU_ERR_CUDA, "sdegpu_run: linear kernel launch: %s", cudaGetErrorString((cudaError_t)cuerr));
    }
    CU(cudaEventRecord(c->ev1, c->stream));

    if (!dev) {
        if (out->path_integrals)
            CU(cudaMemcpyAsync(out->path_integrals, c->out2.p, sizeof(double) * 4 * (size_t)nx * pr->npaths, cudaMemcpyDeviceToHost, c->stream));
        if (out->X) CU(cudaMemcpyAsync(out->X, dX, bytesX, cudaMemcpyDeviceToHost, c->stream));
        if (out->XT) CU(cudaMemcpyAsync(out->XT, dXT, bytesXT, cudaMemcpyDeviceToHost, c->stream));
        if (out->power_sums) CU(cudaMemcpyAsync(out->power_sums, dps, sizeof(double) * nps, cudaMemcpyDeviceToHost, c->stream));
        if (out->hist_counts) CU(cudaMemcpyAsync(out->hist_counts, dh, sizeof(unsigned long long) * nh, cudaMemcpyDeviceToHost, c->stream));
        if (out->tau) CU(cudaMemcpyAsync(out->tau, (double *)c->kill.p + pr->npaths, sizeof(double) * pr->npaths, cudaMemcpyDeviceToHost, c->stream));
        if (out->S_tau) CU(cudaMemcpyAsync(out->S_tau, (double *)c->kill.p + 2 * pr->npaths, sizeof(double) * pr->npaths, cudaMemcpyDeviceToHost, c->stream));
        if (out->S) CU(cudaMemcpyAsync(out->S, (double *)c->kill.p + 3 * pr->npaths, sizeof(double) * pr->npaths * nt, cudaMemcpyDeviceToHost, c->stream));
    }
    if (!dev || out->elapsed_ms) {
        CU(cudaStreamSynchronize(c->stream));
        if (out->elapsed_ms) CU(cudaEventElapsedTime(out->elapsed_ms, c->ev0, c->ev1));
    }
    return SDEGPU_OK;
}

// ---- integrals ----------------------------------------------------------------
static int run_integral(sdegpu_ctx *c, int mode, const double *a, const double *b, uint64_t n, uint64_t ncols,
                        double *l, double *r, double *cc, uint32_t flags, const char *who)
{
    if (!c || !a || !b) return fail(SDEGPU_ERR_ARG, "%s: NULL argument", who);
    if (n == 0 || ncols == 0) return SDEGPU_OK;
    CU(cudaSetDevice(c->device));
    const bool dev = (flags & SDEGPU_DEVICE_PTRS) != 0;
    const size_t bytes = sizeof(double) * n * ncols;
    IntegralArgs q;
    q.n = n; q.ncols = ncols;
    if (dev) { q.a = a; q.b = b; q.l = l; q.r = r; q.c = cc; }
    else {
        CU(c->in0.reserve(bytes));
        CU(cudaMemcpyAsync(c->in0.p, a, bytes, cudaMemcpyHostToDevice, c->stream));
        q.a = (const double *)c->in0.p;
        if (b == a) q.b = q.a;
        else {
            CU(c->in1.reserve(bytes));
            CU(cudaMemcpyAsync(c->in1.p, b, bytes, cudaMemcpyHostToDevice, c->stream));
            q.b = (const double *)c->in1.p;
        }
        q.l = q.r = q.c = nullptr;
        if (l) { CU(c->out0.reserve(bytes)); q.l = (double *)c->out0.p; }
        if (r) { CU(c->out1.reserve(bytes)); q.r = (double *)c->out1.p; }
        if (cc) { CU(c->out2.reserve(bytes)); q.c = (double *)c->out2.p; }
    }
    if (n == 1) {                                               // c(0): nothing to integrate
        if (q.l) CU(cudaMemsetAsync(q.l, 0, bytes, c->stream));
        if (q.r) CU(cudaMemsetAsync(q.r, 0, bytes, c->stream));
        if (q.c) CU(cudaMemsetAsync(q.c, 0, bytes, c->stream));
    } else {
        int e = launch_integral(mode, q, c->prop.multiProcessorCount, c->stream);
        if (e) return fail(SDEGPU_ERR_CUDA, "%s: kernel launch: %s", who, cudaGetErrorString((cudaError_t)e));
    }
    if (!dev) {
        if (l) CU(cudaMemcpyAsync(l, q.l, bytes, cudaMemcpyDeviceToHost, c->stream));
        if (r) CU(cudaMemcpyAsync(r, q.r, bytes, cudaMemcpyDeviceToHost, c->stream));
        if (cc) CU(cudaMemcpyAsync(cc, q.c, bytes, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return SDEGPU_OK;
}

// ou.parameters (R/sdetools.R:706-712) / cir.parameters (:606-617) folded into the kernels' constants
static int fold_law(int32_t law, const double params[3], double t, int32_t stratonovich, LawArgs *q)
{
    if (law != SDEGPU_LAW_OU && law != SDEGPU_LAW_CIR) return fail(SDEGPU_ERR_ARG, "sdegpu law: unknown law %d", law);
    if (!(t >= 0)) return fail(SDEGPU_ERR_ARG, "sdegpu law: t must be >= 0");
    const double lambda = params[0], xi = params[1], s = params[2];
    if (!(s >= 0)) return fail(SDEGPU_ERR_ARG, "sdegpu law: noise intensity must be >= 0");
    if (law == SDEGPU_LAW_CIR && (!(lambda > 0) || !(s > 0) || !(t > 0)))
        return fail(SDEGPU_ERR_ARG, "sdegpu law: CIR needs lambda > 0, gamma > 0, t > 0");
    memset(q, 0, sizeof *q);
    q->law = law;
    q->decay = exp(-lambda * t);                                 // :708 / :613-614
    q->xi = xi;
    q->sd = lambda == 0 ? s * sqrt(t) : s * sqrt((1 - exp(-2 * lambda * t)) / 2 / lambda);           // :709
    const double xi_ito = stratonovich ? xi + s * s / 4 / lambda : xi;                                // :610
    q->two_c = law == SDEGPU_LAW_CIR ? 2 * (2 * lambda / (s * s) / (1 - q->decay)) : 1.0;           // :613
    q->df = law == SDEGPU_LAW_CIR ? 4 * lambda * xi_ito / (s * s) : 0.0;                              // :615
    if (law == SDEGPU_LAW_CIR && !(q->df >= 0)) return fail(SDEGPU_ERR_ARG, "sdegpu law: CIR needs xi >= 0");
    return SDEGPU_OK;
}

extern "C" {

int sdegpu_law_eval(sdegpu_ctx *c, int32_t law, int32_t what, const double params[3], double t, int32_t stratonovich,
                    const double *x0, int64_t x0_stride, const double *x, uint64_t n, double *out, uint32_t flags)
{
    if (!c || !params || !x0 || !x || !out) return fail(SDEGPU_ERR_ARG, "sdegpu_law_eval: NULL argument");
    if (what != 0 && what != 1) return fail(SDEGPU_ERR_ARG, "sdegpu_law_eval: what must be 0 (density) or 1 (distribution)");
    if (x0_stride < 0) return fail(SDEGPU_ERR_ARG, "sdegpu_law_eval: x0_stride must be >= 0");
    LawArgs q;
    int rc = fold_law(law, params, t, stratonovich, &q);
    if (rc) return rc;
    if (law == SDEGPU_LAW_OU && !(q.sd > 0)) return fail(SDEGPU_ERR_ARG, "sdegpu_law_eval: degenerate OU law (sd = 0)");
    if (n == 0) return SDEGPU_OK;
    CU(cudaSetDevice(c->device));
    const bool dev = (flags & SDEGPU_DEVICE_PTRS) != 0;
    q.n = n; q.x0_stride = (uint64_t)x0_stride;
    const double *dx = x;
    double *dout = out;
    if (x0_stride == 0) {
        if (dev) {
            CU(cudaMemcpyAsync(&q.x0s, x0, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
        } else q.x0s = x0[0];
    } else if (dev) q.x0 = x0;
    else {
        const size_t bytes = sizeof(double) * ((n - 1) * (size_t)x0_stride + 1);
        CU(c->in0.reserve(bytes));
        CU(cudaMemcpyAsync(c->in0.p, x0, bytes, cudaMemcpyHostToDevice, c->stream));
        q.x0 = (const double *)c->in0.p;
    }
    if (!dev) {
        CU(c->in1.reserve(sizeof(double) * n));
        CU(cudaMemcpyAsync(c->in1.p, x, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
        dx = (const double *)c->in1.p;
        CU(c->out0.reserve(sizeof(double) * n));
        dout = (double *)c->out0.p;
    }
    int e = launch_law_eval(q, what, dx, dout, c->prop.multiProcessorCount, c->stream);
    if (e) return fail(SDEGPU_ERR_CUDA, "sdegpu_law_eval: kernel launch: %s", cudaGetErrorString((cudaError_t)e));
    if (!dev) {
        CU(cudaMemcpyAsync(out, dout, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return SDEGPU_OK;
}

}  // extern "C"

extern "C" {

int sdegpu_stochint(sdegpu_ctx *c, const double *f, const double *g, uint64_t n, uint64_t ncols, double *l, double *r,
                    double *cc, uint32_t flags)
{
    return run_integral(c, 0, f, g, n, ncols, l, r, cc, flags, "sdegpu_stochint");
}

int sdegpu_itointegral(sdegpu_ctx *c, const double *G, const double *B, uint64_t n, uint64_t ncols, double *out, uint32_t flags)
{
    return run_integral(c, 0, G, B, n, ncols, out, nullptr, nullptr, flags, "sdegpu_itointegral");
}

int sdegpu_qv(sdegpu_ctx *c, const double *X, uint64_t n, uint64_t ncols, double *out, uint32_t flags)
{
    return run_integral(c, 1, X, X, n, ncols, out, nullptr, nullptr, flags, "sdegpu_qv");
}

int sdegpu_crossvar(sdegpu_ctx *c, const double *X, const double *Y, uint64_t n, uint64_t ncols, double *out, uint32_t flags)
{
    return run_integral(c, 1, X, Y, n, ncols, out, nullptr, nullptr, flags, "sdegpu_crossvar");
}

int sdegpu_rvbm(sdegpu_ctx *c, const double *times, int32_t nt, uint64_t ncols, double sigma, double B0, double u,
                uint64_t seed, uint64_t path_offset, double *B, uint32_t flags)
{
    if (!c || !times || !B) return fail(SDEGPU_ERR_ARG, "sdegpu_rvbm: NULL argument");
    if (nt < 1) return fail(SDEGPU_ERR_ARG, "sdegpu_rvbm: nt must be >= 1");
    if (ncols == 0) return SDEGPU_OK;
    if (times[0] < 0) return fail(SDEGPU_ERR_ARG, "sdegpu_rvbm: times[0] must be >= 0 (first increment spans [0,times[0]])");
    for (int i = 0; i + 1 < nt; ++i)
        if (!(times[i + 1] >= times[i])) return fail(SDEGPU_ERR_ARG, "sdegpu_rvbm: times must be non-decreasing (index %d)", i);
    CU(cudaSetDevice(c->device));
    const bool dev = (flags & SDEGPU_DEVICE_PTRS) != 0;
    const double2 *dtsq = nullptr;
    int rc = upload_dtsq(c, times, nt, true, &dtsq);
    if (rc) return rc;
    const size_t bytes = sizeof(double) * ncols * (size_t)nt;
    BMArgs q;
    q.dtsq = dtsq; q.icdf = c->icdf; q.sigma = sigma; q.B0 = B0; q.u = u; q.ncols = ncols; q.path_offset = path_offset; q.seed = make_philox_key(seed); q.nt = nt;
    if (dev) q.B = B;
    else { CU(c->out0.reserve(bytes)); q.B = (double *)c->out0.p; }
    int e = launch_rvbm(q, c->prop.multiProcessorCount, c->stream);
    if (e) return fail(SDEGPU_ERR_CUDA, "sdegpu_rvbm: kernel launch: %s", cudaGetErrorString((cudaError_t)e));
    if (!dev) {
        CU(cudaMemcpyAsync(B, q.B, bytes, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return SDEGPU_OK;
}

int sdegpu_bm_functionals(sdegpu_ctx *c, const double *times, int32_t nt, uint64_t ncols, double sigma, double B0, double u,
                           double level, uint64_t seed, uint64_t path_offset, double *vmax, double *vmin, double *tmax,
                           double *tmin, double *thit, uint32_t flags)
{
    if (!c || !times) return fail(SDEGPU_ERR_ARG, "sdegpu_bm_functionals: NULL argument");
    if (nt < 1) return fail(SDEGPU_ERR_ARG, "sdegpu_bm_functionals: nt must be >= 1");
    if (ncols == 0) return SDEGPU_OK;
    if (times[0] < 0) return fail(SDEGPU_ERR_ARG, "sdegpu_bm_functionals: times[0] must be >= 0 (first increment spans [0,times[0]])");
    for (int i = 0; i + 1 < nt; ++i)
        if (!(times[i + 1] >= times[i])) return fail(SDEGPU_ERR_ARG, "sdegpu_bm_functionals: times must be non-decreasing (index %d)", i);
    CU(cudaSetDevice(c->device));
    const bool dev = (flags & SDEGPU_DEVICE_PTRS) != 0;
    const double2 *dtsq = nullptr;
    int rc = upload_dtsq(c, times, nt, true, &dtsq);
    if (rc) return rc;
    CU(c->times.reserve(sizeof(double) * (size_t)nt));
    CU(cudaMemcpyAsync(c->times.p, times, sizeof(double) * (size_t)nt, cudaMemcpyHostToDevice, c->stream));
    BMFuncArgs q;
    q.bm.dtsq = dtsq; q.bm.icdf = c->icdf; q.bm.B = nullptr; q.bm.sigma = sigma; q.bm.B0 = B0; q.bm.u = u; q.bm.ncols = ncols;
    q.bm.path_offset = path_offset; q.bm.seed = make_philox_key(seed); q.bm.nt = nt;
    q.level = level; q.times = (const double *)c->times.p;
    double *host[5] = {vmax, vmin, tmax, tmin, thit};
    double **dst[5] = {&q.vmax, &q.vmin, &q.tmax, &q.tmin, &q.thit};
    if (dev) for (int k = 0; k < 5; ++k) *dst[k] = host[k];
    else {
        CU(c->out0.reserve(sizeof(double) * 5 * ncols));
        for (int k = 0; k < 5; ++k) *dst[k] = host[k] ? (double *)c->out0.p + (size_t)k * ncols : nullptr;
    }
    int e = launch_bm_functionals(q, c->prop.multiProcessorCount, c->stream);
    if (e) return fail(SDEGPU_ERR_CUDA, "sdegpu_bm_functionals: kernel launch: %s", cudaGetErrorString((cudaError_t)e));
    if (!dev)
        for (int k = 0; k < 5; ++k)
            if (host[k]) CU(cudaMemcpyAsync(host[k], *dst[k], sizeof(double) * ncols, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));                        // the staged `times` is read by the kernel
    return SDEGPU_OK;
}

int sdegpu_rlaw(sdegpu_ctx *c, int32_t law, const double params[3], double t, int32_t stratonovich, const double *x0,
                int64_t x0_stride, uint64_t n, int32_t nsteps, uint64_t seed, uint64_t path_offset, double *out, double *path,
                uint32_t flags)
{
    if (!c || !params || !x0 || !out) return fail(SDEGPU_ERR_ARG, "sdegpu_rlaw: NULL argument");
    if (law != SDEGPU_LAW_OU && law != SDEGPU_LAW_CIR) return fail(SDEGPU_ERR_ARG, "sdegpu_rlaw: unknown law %d", law);
    if (nsteps < 0 || x0_stride < 0) return fail(SDEGPU_ERR_ARG, "sdegpu_rlaw: nsteps and x0_stride must be >= 0");
    if (!(t >= 0)) return fail(SDEGPU_ERR_ARG, "sdegpu_rlaw: t must be >= 0");
    LawArgs q;
    int rc = fold_law(law, params, t, stratonovich, &q);
    if (rc) return rc;
    if (n == 0) return SDEGPU_OK;
    CU(cudaSetDevice(c->device));
    const bool dev = (flags & SDEGPU_DEVICE_PTRS) != 0;
    q.icdf = c->icdf; q.n = n; q.path_offset = path_offset; q.seed = make_philox_key(seed); q.nsteps = nsteps;
    const size_t out_bytes = sizeof(double) * n, path_bytes = path ? sizeof(double) * n * ((size_t)nsteps + 1) : 0;
    q.x0 = nullptr; q.x0s = 0.0; q.x0_stride = (uint64_t)x0_stride;
    if (x0_stride == 0) {
        if (dev) {
            CU(cudaMemcpyAsync(&q.x0s, x0, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
        } else q.x0s = x0[0];
    } else if (dev) q.x0 = x0;
    else {
        const size_t bytes = sizeof(double) * ((n - 1) * (size_t)x0_stride + 1);
        CU(c->in0.reserve(bytes));
        CU(cudaMemcpyAsync(c->in0.p, x0, bytes, cudaMemcpyHostToDevice, c->stream));
        q.x0 = (const double *)c->in0.p;
    }
    if (dev) { q.out = out; q.path = path; }
    else {
        CU(c->out0.reserve(out_bytes));
        q.out = (double *)c->out0.p;
        q.path = nullptr;
        if (path) { CU(c->out1.reserve(path_bytes)); q.path = (double *)c->out1.p; }
    }
    int e = launch_rlaw(q, c->prop.multiProcessorCount, c->stream);
    if (e) return fail(SDEGPU_ERR_CUDA, "sdegpu_rlaw: kernel launch: %s", cudaGetErrorString((cudaError_t)e));
    if (!dev) {
        CU(cudaMemcpyAsync(out, q.out, out_bytes, cudaMemcpyDeviceToHost, c->stream));
        if (path) CU(cudaMemcpyAsync(path, q.path, path_bytes, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return SDEGPU_OK;
}

}  // extern "C"

// ---- generator introspection ----------------------------------------------------
__global__ void k_philox_kat(uint4 ctr, uint2 key, uint32_t *out)
{
    uint32_t c0 = ctr.x, c1 = ctr.y, c2 = ctr.z, c3 = ctr.w;
    PhiloxKey pk;
    uint32_t k0 = key.x, k1 = key.y;
    for (int r = 0; r < 10; ++r) { pk.k[2 * r] = k0; pk.k[2 * r + 1] = k1; k0 += PHILOX_W0; k1 += PHILOX_W1; }
    philox4x32_10(c0, c1, c2, c3, pk);
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void k_normals(const double *icdf, const PhiloxKey seed, uint64_t path_offset, uint64_t npaths, uint32_t channel,
                          int nsteps, double *z)
{
    extern __shared__ __align__(16) double sh_tab[];
    icdf_stage(sh_tab, icdf, threadIdx.x, blockDim.x);
    __syncthreads();
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npaths) return;
    for (int i = 0; i < nsteps; i += 4) {
        double zq[4];
        normal_quad(seed, path_offset + p, (uint32_t)(i >> 2), channel, sh_tab, zq);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i + k < nsteps) z[(size_t)(i + k) * npaths + p] = zq[k];
    }
}

// ---- roofline denominators --------------------------------------------------------
template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma_peak(int iters, double seed, double *sink)
{
    double acc[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) acc[k] = seed + k + threadIdx.x * 1e-9;
    const double a = 1.0000000001, b = 1e-12;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) acc[k] = fma(acc[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += acc[k];
    if (s == 12345.678) sink[0] = s;      // never true; keeps the chains alive
}

__global__ void __launch_bounds__(256) k_copy(const double2 *__restrict__ src, double2 *__restrict__ dst, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

extern "C" {

int sdegpu_philox4x32_10(sdegpu_ctx *c, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    if (!c || !ctr || !key || !out) return fail(SDEGPU_ERR_ARG, "sdegpu_philox4x32_10: NULL argument");
    CU(cudaSetDevice(c->device));
    CU(c->stats.reserve(64));
    k_philox_kat<<<1, 1, 0, c->stream>>>(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), make_uint2(key[0], key[1]), (uint32_t *)c->stats.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, c->stats.p, 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SDEGPU_OK;
}

int sdegpu_normals(sdegpu_ctx *c, uint64_t seed, uint64_t path_offset, uint64_t npaths, uint32_t channel, int32_t nsteps, double *z)
{
    if (!c || !z || nsteps < 0) return fail(SDEGPU_ERR_ARG, "sdegpu_normals: bad argument");
    if (npaths == 0 || nsteps == 0) return SDEGPU_OK;
    CU(cudaSetDevice(c->device));
    const size_t bytes = sizeof(double) * npaths * (size_t)nsteps;
    CU(c->out0.reserve(bytes));
    cudaFuncSetAttribute(k_normals, cudaFuncAttributeMaxDynamicSharedMemorySize, ICDF_TABLE_BYTES);
    k_normals<<<(unsigned)((npaths + 127) / 128), 128, ICDF_TABLE_BYTES, c->stream>>>(c->icdf, make_philox_key(seed), path_offset, npaths, channel,
                                                                                     nsteps, (double *)c->out0.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(z, c->out0.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SDEGPU_OK;
}

int sdegpu_measure_fp64_peak(sdegpu_ctx *c, double *tflops)
{
    if (!c || !tflops) return fail(SDEGPU_ERR_ARG, "sdegpu_measure_fp64_peak: NULL argument");
    CU(cudaSetDevice(c->device));
    CU(c->stats.reserve(64));
    constexpr int CH = 8;
    const int iters = 1 << 16, blocks = c->prop.multiProcessorCount * 8;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CU(cudaEventRecord(c->ev0, c->stream));
        k_dfma_peak<CH><<<blocks, 256, 0, c->stream>>>(iters, 0.5, (double *)c->stats.p);
        CU(cudaEventRecord(c->ev1, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        const double tf = 2.0 * CH * (double)iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    *tflops = best;
    return SDEGPU_OK;
}

int sdegpu_measure_hbm_peak(sdegpu_ctx *c, double *gbs)
{
    if (!c || !gbs) return fail(SDEGPU_ERR_ARG, "sdegpu_measure_hbm_peak: NULL argument");
    CU(cudaSetDevice(c->device));
    const size_t bytes = (size_t)2 << 30;
    void *a = nullptr, *b = nullptr;
    CU(cudaMalloc(&a, bytes));
    if (cudaMalloc(&b, bytes) != cudaSuccess) { cudaFree(a); return fail(SDEGPU_ERR_NOMEM, "sdegpu_measure_hbm_peak: cudaMalloc"); }
    cudaMemsetAsync(a, 1, bytes, c->stream);
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(c->ev0, c->stream);
        k_copy<<<c->prop.multiProcessorCount * 16, 256, 0, c->stream>>>((const double2 *)a, (double2 *)b, bytes / 16);
        cudaEventRecord(c->ev1, c->stream);
        cudaStreamSynchronize(c->stream);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        const double g = 2.0 * bytes / (ms * 1e-3) / 1e9;
        if (rep > 0 && g > best) best = g;
    }
    cudaFree(a); cudaFree(b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SDEGPU_ERR_CUDA, "sdegpu_measure_hbm_peak: %s", cudaGetErrorString(e));
    *gbs = best;
    return SDEGPU_OK;
}

}  // extern "C"
