// C ABI of libsdegpu.so (include/sdegpu.h).  No R / torch / C++ types cross this file's exports.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <dlfcn.h>
#include <emmintrin.h>                 // SSE2 streaming stores for the staging copies (x86-64 baseline)
#include <nccl.h>                      // types only: the library is loaded with dlopen when a multi-device context needs it
#include <stdlib.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

// the half-normal quantile table of the fused generator lives in this translation unit
#define ICDF_TABLE_DEFINE __device__ const unsigned long long g_icdf_table
#define ICDF32_TABLE_DEFINE __device__ const unsigned long long g_icdf32_table
#include "../../include/sdegpu.h"
#include "ensemble.cuh"
#include "integrals.cuh"
#include "laws.cuh"
#include "linear.cuh"
#include "linlaws.cuh"

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
// grow-only pinned host staging
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct NcclApi;                                                  // function table of a libnccl loaded at run time (below)

struct sdegpu_ctx {
    int device = 0;
    cudaDeviceProp prop;
    cudaStream_t own_stream = nullptr, stream = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k0[2] = {nullptr, nullptr}, ev_k1[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    cudaEvent_t ev_pin_in[2] = {nullptr, nullptr}, ev_pin_out[2] = {nullptr, nullptr};      // a pinned piece has left / arrived
    unsigned pin_in_turn = 0, pin_out_turn = 0;
    DevBuf dtsq, partials, stats, in0, in1, out0, out1, out2, params, scratch, times, kill, coef, tdep, lin_x0, lin_center, lin_partials;
    DevBuf pipe_in[2], pipe_out[2];                              // device side of the chunk pipeline
    PinBuf pin_in[2], pin_out[2];                                // its pinned host side
    const double *icdf = nullptr, *icdf32 = nullptr;
    int max_grid = 0;
    // several GPUs in one process: this context is devices[0]; peers are devices[1..]
    std::vector<sdegpu_ctx *> peers;
    sdegpu_ctx *leader = nullptr;                                // set in a peer
    void *nccl_comm = nullptr;                                   // ncclComm_t of this device (leader and peers)
    bool nccl_tried = false;
    const char *last_reduction = "none";
    // interruption and chunking of host-pointer runs
    sdegpu_interrupt_fn intr = nullptr;
    void *intr_user = nullptr;
    std::atomic<int> abort_flag{0};
    uint64_t chunk_paths = 0;
};

static const int MODEL_NX[9] = {1, 1, 2, 1, 1, 1, 0, 1, 1};      // [6] linear: nx is free
static const int MODEL_NP[9] = {3, 3, 2, 2, 2, 1, 0, 0, 1};

static void copy_pool_acquire();
static void copy_pool_release();

static void release_device(sdegpu_ctx *c)
{
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    DevBuf *bufs[] = {&c->dtsq, &c->partials, &c->stats, &c->in0, &c->in1, &c->out0, &c->out1, &c->out2, &c->params, &c->scratch, &c->times,
                      &c->kill, &c->coef, &c->tdep, &c->lin_x0, &c->lin_center, &c->lin_partials, &c->pipe_in[0], &c->pipe_in[1], &c->pipe_out[0], &c->pipe_out[1]};
    for (DevBuf *b : bufs) b->release();
    for (int k = 0; k < 2; ++k) {
        c->pin_in[k].release(); c->pin_out[k].release();
        cudaEvent_t *evs[] = {&c->ev_in[k], &c->ev_k0[k], &c->ev_k1[k], &c->ev_out[k], &c->ev_pin_in[k], &c->ev_pin_out[k]};
        for (cudaEvent_t *e : evs) if (*e) cudaEventDestroy(*e);
    }
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
}

static int create_device(int device, sdegpu_ctx **out)
{
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
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
        e = cudaEventCreateWithFlags(&c->ev_in[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreate(&c->ev_k0[k]);
        if (e == cudaSuccess) e = cudaEventCreate(&c->ev_k1[k]);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_out[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_pin_in[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_pin_out[k], cudaEventDisableTiming);
    }
    c->max_grid = c->prop.multiProcessorCount * 32;
    if (e == cudaSuccess) e = c->partials.reserve(sizeof(double) * (size_t)c->max_grid * 5 * MAX_NX);
    if (e == cudaSuccess) e = cudaGetSymbolAddress((void **)&c->icdf, g_icdf_table);
    if (e == cudaSuccess) e = cudaGetSymbolAddress((void **)&c->icdf32, g_icdf32_table);
    if (e != cudaSuccess) {
        release_device(c);
        delete c;
        return fail(SDEGPU_ERR_CUDA, "sdegpu_create: %s", cudaGetErrorString(e));
    }
    c->stream = c->own_stream;
    *out = c;
    return SDEGPU_OK;
}

// ---- NCCL, loaded at run time (libsdegpu.so itself does not link against it) ------------------------------------
struct NcclApi {
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
static NcclApi &nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    if (const char *e = getenv("SDEGPU_NO_NCCL")) if (atoi(e) != 0) return api;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return api;
    api.CommInitAll = (decltype(api.CommInitAll))dlsym(h, "ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
    api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
    api.ok = api.CommInitAll && api.CommDestroy && api.AllReduce && api.GroupStart && api.GroupEnd;
    return api;
}

extern "C" {

int sdegpu_abi_version(void) { return SDEGPU_ABI_VERSION; }
const char *sdegpu_last_error(void) { return g_err; }

int sdegpu_create_multi(const int *devices, int ndev, sdegpu_ctx **out)
{
    if (!out) return fail(SDEGPU_ERR_ARG, "sdegpu_create: ctx is NULL");
    *out = nullptr;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) {
        cudaGetLastError();
        return fail(SDEGPU_ERR_NODEVICE, "sdegpu_create: no CUDA device visible");
    }
    if (ndev < 1 || ndev > 64) return fail(SDEGPU_ERR_ARG, "sdegpu_create_multi: ndev must be in [1,64], got %d", ndev);
    std::vector<int> ids((size_t)ndev);
    for (int k = 0; k < ndev; ++k) {
        ids[k] = devices ? devices[k] : k;
        if (ids[k] < 0 || ids[k] >= have) return fail(SDEGPU_ERR_ARG, "sdegpu_create: device %d out of range [0,%d)", ids[k], have);
        for (int m = 0; m < k; ++m)
            if (ids[m] == ids[k]) return fail(SDEGPU_ERR_ARG, "sdegpu_create_multi: device %d listed twice", ids[k]);
    }
    sdegpu_ctx *lead = nullptr;
    int rc = create_device(ids[0], &lead);
    if (rc) return rc;
    for (int k = 1; k < ndev; ++k) {
        sdegpu_ctx *peer = nullptr;
        rc = create_device(ids[k], &peer);
        if (rc) { sdegpu_destroy(lead); return rc; }
        peer->leader = lead;
        lead->peers.push_back(peer);
    }
    cudaSetDevice(ids[0]);
    copy_pool_acquire();
    *out = lead;
    return SDEGPU_OK;
}

int sdegpu_create(int device, sdegpu_ctx **out) { return sdegpu_create_multi(&device, 1, out); }

int sdegpu_destroy(sdegpu_ctx *c)
{
    if (!c) return SDEGPU_OK;
    NcclApi &n = nccl_api();
    for (sdegpu_ctx *p : c->peers) {
        if (p->nccl_comm && n.ok) n.CommDestroy((ncclComm_t)p->nccl_comm);
        release_device(p);
        delete p;
    }
    if (c->nccl_comm && n.ok) n.CommDestroy((ncclComm_t)c->nccl_comm);
    release_device(c);
    delete c;
    copy_pool_release();
    return SDEGPU_OK;
}

int sdegpu_device_count(sdegpu_ctx *c) { return c ? 1 + (int)c->peers.size() : 0; }
const char *sdegpu_last_reduction(sdegpu_ctx *c) { return c ? c->last_reduction : "none"; }

int sdegpu_set_interrupt(sdegpu_ctx *c, sdegpu_interrupt_fn fn, void *user)
{
    if (!c) return fail(SDEGPU_ERR_ARG, "sdegpu_set_interrupt: ctx is NULL");
    c->intr = fn; c->intr_user = user;
    return SDEGPU_OK;
}

int sdegpu_set_chunk_paths(sdegpu_ctx *c, uint64_t paths)
{
    if (!c) return fail(SDEGPU_ERR_ARG, "sdegpu_set_chunk_paths: ctx is NULL");
    c->chunk_paths = paths;
    for (sdegpu_ctx *p : c->peers) p->chunk_paths = paths;
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
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));  // kernels enqueued on the old stream may still read the context's tables
    c->stream = (cudaStream_t)s;          // NULL is CUDA's default stream, a valid choice
    return SDEGPU_OK;
}

int sdegpu_reset_stream(sdegpu_ctx *c)
{
    if (!c) return fail(SDEGPU_ERR_ARG, "sdegpu_reset_stream: ctx is NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
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

// ---- sdegpu_run ---------------------------------------------------------------------------------------------------
// validate()  host-only checks of a problem/output pair
// prepare()   uploads the tables that depend on (times, params) to one device: one stream synchronisation
// enqueue()   launches the kernels of one path range on c->stream; every per-path pointer is a device pointer
// run_device_ptrs() / run_host() / run_multi()  the three ways a call reaches enqueue()

static int validate(const sdegpu_problem *pr, const sdegpu_output *out)
{
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
    const bool tdep = pr->drift_forcing || pr->diffusion_scale;
    const bool linear = pr->model == SDEGPU_MODEL_LINEAR;
    if (pr->kill_kind < 0 || pr->kill_kind > SDEGPU_KILL_QUAD || pr->stop_kind < 0 || pr->stop_kind > SDEGPU_STOP_OUTSIDE)
        return fail(SDEGPU_ERR_ARG, "sdegpu_run: unknown kill/stop kind (%d, %d)", pr->kill_kind, pr->stop_kind);
    if (killed && (pr->kill_comp < 0 || pr->kill_comp >= pr->nx || pr->stop_comp < 0 || pr->stop_comp >= pr->nx))
        return fail(SDEGPU_ERR_ARG, "sdegpu_run: kill/stop component out of range");
    if (killed && linear) return fail(SDEGPU_ERR_ARG, "sdegpu_run: killing/stopping is not available for the linear model");
    if ((out->path_integrals || out->model_integrals) && (killed || out->X || linear))
        return fail(SDEGPU_ERR_ARG, "sdegpu_run: path_integrals/model_integrals need a catalogue model without kill/stop rule and X = NULL");
    if ((out->tau || out->S_tau || out->S) && !killed) return fail(SDEGPU_ERR_ARG, "sdegpu_run: tau/S_tau/S outputs need a kill or stop rule");
    if (!pr->B && pr->arith == SDEGPU_ARITH_STRICT && !killed && !tdep)
        return fail(SDEGPU_ERR_ARG, "sdegpu_run: STRICT arithmetic is defined for a supplied B; fused-generator runs are FAST");
    if (linear && pr->diffusion_scale) return fail(SDEGPU_ERR_ARG, "sdegpu_run: diffusion_scale is not available for the linear model (drift_forcing is)");
    for (int i = 0; i + 1 < pr->nt; ++i)
        if (!(pr->times[i + 1] >= pr->times[i])) return fail(SDEGPU_ERR_ARG, "sdegpu_run: times must be non-decreasing (index %d)", i);
    if (out->hist_counts && (out->hist_nbins < 1 || out->hist_nbins > 8192 || !(out->hist_hi > out->hist_lo) ||
                             out->hist_component < 0 || out->hist_component >= pr->nx))
        return fail(SDEGPU_ERR_ARG, "sdegpu_run: bad histogram spec (nbins in [1,8192], hi > lo, component < nx)");
    if (!linear) {
        if (pr->nx != MODEL_NX[pr->model] || pr->nB != 1)
            return fail(SDEGPU_ERR_ARG, "sdegpu_run: model %d has nx=%d, nB=1 (got nx=%d, nB=%d)", pr->model, MODEL_NX[pr->model], pr->nx, pr->nB);
        if (pr->nparams != MODEL_NP[pr->model])
            return fail(SDEGPU_ERR_ARG, "sdegpu_run: model %d takes %d parameters, got %d", pr->model, MODEL_NP[pr->model], pr->nparams);
    } else {
        if (pr->nx < 1 || pr->nB < 1 || pr->nparams != pr->nx * pr->nx + pr->nx * pr->nB)
            return fail(SDEGPU_ERR_ARG, "sdegpu_run: linear model needs nparams = nx*nx + nx*nB");
        if (out->hist_counts) return fail(SDEGPU_ERR_ARG, "sdegpu_run: histogram output is not available for the linear model");
    }
    return SDEGPU_OK;
}

struct Prepared {
    const double2 *dtsq = nullptr;
    bool uniform = false, killed = false, tdep = false, linear = false;
    double ucoef[FAST_NCOEF] = {0, 0, 0, 0, 0, 0};
    const double *coef = nullptr, *times = nullptr, *forcing = nullptr, *gscale = nullptr;
    const double *linA = nullptr, *linG = nullptr, *lin_x0 = nullptr, *lin_center = nullptr;
};

// every per-path array of one path range, on the device
struct DevIO {
    const double *B = nullptr, *x0 = nullptr, *Stau = nullptr;
    double *X = nullptr, *XT = nullptr, *tau = nullptr, *S_tau = nullptr, *S = nullptr, *func = nullptr, *mfunc = nullptr;
    double *ps = nullptr;                                        // 5*nx power sums (device)
    unsigned long long *hist = nullptr;                          // nbins+3 counts (device)
    uint64_t npaths = 0, path_offset = 0;
};

static int prepare(sdegpu_ctx *c, const sdegpu_problem *pr, const sdegpu_output *out, Prepared *pp)
{
    const int nt = pr->nt, nx = pr->nx;
    pp->killed = pr->kill_kind != SDEGPU_KILL_NONE || pr->stop_kind != SDEGPU_STOP_NONE;
    pp->tdep = pr->drift_forcing || pr->diffusion_scale;
    pp->linear = pr->model == SDEGPU_MODEL_LINEAR;
    std::vector<double2> hdt((size_t)nt - 1);
    for (int i = 0; i + 1 < nt; ++i) {
        const double dt = pr->times[i + 1] - pr->times[i];
        hdt[i] = make_double2(dt, sqrt(dt));
    }
    CU(c->dtsq.reserve(sizeof(double2) * hdt.size()));
    CU(cudaMemcpyAsync(c->dtsq.p, hdt.data(), sizeof(double2) * hdt.size(), cudaMemcpyHostToDevice, c->stream));
    pp->dtsq = (const double2 *)c->dtsq.p;
    std::vector<double> hc, hcen;
    if (!pp->linear && !pr->B && !pp->killed && !pp->tdep && !out->path_integrals && !out->model_integrals) {
        // k_terminal, k_traj: folded step coefficients (FAST only).  A grid whose increments agree to the rounding of its
        // own time stamps (seq(0, T, by = h) gives t_i = i*h rounded, so diff(times) scatters by ~ulp(T)) is integrated as
        // the uniform grid it stands for, dt = (t_n - t_0)/n: one coefficient row as a kernel parameter instead of a table.
        const double tmax = fmax(fabs(pr->times[0]), fabs(pr->times[nt - 1]));
        const double dtbar = (pr->times[nt - 1] - pr->times[0]) / (double)(nt - 1), tol = 4.0 * 2.220446049250313e-16 * tmax;
        bool uniform = !(pr->flags & SDEGPU_EXACT_GRID) && dtbar > 0.0;
        for (int i = 0; i + 1 < nt && uniform; ++i) uniform = fabs((pr->times[i + 1] - pr->times[i]) - dtbar) <= tol;
        pp->uniform = uniform;
        if (uniform) fast_step_coefficients(pr->model, pr->params, dtbar, sqrt(dtbar), pp->ucoef);
        else {
            hc.resize((size_t)(nt - 1) * FAST_NCOEF);
            for (int i = 0; i + 1 < nt; ++i) {
                const double dt = pr->times[i + 1] - pr->times[i];
                fast_step_coefficients(pr->model, pr->params, dt, sqrt(dt), &hc[(size_t)i * FAST_NCOEF]);
            }
            CU(c->coef.reserve(sizeof(double) * hc.size()));
            CU(cudaMemcpyAsync(c->coef.p, hc.data(), sizeof(double) * hc.size(), cudaMemcpyHostToDevice, c->stream));
            pp->coef = (const double *)c->coef.p;
        }
    }
    if (pp->killed) {
        CU(c->times.reserve(sizeof(double) * nt));
        CU(cudaMemcpyAsync(c->times.p, pr->times, sizeof(double) * nt, cudaMemcpyHostToDevice, c->stream));
        pp->times = (const double *)c->times.p;
    }
    if (pp->tdep) {
        CU(c->tdep.reserve(sizeof(double) * (size_t)nt * (nx + 1)));
        double *d = (double *)c->tdep.p;
        if (pr->drift_forcing) {
            CU(cudaMemcpyAsync(d, pr->drift_forcing, sizeof(double) * (size_t)nt * nx, cudaMemcpyHostToDevice, c->stream));
            pp->forcing = d;
        }
        if (pr->diffusion_scale) {
            CU(cudaMemcpyAsync(d + (size_t)nt * nx, pr->diffusion_scale, sizeof(double) * nt, cudaMemcpyHostToDevice, c->stream));
            pp->gscale = d + (size_t)nt * nx;
        }
    }
    if (pp->linear) {
        CU(c->params.reserve(sizeof(double) * (size_t)pr->nparams));
        CU(cudaMemcpyAsync(c->params.p, pr->params, sizeof(double) * (size_t)pr->nparams, cudaMemcpyHostToDevice, c->stream));
        pp->linA = (const double *)c->params.p;
        pp->linG = pp->linA + (size_t)nx * nx;
        if (!pr->x0_stride) {
            CU(c->lin_x0.reserve(sizeof(double) * nx));
            CU(cudaMemcpyAsync(c->lin_x0.p, pr->x0, sizeof(double) * nx, cudaMemcpyHostToDevice, c->stream));
            pp->lin_x0 = (const double *)c->lin_x0.p;
        }
        hcen.assign((size_t)nx, 0.0);
        if (out->center) for (int j = 0; j < nx; ++j) hcen[j] = out->center[j];
        CU(c->lin_center.reserve(sizeof(double) * nx));
        CU(cudaMemcpyAsync(c->lin_center.p, hcen.data(), sizeof(double) * nx, cudaMemcpyHostToDevice, c->stream));
        pp->lin_center = (const double *)c->lin_center.p;
    }
    CU(cudaStreamSynchronize(c->stream));                        // the host vectors above go out of scope
    return SDEGPU_OK;
}

static int enqueue(sdegpu_ctx *c, const sdegpu_problem *pr, const sdegpu_output *out, const Prepared &pp, const DevIO &io, bool accum)
{
    const int nx = pr->nx, nt = pr->nt;
    int grid_used = 0, cuerr = 0;
    if (!pp.linear) {
        RunArgs a;
        memset(&a, 0, sizeof a);
        for (int i = 0; i < pr->nparams; ++i) a.P.p[i] = pr->params[i];
        a.dtsq = pp.dtsq;
        a.icdf = c->icdf;
        a.icdf32 = c->icdf32;
        for (int k = 0; k < FAST_NCOEF; ++k) a.ucoef[k] = pp.ucoef[k];
        a.coef = pp.coef;
        a.x0 = io.x0;
        if (!pr->x0_stride) for (int j = 0; j < nx; ++j) a.x0s[j] = pr->x0[j];
        a.B = io.B; a.X = io.X; a.XT = io.XT;
        a.partials = io.ps ? (double *)c->partials.p : nullptr;
        a.hist = io.hist;
        for (int j = 0; j < nx; ++j) a.center[j] = out->center ? out->center[j] : 0.0;
        a.hist_lo = out->hist_lo; a.hist_hi = out->hist_hi;
        a.hist_invw = io.hist ? (double)out->hist_nbins / (out->hist_hi - out->hist_lo) : 0.0;
        a.npaths = io.npaths; a.path_offset = io.path_offset; a.seed = make_philox_key(pr->seed);
        a.nt = nt; a.hist_nbins = out->hist_nbins; a.hist_comp = out->hist_component;
        a.forcing = pp.forcing; a.gscale = pp.gscale;
        if (pp.killed) {
            a.kill_kind = pr->kill_kind; a.kill_comp = pr->kill_comp; a.kill_a = pr->kill_a; a.kill_b = pr->kill_b;
            a.stop_kind = pr->stop_kind; a.stop_comp = pr->stop_comp; a.stop_lo = pr->stop_lo; a.stop_hi = pr->stop_hi;
            a.S0 = pr->S0;
            a.times = pp.times;
            a.Stau = io.Stau; a.tau = io.tau; a.S_tau = io.S_tau; a.Spath = io.S;
        }
        a.func = io.func; a.mfunc = io.mfunc;
        LaunchCfg cfg;
        cfg.killed = pp.killed; cfg.want_func = io.func != nullptr || io.mfunc != nullptr; cfg.uniform = pp.uniform;
        cfg.general = pp.tdep;
        cfg.scheme = pr->scheme; cfg.proj = pr->projection; cfg.arith = pr->B ? pr->arith : SDEGPU_ARITH_FAST;
        if (pp.tdep) cfg.arith = SDEGPU_ARITH_STRICT;
        cfg.src_b = io.B != nullptr; cfg.want_x = io.X != nullptr;
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
        if (io.ps) {
            k_reduce_partials<<<1, 32, 0, c->stream>>>((const double *)c->partials.p, grid_used, 5 * nx, io.ps, accum ? 1 : 0);
            CU(cudaGetLastError());
        }
    } else {
        LinearArgs la;
        memset(&la, 0, sizeof la);
        la.A = pp.linA; la.G = pp.linG;
        la.dtsq = pp.dtsq; la.icdf = c->icdf; la.x0 = pr->x0_stride ? io.x0 : nullptr;
        la.x0_shared = pp.lin_x0;
        la.B = io.B; la.X = io.X; la.XT = io.XT;
        la.npaths = io.npaths; la.path_offset = io.path_offset; la.seed = make_philox_key(pr->seed);
        la.nt = nt; la.d = nx; la.nB = pr->nB; la.scheme = pr->scheme; la.strict = pr->B && pr->arith == SDEGPU_ARITH_STRICT;
        la.moments = io.ps; la.accumulate = accum ? 1 : 0;
        if (io.ps) {                                             // per-block slices of the power sums, reduced in block order
            la.mom_capacity = (nx <= 128 ? 32 : 8) * c->prop.multiProcessorCount;     // d <= 128: up to 16 path tiles per CTA, each with its slice
            CU(c->lin_partials.reserve(sizeof(double) * (size_t)la.mom_capacity * 5 * nx));
            la.mom_partials = (double *)c->lin_partials.p;
        }
        la.center = pp.lin_center;
        la.forcing = pp.forcing;
        int mode = dmma_mode(la, pr->params + (size_t)nx * nx);
        // 32 < d <= 128: the general tensor kernel with two or four path tiles per CTA behind one quantile table keeps 8 warps per SM
        // where the tuned C4 kernels keep 4 (measured, euler, fused generator: d = 64 0.45 -> 0.50 of FP64 peak, d = 100 0.36 -> 0.46,
        // d = 128 0.59 -> 0.64; dense noise d = 64 0.26 -> 0.45, d = 100 level, d = 128 0.62 stays with k_linear_dmma_dense)
        if (((mode == 1 && nx <= 128) || (mode == 2 && nx <= 96)) && !getenv("SDEGPU_LINEAR_DMMA") &&
            linear_tc_eligible(la, pr->params + (size_t)nx * nx))
            mode = 0;
        const size_t dp = ((size_t)nx + 31) / 32 * 32;              // the tensor kernels pad to a multiple of 32 rows
        if (mode == 1) {
            CU(c->scratch.reserve(sizeof(double) * (dp * dp + dp)));
            cuerr = launch_linear_dmma(la, (double *)c->scratch.p, c->prop.multiProcessorCount, c->stream);
        } else if (mode == 2) {
            CU(c->scratch.reserve(sizeof(double) * (dp * dp + dp * 4 * (((size_t)pr->nB + 3) / 4))));
            cuerr = launch_linear_dmma_dense(la, (double *)c->scratch.p, c->prop.multiProcessorCount, c->stream);
        } else if (linear_tc_eligible(la, pr->params + (size_t)nx * nx)) {
            CU(c->scratch.reserve(sizeof(double) * linear_tc_scratch_doubles(la)));
            cuerr = launch_linear_tc(la, pr->params + (size_t)nx * nx, (double *)c->scratch.p, c->prop.multiProcessorCount, c->stream);
        } else
            cuerr = launch_linear(la, c->prop.multiProcessorCount, c->stream);
        if (cuerr) return fail(SDEGPU_ERR_CUDA, "sdegpu_run: linear kernel launch: %s", cudaGetErrorString((cudaError_t)cuerr));
    }
    return SDEGPU_OK;
}

// SDEGPU_DEVICE_PTRS: one launch on the caller's buffers; returns with the work enqueued unless elapsed_ms is asked for
static int run_device_ptrs(sdegpu_ctx *c, const sdegpu_problem *pr, sdegpu_output *out)
{
    Prepared pp;
    int rc = prepare(c, pr, out, &pp);
    if (rc) return rc;
    const bool accum = (pr->flags & SDEGPU_ACCUMULATE) != 0;
    DevIO io;
    io.B = pr->B; io.x0 = pr->x0_stride ? pr->x0 : nullptr; io.Stau = pr->Stau;
    io.X = out->X; io.XT = out->XT; io.tau = out->tau; io.S_tau = out->S_tau; io.S = out->S;
    io.func = out->path_integrals; io.mfunc = out->model_integrals;
    io.ps = out->power_sums; io.hist = (unsigned long long *)out->hist_counts;
    io.npaths = pr->npaths; io.path_offset = pr->path_offset;
    if (io.hist && !accum) CU(cudaMemsetAsync(io.hist, 0, sizeof(unsigned long long) * (out->hist_nbins + 3), c->stream));
    CU(cudaEventRecord(c->ev0, c->stream));
    rc = enqueue(c, pr, out, pp, io, accum);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev1, c->stream));
    if (out->elapsed_ms) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaEventElapsedTime(out->elapsed_ms, c->ev0, c->ev1));
    }
    return SDEGPU_OK;
}

// ---- host pointers: the chunk pipeline -------------------------------------------------------------------------
// Per-path arrays are path-major (X[(p*nx+j)*nt+i], XT[p*nx+j], tau[p], ...), so a path range is one contiguous slice of
// each.  Two granularities:
//   chunk  what one kernel launch integrates.  Kernels with one thread (or lane) per path need tens of thousands of paths
//          to fill the machine, so chunks are large: a quarter of the run, between 256 MB and 4 GB of per-path I/O, two sets
//          of device buffers (chunk k computes while chunk k-1 drains and chunk k+1 arrives).
//   piece  what one PCIe copy moves: 64 MB through two pinned buffers per direction on the copy streams.  Host copies
//          between the caller's arrays and the pinned pieces run on several threads: first touches of a freshly allocated
//          result (R's allocVector, numpy's empty) are page faults, which one thread takes at ~2 GB/s and PCIe delivers
//          at ~50; measured on the B200 box: 14 GB/s with one copy thread (= the driver's pageable path), 35 GB/s with 16.
static const size_t PIECE_BYTES = (size_t)64 << 20;
static const size_t STAGE_MIN_BYTES = (size_t)32 << 20;          // smaller slices take the driver's pageable path (C1: 8 MB each way)

// Persistent host threads for those copies (creating 15 threads per 64 MB piece cost about a third of the copy itself).
// The pool lives while at least one context exists: sdegpu_create takes a reference, sdegpu_destroy drops it and the last
// one joins the threads, so nothing of this library runs after its last context is gone (R may dlclose it then).
static void stream_copy(char *dst, const char *src, size_t n);
class CopyPool {
public:
    struct Task { void *dst; const void *src; size_t n; std::atomic<int> *left; };
    static void acquire()
    {
        std::lock_guard<std::mutex> g(life());
        if (refs()++ == 0) instance() = new CopyPool();
    }
    static void release()
    {
        std::lock_guard<std::mutex> g(life());
        if (--refs() == 0) { delete instance(); instance() = nullptr; }
    }
    static CopyPool *get() { return instance(); }
    unsigned size() const { return (unsigned)th_.size(); }
    void submit(const Task &t)
    {
        { std::lock_guard<std::mutex> g(m_); q_.push_back(t); }
        cv_.notify_one();
    }
private:
    CopyPool()
    {
        // Copy threads: three quarters of the logical CPUs, 16 at most; the submitting thread is one of them.  Measured on two B200
        // boxes of the pool (16 logical CPUs, streaming stores; profiles/r2_e2e_staging_experiment.txt): box A 8 threads 55 GB/s,
        // 16 threads 52, 32 threads 49; box B 6 threads 38.5, 8 threads 41.9, 12 threads 44.0, 16 threads 42.8.
        unsigned hw = std::thread::hardware_concurrency();
        unsigned n = hw >= 4 ? hw * 3 / 4 - 1 : (hw > 1 ? hw - 1 : 0);
        if (n > 15) n = 15;
        if (const char *e = getenv("SDEGPU_COPY_THREADS")) if (atoi(e) > 0) n = (unsigned)atoi(e) - 1;
        for (unsigned k = 0; k < n; ++k) th_.emplace_back([this] { work(); });
    }
    ~CopyPool()
    {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    void work()
    {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [this] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                t = q_.front();
                q_.pop_front();
            }
            stream_copy((char *)t.dst, (const char *)t.src, t.n);
            t.left->fetch_sub(1, std::memory_order_release);
        }
    }
    static CopyPool *&instance() { static CopyPool *p = nullptr; return p; }
    static int &refs() { static int r = 0; return r; }
    static std::mutex &life() { static std::mutex m; return m; }
    std::vector<std::thread> th_;
    std::deque<Task> q_;
    std::mutex m_;
    std::condition_variable cv_;
    bool stop_ = false;
};

static void copy_pool_acquire() { CopyPool::acquire(); }
static void copy_pool_release() { CopyPool::release(); }

// One thread's slice of a staging copy.  Destinations are written once and not read again by this core (the caller's result
// array, or a pinned buffer the DMA engine reads), so non-temporal stores skip the read-for-ownership of every destination line:
// a third less DRAM traffic per byte copied.  SDEGPU_COPY_NT=0 falls back to memcpy.
static bool copy_nt_enabled()
{
    static const bool on = [] { const char *e = getenv("SDEGPU_COPY_NT"); return !(e && e[0] == '0'); }();
    return on;
}
static void stream_copy(char *dst, const char *src, size_t n)
{
    if (n < 65536 || !copy_nt_enabled()) { memcpy(dst, src, n); return; }
    const size_t head = (size_t)(-(uintptr_t)dst) & 63;
    if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
    const size_t blocks = n >> 6;
    for (size_t i = 0; i < blocks; ++i) {
        const __m128i a = _mm_loadu_si128((const __m128i *)src), b = _mm_loadu_si128((const __m128i *)(src + 16));
        const __m128i c = _mm_loadu_si128((const __m128i *)(src + 32)), d = _mm_loadu_si128((const __m128i *)(src + 48));
        _mm_stream_si128((__m128i *)dst, a);
        _mm_stream_si128((__m128i *)(dst + 16), b);
        _mm_stream_si128((__m128i *)(dst + 32), c);
        _mm_stream_si128((__m128i *)(dst + 48), d);
        src += 64;
        dst += 64;
    }
    _mm_sfence();
    if (n & 63) memcpy(dst, src, n & 63);
}

static void parallel_copy(void *dst, const void *src, size_t bytes)
{
    CopyPool *pool = CopyPool::get();
    unsigned nth = pool ? pool->size() + 1 : 1;
    const unsigned by_size = (unsigned)(bytes >> 22);            // a thread per 4 MB
    if (nth > by_size) nth = by_size;
    if (nth <= 1) { memcpy(dst, src, bytes); return; }
    const size_t slice = ((bytes + nth - 1) / nth + 4095) & ~(size_t)4095;
    std::atomic<int> left{0};
    int handed = 0;
    for (unsigned k = 1; k < nth; ++k) {
        const size_t off = slice * k;
        if (off >= bytes) break;
        ++handed;
    }
    left.store(handed);
    for (int k = 1; k <= handed; ++k) {
        const size_t off = slice * (size_t)k, n = bytes - off < slice ? bytes - off : slice;
        pool->submit({(char *)dst + off, (const char *)src + off, n, &left});
    }
    stream_copy((char *)dst, (const char *)src, slice < bytes ? slice : bytes);
    while (left.load(std::memory_order_acquire) > 0) std::this_thread::yield();
}

// host -> device on the H2D stream.  Large slices go piece by piece through the pinned pair; small ones directly.
static int send_slice(sdegpu_ctx *c, char *dev, const char *host, size_t bytes)
{
    if (bytes < STAGE_MIN_BYTES) {
        CU(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->s_h2d));
        return SDEGPU_OK;
    }
    for (int b = 0; b < 2; ++b) CU(c->pin_in[b].reserve(PIECE_BYTES));
    for (size_t off = 0; off < bytes; off += PIECE_BYTES) {
        const size_t n = bytes - off < PIECE_BYTES ? bytes - off : PIECE_BYTES;
        const int b = (int)(c->pin_in_turn++ & 1);
        CU(cudaEventSynchronize(c->ev_pin_in[b]));              // the copy that last read this pinned buffer has finished
        parallel_copy(c->pin_in[b].p, host + off, n);
        CU(cudaMemcpyAsync(dev + off, c->pin_in[b].p, n, cudaMemcpyHostToDevice, c->s_h2d));
        CU(cudaEventRecord(c->ev_pin_in[b], c->s_h2d));
    }
    return SDEGPU_OK;
}

// device -> host on the D2H stream (which already waits for the producing kernel); returns when the slice is in `host`
static int fetch_slice(sdegpu_ctx *c, char *host, const char *dev, size_t bytes)
{
    if (bytes < STAGE_MIN_BYTES) {
        CU(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->s_d2h));
        CU(cudaStreamSynchronize(c->s_d2h));
        return SDEGPU_OK;
    }
    for (int b = 0; b < 2; ++b) CU(c->pin_out[b].reserve(PIECE_BYTES));
    const size_t npieces = (bytes + PIECE_BYTES - 1) / PIECE_BYTES;
    const unsigned turn0 = c->pin_out_turn;
    auto issue = [&](size_t j) -> cudaError_t {
        const size_t off = j * PIECE_BYTES, n = bytes - off < PIECE_BYTES ? bytes - off : PIECE_BYTES;
        const int b = (int)((turn0 + j) & 1);
        cudaError_t e = cudaMemcpyAsync(c->pin_out[b].p, dev + off, n, cudaMemcpyDeviceToHost, c->s_d2h);
        if (e == cudaSuccess) e = cudaEventRecord(c->ev_pin_out[b], c->s_d2h);
        return e;
    };
    CU(issue(0));
    for (size_t j = 0; j < npieces; ++j) {
        if (j + 1 < npieces) CU(issue(j + 1));                   // the other pinned buffer was emptied in the previous turn
        const size_t off = j * PIECE_BYTES, n = bytes - off < PIECE_BYTES ? bytes - off : PIECE_BYTES;
        const int b = (int)((turn0 + j) & 1);
        CU(cudaEventSynchronize(c->ev_pin_out[b]));
        parallel_copy(host + off, c->pin_out[b].p, n);
    }
    c->pin_out_turn = turn0 + (unsigned)npieces;
    return SDEGPU_OK;
}

struct Slice { size_t per_path; size_t off; };                   // bytes per path, offset of the slice in a chunk buffer
static size_t lay(Slice &s, size_t per_path, bool present, size_t &cursor, uint64_t cp)
{
    s.per_path = present ? per_path : 0;
    s.off = cursor;
    if (present) cursor += (per_path * cp + 255) & ~(size_t)255;
    return s.off;
}

static int run_host(sdegpu_ctx *c, const sdegpu_problem *pr, sdegpu_output *out, bool leave_stats, std::atomic<int> *abort_flag)
{
    CU(cudaSetDevice(c->device));
    const int nx = pr->nx, nt = pr->nt;
    const uint64_t np = pr->npaths;
    const bool accum = (pr->flags & SDEGPU_ACCUMULATE) != 0;
    Prepared pp;
    int rc = prepare(c, pr, out, &pp);
    if (rc) return rc;

    // statistics accumulators on the device: [5*nx power sums][nbins+3 counts]
    const int nps = 5 * nx, nh = out->hist_counts ? out->hist_nbins + 3 : 0;
    double *dps = nullptr;
    unsigned long long *dh = nullptr;
    if (out->power_sums || out->hist_counts) {
        CU(c->stats.reserve(sizeof(double) * nps + sizeof(unsigned long long) * (size_t)(nh + 1)));
        if (out->power_sums) dps = (double *)c->stats.p;
        if (out->hist_counts) dh = (unsigned long long *)((double *)c->stats.p + nps);
        if (dh) {
            if (accum) CU(cudaMemcpyAsync(dh, out->hist_counts, sizeof(unsigned long long) * nh, cudaMemcpyHostToDevice, c->stream));
            else CU(cudaMemsetAsync(dh, 0, sizeof(unsigned long long) * nh, c->stream));
        }
        if (dps && accum) CU(cudaMemcpyAsync(dps, out->power_sums, sizeof(double) * nps, cudaMemcpyHostToDevice, c->stream));
    }

    // chunk geometry
    const size_t in_pp = (pr->B ? sizeof(double) * (size_t)pr->nB * nt : 0) + (pr->x0_stride ? sizeof(double) * nx : 0) +
                         (pp.killed && pr->Stau ? sizeof(double) : 0);
    const size_t out_pp = (out->X ? sizeof(double) * (size_t)nx * nt : 0) + (out->XT ? sizeof(double) * nx : 0) + (out->tau ? 8 : 0) +
                          (out->S_tau ? 8 : 0) + (out->S ? sizeof(double) * nt : 0) + (out->path_integrals ? 32 * (size_t)nx : 0) +
                          (out->model_integrals ? 16 * (size_t)nx : 0);
    const bool can_stop = c->intr != nullptr || abort_flag != nullptr;
    uint64_t cp = np;
    if (c->chunk_paths) cp = c->chunk_paths;
    else if (in_pp + out_pp > 0) {
        const double total = (double)(in_pp + out_pp) * (double)np;
        if (total > (double)((size_t)512 << 20)) {
            double target = total / 4.0;                         // four chunks overlap compute and both copy directions
            const double lo = (double)((size_t)256 << 20), hi = (double)((size_t)4 << 30);
            target = target < lo ? lo : (target > hi ? hi : target);
            cp = (uint64_t)(target / (double)(in_pp + out_pp));
        }
    } else if (can_stop) {
        const uint64_t unit = (uint64_t)c->prop.multiProcessorCount * 768;     // whole waves of the register-resident kernel
        cp = (uint64_t)(2e10 / (double)(nt - 1));
        cp = cp / unit * unit;
        if (cp < unit) cp = unit;
    }
    if (cp < 1) cp = 1;
    if (cp > np) cp = np;
    const uint64_t nchunks = (np + cp - 1) / cp;

    size_t cin = 0, cout = 0;
    Slice sB, sx0, sStau, sX, sXT, stau, sStaus, sS, sF, sMF;
    lay(sB, sizeof(double) * (size_t)pr->nB * nt, pr->B != nullptr, cin, cp);
    lay(sx0, sizeof(double) * nx, pr->x0_stride != 0, cin, cp);
    lay(sStau, 8, pp.killed && pr->Stau, cin, cp);
    lay(sX, sizeof(double) * (size_t)nx * nt, out->X != nullptr, cout, cp);
    lay(sXT, sizeof(double) * nx, out->XT != nullptr, cout, cp);
    lay(stau, 8, out->tau != nullptr, cout, cp);
    lay(sStaus, 8, out->S_tau != nullptr, cout, cp);
    lay(sS, sizeof(double) * nt, out->S != nullptr, cout, cp);
    lay(sF, 32 * (size_t)nx, out->path_integrals != nullptr, cout, cp);
    lay(sMF, 16 * (size_t)nx, out->model_integrals != nullptr, cout, cp);
    const int nbuf = nchunks > 1 ? 2 : 1;
    for (int b = 0; b < nbuf; ++b) {
        if (cin) CU(c->pipe_in[b].reserve(cin));
        if (cout) CU(c->pipe_out[b].reserve(cout));
    }
    struct HostIn { const void *p; const Slice *s; } hin[3] = {{pr->B, &sB}, {pr->x0_stride ? pr->x0 : nullptr, &sx0}, {pp.killed ? pr->Stau : nullptr, &sStau}};
    struct HostOut { void *p; const Slice *s; } hout[7] = {{out->X, &sX}, {out->XT, &sXT}, {out->tau, &stau}, {out->S_tau, &sStaus},
                                                           {out->S, &sS}, {out->path_integrals, &sF}, {out->model_integrals, &sMF}};
    float total_ms = 0.f;
    int status = SDEGPU_OK;
    // drain chunk j: its per-path outputs to the caller's arrays (blocking on the host side), then its kernel time
    auto drain = [&](uint64_t j, float *ms_out) -> int {
        const int b = (int)(j & 1) % nbuf;
        const uint64_t p0 = j * cp, n = np - p0 < cp ? np - p0 : cp;
        const char *dout = (const char *)c->pipe_out[b].p;
        if (cout) {
            CU(cudaStreamWaitEvent(c->s_d2h, c->ev_k1[b], 0));
            for (const HostOut &h : hout)
                if (h.p && h.s->per_path) {
                    int r = fetch_slice(c, (char *)h.p + h.s->per_path * p0, dout + h.s->off, h.s->per_path * n);
                    if (r) return r;
                }
        } else
            CU(cudaEventSynchronize(c->ev_k1[b]));
        CU(cudaEventElapsedTime(ms_out, c->ev_k0[b], c->ev_k1[b]));
        return SDEGPU_OK;
    };
    // With bulk traffic in BOTH directions (a supplied B and a trajectory out) the host side of chunk k's upload and of chunk
    // k-1's download run on two threads: PCIe is full duplex and each direction has its own pinned pieces and copy stream.
    const bool duplex = cin >= STAGE_MIN_BYTES && cout >= STAGE_MIN_BYTES && nchunks > 1 && !getenv("SDEGPU_NO_DUPLEX");
    for (uint64_t k = 0; k <= nchunks && status == SDEGPU_OK; ++k) {
        std::thread drainer;
        int drain_rc = SDEGPU_OK;
        float drain_ms = 0.f;
        std::string drain_err;
        if (duplex && k >= 1 && k < nchunks)
            drainer = std::thread([&, k] {
                drain_rc = cudaSetDevice(c->device) == cudaSuccess ? drain(k - 1, &drain_ms) : fail(SDEGPU_ERR_CUDA, "cudaSetDevice");
                if (drain_rc) drain_err = g_err;
            });
        if (k < nchunks) {
            const int b = (int)(k & 1) % nbuf;
            const uint64_t p0 = k * cp, n = np - p0 < cp ? np - p0 : cp;
            char *din = (char *)c->pipe_in[b].p, *dout = (char *)c->pipe_out[b].p;
            if (cin) {
                for (const HostIn &h : hin)
                    if (h.p && h.s->per_path) {
                        rc = send_slice(c, din + h.s->off, (const char *)h.p + h.s->per_path * p0, h.s->per_path * n);
                        if (rc) { status = rc; break; }
                    }
                if (status == SDEGPU_OK) {
                    cudaEventRecord(c->ev_in[b], c->s_h2d);
                    cudaStreamWaitEvent(c->stream, c->ev_in[b], 0);
                }
            }
            if (status == SDEGPU_OK) {
                DevIO io;
                io.B = pr->B ? (const double *)(din + sB.off) : nullptr;
                io.x0 = pr->x0_stride ? (const double *)(din + sx0.off) : nullptr;
                io.Stau = pp.killed && pr->Stau ? (const double *)(din + sStau.off) : nullptr;
                io.X = out->X ? (double *)(dout + sX.off) : nullptr;
                io.XT = out->XT ? (double *)(dout + sXT.off) : nullptr;
                io.tau = out->tau ? (double *)(dout + stau.off) : nullptr;
                io.S_tau = out->S_tau ? (double *)(dout + sStaus.off) : nullptr;
                io.S = out->S ? (double *)(dout + sS.off) : nullptr;
                io.func = out->path_integrals ? (double *)(dout + sF.off) : nullptr;
                io.mfunc = out->model_integrals ? (double *)(dout + sMF.off) : nullptr;
                io.ps = dps; io.hist = dh;
                io.npaths = n; io.path_offset = pr->path_offset + p0;
                cudaEventRecord(c->ev_k0[b], c->stream);
                rc = enqueue(c, pr, out, pp, io, accum || k > 0);
                if (rc) status = rc;
                else cudaEventRecord(c->ev_k1[b], c->stream);
            }
        }
        if (drainer.joinable()) {
            drainer.join();
            if (drain_rc && status == SDEGPU_OK) status = fail(drain_rc, "%s", drain_err.c_str());
            total_ms += drain_ms;
        } else if (k >= 1 && status == SDEGPU_OK) {              // drain chunk k-1 while chunk k computes
            float ms = 0.f;
            rc = drain(k - 1, &ms);
            if (rc) { status = rc; break; }
            total_ms += ms;
        }
        if (k >= 1 && k < nchunks && status == SDEGPU_OK) {      // between chunks: may the run go on?
            if (abort_flag && abort_flag->load()) status = SDEGPU_ERR_INTERRUPT;
            else if (!abort_flag && c->intr && c->intr(c->intr_user)) status = SDEGPU_ERR_INTERRUPT;
        }
    }
    if (status != SDEGPU_OK) {
        cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->s_d2h); cudaStreamSynchronize(c->s_h2d);
        if (status == SDEGPU_ERR_INTERRUPT) return fail(status, "sdegpu_run: interrupted");
        return status;
    }
    if (!leave_stats) {
        if (out->power_sums) CU(cudaMemcpyAsync(out->power_sums, dps, sizeof(double) * nps, cudaMemcpyDeviceToHost, c->stream));
        if (out->hist_counts) CU(cudaMemcpyAsync(out->hist_counts, dh, sizeof(unsigned long long) * nh, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->s_h2d));
    if (out->elapsed_ms) *out->elapsed_ms = total_ms;
    return SDEGPU_OK;
}

// ---- several devices: contiguous path ranges, one host thread per device, one all-reduce of the accumulators --------
static int run_multi(sdegpu_ctx *c, const sdegpu_problem *pr, sdegpu_output *out)
{
    const int ndev = 1 + (int)c->peers.size(), nx = pr->nx, nt = pr->nt;
    const uint64_t np = pr->npaths, per = (np + ndev - 1) / ndev;
    const int nps = 5 * nx, nh = out->hist_counts ? out->hist_nbins + 3 : 0;
    std::vector<sdegpu_ctx *> devs{c};
    devs.insert(devs.end(), c->peers.begin(), c->peers.end());
    std::vector<sdegpu_problem> prs((size_t)ndev, *pr);
    std::vector<sdegpu_output> outs((size_t)ndev, *out);
    std::vector<float> ms((size_t)ndev, 0.f);
    std::vector<int> rcs((size_t)ndev, SDEGPU_OK);
    std::vector<std::string> errs((size_t)ndev);
    std::vector<std::vector<double>> hps((size_t)ndev);
    std::vector<std::vector<unsigned long long>> hhc((size_t)ndev);
    for (int k = 0; k < ndev; ++k) {
        const uint64_t p0 = per * k < np ? per * k : np, n = np - p0 < per ? np - p0 : per;
        sdegpu_problem &q = prs[k];
        sdegpu_output &o = outs[k];
        q.npaths = n; q.path_offset = pr->path_offset + p0;
        if (k > 0) q.flags &= ~(uint32_t)SDEGPU_ACCUMULATE;      // what the caller passed in is added once, on the first device
        if (pr->B) q.B = pr->B + (size_t)p0 * pr->nB * nt;
        if (pr->x0_stride) q.x0 = pr->x0 + (size_t)p0 * nx;
        if (pr->Stau) q.Stau = pr->Stau + p0;
        if (out->X) o.X = out->X + (size_t)p0 * nx * nt;
        if (out->XT) o.XT = out->XT + (size_t)p0 * nx;
        if (out->tau) o.tau = out->tau + p0;
        if (out->S_tau) o.S_tau = out->S_tau + p0;
        if (out->S) o.S = out->S + (size_t)p0 * nt;
        if (out->path_integrals) o.path_integrals = out->path_integrals + (size_t)p0 * nx * 4;
        if (out->model_integrals) o.model_integrals = out->model_integrals + (size_t)p0 * nx * 2;
        o.elapsed_ms = &ms[k];
        if (k > 0) {                                             // device-local accumulators for the host-side fallback
            if (out->power_sums) { hps[k].assign((size_t)nps, 0.0); o.power_sums = hps[k].data(); }
            if (out->hist_counts) { hhc[k].assign((size_t)nh, 0ull); o.hist_counts = (uint64_t *)hhc[k].data(); }
        }
    }
    // the all-reduce needs every device's accumulators resident: decide before the shards run
    NcclApi &nc = nccl_api();
    bool use_nccl = nc.ok && (out->power_sums || out->hist_counts);
    if (use_nccl && !c->nccl_comm && !c->nccl_tried) {
        c->nccl_tried = true;
        std::vector<int> ids((size_t)ndev);
        std::vector<ncclComm_t> comms((size_t)ndev);
        for (int k = 0; k < ndev; ++k) ids[k] = devs[k]->device;
        if (nc.CommInitAll(comms.data(), ndev, ids.data()) == ncclSuccess)
            for (int k = 0; k < ndev; ++k) devs[k]->nccl_comm = comms[k];
    }
    use_nccl = use_nccl && c->nccl_comm != nullptr;

    c->abort_flag.store(0);
    std::atomic<int> running{ndev};
    std::mutex done_m;
    std::condition_variable done_cv;
    auto finished = [&] { { std::lock_guard<std::mutex> g(done_m); running.fetch_sub(1); } done_cv.notify_one(); };
    std::vector<std::thread> th;
    for (int k = 0; k < ndev; ++k)
        th.emplace_back([&, k] {
            if (prs[k].npaths == 0 && !use_nccl) { finished(); return; }
            if (prs[k].npaths == 0) {                            // an empty range still takes part in the all-reduce: zeros
                sdegpu_ctx *d = devs[k];
                cudaSetDevice(d->device);
                d->stats.reserve(sizeof(double) * nps + sizeof(unsigned long long) * (size_t)(nh + 1));
                cudaMemsetAsync(d->stats.p, 0, sizeof(double) * nps + sizeof(unsigned long long) * (size_t)(nh + 1), d->stream);
                cudaStreamSynchronize(d->stream);
            } else {
                rcs[k] = run_host(devs[k], &prs[k], &outs[k], use_nccl, &c->abort_flag);
                if (rcs[k]) errs[k] = g_err;
            }
            finished();
        });
    {                                                            // the interrupt callback runs on the calling thread only: poll it
        std::unique_lock<std::mutex> g(done_m);                  // every 20 ms while waiting for the device threads
        while (running.load() > 0) {
            if (c->intr) {
                done_cv.wait_for(g, std::chrono::milliseconds(20));
                if (running.load() > 0 && !c->abort_flag.load()) {
                    g.unlock();
                    if (c->intr(c->intr_user)) c->abort_flag.store(1);
                    g.lock();
                }
            } else
                done_cv.wait(g, [&] { return running.load() == 0; });
        }
    }
    for (auto &t : th) t.join();
    for (int k = 0; k < ndev; ++k)
        if (rcs[k]) return fail(rcs[k], "device %d: %s", devs[k]->device, errs[k].c_str());
    float worst = 0.f;
    for (int k = 0; k < ndev; ++k) worst = ms[k] > worst ? ms[k] : worst;
    if (out->elapsed_ms) *out->elapsed_ms = worst;               // devices run side by side: the slowest shard

    if (out->power_sums || out->hist_counts) {
        if (use_nccl) {
            ncclResult_t r = nc.GroupStart();
            for (int k = 0; k < ndev && r == ncclSuccess; ++k) {
                sdegpu_ctx *d = devs[k];
                cudaSetDevice(d->device);
                double *ps = (double *)d->stats.p;
                unsigned long long *hc = (unsigned long long *)(ps + nps);
                if (out->power_sums) r = nc.AllReduce(ps, ps, (size_t)nps, ncclDouble, ncclSum, (ncclComm_t)d->nccl_comm, d->stream);
                if (out->hist_counts && r == ncclSuccess) r = nc.AllReduce(hc, hc, (size_t)nh, ncclUint64, ncclSum, (ncclComm_t)d->nccl_comm, d->stream);
            }
            if (r == ncclSuccess) r = nc.GroupEnd(); else nc.GroupEnd();
            if (r != ncclSuccess) return fail(SDEGPU_ERR_CUDA, "sdegpu_run: ncclAllReduce: %s", nc.GetErrorString ? nc.GetErrorString(r) : "error");
            for (int k = 0; k < ndev; ++k) { cudaSetDevice(devs[k]->device); CU(cudaStreamSynchronize(devs[k]->stream)); }
            CU(cudaSetDevice(c->device));
            double *ps = (double *)c->stats.p;
            if (out->power_sums) CU(cudaMemcpyAsync(out->power_sums, ps, sizeof(double) * nps, cudaMemcpyDeviceToHost, c->stream));
            if (out->hist_counts) CU(cudaMemcpyAsync(out->hist_counts, ps + nps, sizeof(unsigned long long) * nh, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            c->last_reduction = "nccl";
        } else {
            for (int k = 1; k < ndev; ++k) {                     // device order: deterministic
                if (out->power_sums) for (int v = 0; v < nps; ++v) out->power_sums[v] += hps[k][v];
                if (out->hist_counts) for (int v = 0; v < nh; ++v) out->hist_counts[v] += hhc[k][v];
            }
            c->last_reduction = "host";
        }
    } else
        c->last_reduction = "none";
    CU(cudaSetDevice(c->device));
    return SDEGPU_OK;
}

extern "C" int sdegpu_run(sdegpu_ctx *c, const sdegpu_problem *pr, sdegpu_output *out)
{
    if (!c || !pr || !out) return fail(SDEGPU_ERR_ARG, "sdegpu_run: NULL argument");
    int rc = validate(pr, out);
    if (rc) return rc;
    const bool dev = (pr->flags & SDEGPU_DEVICE_PTRS) != 0;
    if (dev && !c->peers.empty()) return fail(SDEGPU_ERR_ARG, "sdegpu_run: SDEGPU_DEVICE_PTRS needs a one-device context");
    CU(cudaSetDevice(c->device));
    if (pr->npaths == 0) {
        if (out->elapsed_ms) *out->elapsed_ms = 0.f;
        return SDEGPU_OK;
    }
    if (dev) return run_device_ptrs(c, pr, out);
    if (!c->peers.empty()) return run_multi(c, pr, out);
    return run_host(c, pr, out, false, nullptr);
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

// ---- linear-SDE laws --------------------------------------------------------------
extern "C" {

int sdegpu_lyap(sdegpu_ctx *c, const double *A, const double *Q, int32_t d, double *X)
{
    if (!c || !A || !Q || !X) return fail(SDEGPU_ERR_ARG, "sdegpu_lyap: NULL argument");
    if (d < 1 || d > 8192) return fail(SDEGPU_ERR_ARG, "sdegpu_lyap: d must be in [1, 8192]");
    CU(cudaSetDevice(c->device));
    const char *why = "";
    const int rc = linlaws_lyap(c->stream, A, Q, d, X, &why);
    if (rc) return fail(rc, "sdegpu_lyap: %s", why);
    return SDEGPU_OK;
}

int sdegpu_dlinsde(sdegpu_ctx *c, const double *A, const double *G, int32_t d, int32_t nB, double t, const double *x0, const double *u,
                   int32_t u_cols, const double *S0, double *eAt, double *EX, double *St)
{
    if (!c || !A || !G || !St) return fail(SDEGPU_ERR_ARG, "sdegpu_dlinsde: NULL argument");
    if (d < 1 || d > 2048 || nB < 1) return fail(SDEGPU_ERR_ARG, "sdegpu_dlinsde: d must be in [1, 2048] and nB >= 1");
    if (!isfinite(t)) return fail(SDEGPU_ERR_ARG, "sdegpu_dlinsde: t must be finite");
    if (x0 && !EX) return fail(SDEGPU_ERR_ARG, "sdegpu_dlinsde: EX is NULL although x0 is given");
    if (u && (u_cols != 1 && u_cols != 2)) return fail(SDEGPU_ERR_ARG, "sdegpu_dlinsde: u_cols must be 1 (constant) or 2 (affine in time)");
    if (u && u_cols == 2 && t == 0.0) return fail(SDEGPU_ERR_ARG, "sdegpu_dlinsde: first-order hold needs t != 0");
    CU(cudaSetDevice(c->device));
    const char *why = "";
    const int rc = linlaws_dlinsde(c->stream, A, G, d, nB, t, x0, u, u ? u_cols : 0, S0, eAt, EX, St, &why);
    if (rc) return fail(rc, "sdegpu_dlinsde: %s", why);
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
