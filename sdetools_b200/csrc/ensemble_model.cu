// One translation unit per catalogue model (compiled with -DSDEGPU_TU_MODEL=<id>):
// instantiates k_terminal / k_traj / k_stream / k_killed(_refill) / k_functionals for every scheme x projection x arithmetic
// combination of that model and exposes one launcher.
#include <cuda.h>
#include <cstdlib>

#include "ensemble.cuh"

#ifndef SDEGPU_TU_MODEL
#error "compile with -DSDEGPU_TU_MODEL=<sdegpu_model id>"
#endif

namespace sdegpu {

using M = Model<SDEGPU_TU_MODEL>;

template <class K>
static int occupancy_grid(K kernel, int threads, size_t smem, int sm_count, uint64_t work_blocks, int max_per_sm = 0)
{
    int per_sm = 0;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (per_sm < 1) per_sm = 1;
    if (max_per_sm > 0 && per_sm > max_per_sm) per_sm = max_per_sm;
    if (per_sm > 32) per_sm = 32;                  // the context's block-partials buffer holds 32 blocks per SM
    uint64_t g = (uint64_t)per_sm * sm_count;
    if (g > work_blocks) g = work_blocks;
    return (int)(g < 1 ? 1 : g);
}

template <int SCHEME, int PROJ>
static int launch_terminal(const RunArgs &a, const LaunchCfg &c, int *grid_used)
{
    auto kernel = c.uniform ? k_terminal<M, SCHEME, PROJ, true> : k_terminal<M, SCHEME, PROJ, false>;
    const size_t smem = Icdf32Layout<TERMINAL_REPL>::BYTES + (a.hist ? (size_t)(a.hist_nbins + 3) * 4 : 0);
    const uint64_t work = (a.npaths + TERMINAL_THREADS - 1) / TERMINAL_THREADS;
    const int grid = c.grid > 0 ? min(c.grid, 32 * c.sm_count) : occupancy_grid(kernel, TERMINAL_THREADS, smem, c.sm_count, work);
    kernel<<<grid, TERMINAL_THREADS, smem, c.stream>>>(a);
    *grid_used = grid;
    return (int)cudaGetLastError();
}

// Warps per CTA of k_traj (one CTA per SM).  Rows in flight per SM = 32 * warps * nx; the memory system prefers ~256
// (profiles/r1_tma_store_microbench.txt), the generator wants warps.  The environment variable is a tuning knob.
#ifndef SDEGPU_TRAJ_WARPS_NX1
#define SDEGPU_TRAJ_WARPS_NX1 12
#endif
#ifndef SDEGPU_TRAJ_WARPS_NX2
#define SDEGPU_TRAJ_WARPS_NX2 4
#endif

typedef CUresult (*tensor_map_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                         const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tensor_map_encode_fn tensor_map_encoder()
{
    static tensor_map_encode_fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<tensor_map_encode_fn>(p);
    }();
    return fn;
}

static int gcd_int(int a, int b) { return b ? gcd_int(b, a % b) : a; }

template <int SCHEME, int PROJ>
static int launch_traj(const RunArgs &a, const LaunchCfg &c, int *grid_used)
{
    constexpr int NX = M::NX;
    static_assert(sizeof(CUtensorMap) == 128 && NX <= 2, "TrajMaps layout");
    auto kernel = c.uniform ? k_traj<M, SCHEME, PROJ, true> : k_traj<M, SCHEME, PROJ, false>;
    if (reinterpret_cast<uintptr_t>(a.X) & 7) return (int)cudaErrorMisalignedAddress;
    // row classes: S consecutive rows start at S different offsets mod 4 elements, then the pattern repeats
    const int Q = 4 / gcd_int(a.nt & 3 ? a.nt & 3 : 4, 4), S = Q / gcd_int(Q, NX) * NX, sp = S / NX;
    const uint32_t xoff = (uint32_t)((reinterpret_cast<uintptr_t>(a.X) >> 3) & 3u);
    TrajGeom geo;
    TrajMaps maps;
    geo.sp = sp;
    int nchunks = a.nt;
    for (int cls = 0; cls < S; ++cls) {
        const uint32_t phase = (xoff + (uint32_t)(((uint64_t)cls * (uint64_t)a.nt) & 3u)) & 3u;
        geo.head[cls] = 4 - (int)phase;                          // 1..4: the x0 element is always written directly
        const int n = (a.nt - geo.head[cls]) / 16;
        nchunks = n < nchunks ? n : nchunks;
    }
    for (int cls = S; cls < TRAJ_MAX_CLASSES; ++cls) geo.head[cls] = 4;
    geo.nchunks = nchunks < 0 ? 0 : nchunks;
    if (geo.nchunks > 0) {
        tensor_map_encode_fn enc = tensor_map_encoder();
        if (!enc) return (int)cudaErrorNotSupported;
        const uint64_t nrows = a.npaths * NX;
        for (int cls = 0; cls < S; ++cls) {
            const uint64_t nq = nrows > (uint64_t)cls ? (nrows - cls + S - 1) / S : 1;
            cuuint64_t dims[3] = {16, (cuuint64_t)nq, (cuuint64_t)geo.nchunks};
            cuuint64_t strides[2] = {(cuuint64_t)S * (cuuint64_t)a.nt * 8u, 128u};
            cuuint32_t box[3] = {16, 32, (cuuint32_t)TrajCfg<NX>::CH}, es[3] = {1, 1, 1};
            const CUresult r = enc(reinterpret_cast<CUtensorMap *>(maps.m[cls]), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3,
                                   a.X + (uint64_t)cls * a.nt + geo.head[cls], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
        }
    }
    const size_t fixed = Icdf32Layout<TRAJ_REPL>::BYTES + (a.hist ? (size_t)((a.hist_nbins + 3 + 3) & ~3) * 4 : 0);
    const size_t per_warp = (size_t)NX * TrajCfg<NX>::NSLOT * 4096 + sizeof(double) * 5 * NX;
    int dev = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    int warps = NX == 1 ? SDEGPU_TRAJ_WARPS_NX1 : SDEGPU_TRAJ_WARPS_NX2;
    if (const char *e = getenv("SDEGPU_TRAJ_WARPS")) warps = atoi(e) > 0 ? atoi(e) : warps;
    const int fit = (int)(((size_t)max_smem - 1024 - fixed) / per_warp);
    if (warps > fit) warps = fit;
    if (warps > TrajCfg<NX>::MAX_THREADS / 32) warps = TrajCfg<NX>::MAX_THREADS / 32;
    if (warps < 1) return (int)cudaErrorInvalidConfiguration;
    const size_t smem = fixed + per_warp * warps;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const uint64_t tasks = (a.npaths + 31) / 32, work = (tasks + warps - 1) / warps;
    uint64_t g = (uint64_t)c.sm_count;                           // one CTA per SM (the rings fill its shared memory)
    if (g > work) g = work;
    const int grid = c.grid > 0 ? min(c.grid, 32 * c.sm_count) : (int)g;
    kernel<<<grid, 32 * warps, smem, c.stream>>>(a, maps, geo);
    *grid_used = grid;
    return (int)cudaGetLastError();
}

template <int SCHEME, int PROJ, class A>
static int launch_stream(const RunArgs &a, const LaunchCfg &c, int *grid_used)
{
    auto kernel = k_stream<M, SCHEME, PROJ, A>;
    const size_t smem = a.hist ? (size_t)(a.hist_nbins + 3) * 4 : 0;
    const uint64_t work = (a.npaths + STREAM_THREADS - 1) / STREAM_THREADS;
    const int grid = c.grid > 0 ? min(c.grid, 32 * c.sm_count) : occupancy_grid(kernel, STREAM_THREADS, smem, c.sm_count, work, SDEGPU_STREAM_BLOCKS_PER_SM);
    const int vec_in = (reinterpret_cast<uintptr_t>(a.B) & 31) == 0;
    const int vec_out = a.X && (reinterpret_cast<uintptr_t>(a.X) & 31) == 0;
    kernel<<<grid, STREAM_THREADS, smem, c.stream>>>(a, vec_in, vec_out);
    *grid_used = grid;
    return (int)cudaGetLastError();
}

template <int SCHEME, int PROJ>
static int launch_killed(const RunArgs &a, const LaunchCfg &c, int *grid_used)
{
    // no trajectory: lanes refill (pointless without a rule: every path runs the whole grid)
    auto kernel = a.X || a.Spath || !c.killed ? k_killed<M, SCHEME, PROJ> : k_killed_refill<M, SCHEME, PROJ>;
    const size_t smem = (a.B ? 0 : ICDF_TABLE_BYTES) + (a.hist ? (size_t)(a.hist_nbins + 3) * 4 : 0);
    const uint64_t work = (a.npaths + KILLED_THREADS - 1) / KILLED_THREADS;
    const int grid = c.grid > 0 ? min(c.grid, 32 * c.sm_count) : occupancy_grid(kernel, KILLED_THREADS, smem, c.sm_count, work);
    kernel<<<grid, KILLED_THREADS, smem, c.stream>>>(a);
    *grid_used = grid;
    return (int)cudaGetLastError();
}

template <int SCHEME, int PROJ, class A>
static int launch_functionals(const RunArgs &a, const LaunchCfg &c, int *grid_used)
{
    auto kernel = k_functionals<M, SCHEME, PROJ, A>;
    const size_t smem = (a.B ? 0 : ICDF_TABLE_BYTES) + (a.hist ? (size_t)(a.hist_nbins + 3) * 4 : 0);
    const uint64_t work = (a.npaths + FUNC_THREADS - 1) / FUNC_THREADS;
    const int grid = c.grid > 0 ? min(c.grid, 32 * c.sm_count) : occupancy_grid(kernel, FUNC_THREADS, smem, c.sm_count, work);
    kernel<<<grid, FUNC_THREADS, smem, c.stream>>>(a);
    *grid_used = grid;
    return (int)cudaGetLastError();
}

template <int SCHEME, int PROJ>
static int dispatch_mode(const RunArgs &a, const LaunchCfg &c, int *g)
{
    if (c.want_func)
        return c.src_b && c.arith == SDEGPU_ARITH_STRICT ? launch_functionals<SCHEME, PROJ, Strict>(a, c, g)
                                                         : launch_functionals<SCHEME, PROJ, Fast>(a, c, g);
    if (c.killed || c.general) return launch_killed<SCHEME, PROJ>(a, c, g);
    if (c.src_b)
        return c.arith == SDEGPU_ARITH_STRICT ? launch_stream<SCHEME, PROJ, Strict>(a, c, g)
                                              : launch_stream<SCHEME, PROJ, Fast>(a, c, g);
    if (c.want_x) return launch_traj<SCHEME, PROJ>(a, c, g);
    return launch_terminal<SCHEME, PROJ>(a, c, g);
}

template <int SCHEME>
static int dispatch_proj(const RunArgs &a, const LaunchCfg &c, int *g)
{
    switch (c.proj) {
    case SDEGPU_PROJ_IDENTITY: return dispatch_mode<SCHEME, SDEGPU_PROJ_IDENTITY>(a, c, g);
    case SDEGPU_PROJ_ABS: return dispatch_mode<SCHEME, SDEGPU_PROJ_ABS>(a, c, g);
    default: return dispatch_mode<SCHEME, SDEGPU_PROJ_RELU>(a, c, g);
    }
}

#define SDEGPU_CAT2(a, b) a##b
#define SDEGPU_CAT(a, b) SDEGPU_CAT2(a, b)
int SDEGPU_CAT(launch_model_, SDEGPU_TU_MODEL)(const RunArgs &a, const LaunchCfg &c, int *g)
{
    return c.scheme == SDEGPU_EULER ? dispatch_proj<SDEGPU_EULER>(a, c, g) : dispatch_proj<SDEGPU_HEUN>(a, c, g);
}

}  // namespace sdegpu
