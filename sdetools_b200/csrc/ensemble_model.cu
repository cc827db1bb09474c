// One translation unit per catalogue model (compiled with -DSDEGPU_TU_MODEL=<id>):
// instantiates k_terminal / k_traj / k_stream / k_killed(_refill) / k_functionals for every scheme x projection x arithmetic
// combination of that model and exposes one launcher.
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

template <int SCHEME, int PROJ>
static int launch_traj(const RunArgs &a, const LaunchCfg &c, int *grid_used)
{
    auto kernel = k_traj<M, SCHEME, PROJ>;
    const size_t smem = Icdf32Layout<TRAJ_REPL>::BYTES + sizeof(double2) * TRAJ_CHUNK + (a.hist ? (size_t)(a.hist_nbins + 3) * 4 : 0);
    constexpr int TRAJ_THREADS = TrajThreads<M::NX>::value;
    const uint64_t work = (a.npaths + TRAJ_THREADS - 1) / TRAJ_THREADS;
    const int grid = c.grid > 0 ? min(c.grid, 32 * c.sm_count) : occupancy_grid(kernel, TRAJ_THREADS, smem, c.sm_count, work);
    const int vec_ok = (reinterpret_cast<uintptr_t>(a.X) & 31) == 0;
    const unsigned long long span = (unsigned long long)M::NX * (unsigned long long)a.nt;
    const int Q = (span % 4 == 0) ? 1 : ((span % 2 == 0) ? 2 : 4);   // lanes take every Q-th path: uniform sector phase
    kernel<<<grid, TRAJ_THREADS, smem, c.stream>>>(a, vec_ok, Q);
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
    auto kernel = a.X || a.Spath ? k_killed<M, SCHEME, PROJ> : k_killed_refill<M, SCHEME, PROJ>;   // no trajectory: lanes refill
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
    if (c.killed) return launch_killed<SCHEME, PROJ>(a, c, g);
    if (c.want_func)
        return c.src_b && c.arith == SDEGPU_ARITH_STRICT ? launch_functionals<SCHEME, PROJ, Strict>(a, c, g)
                                                         : launch_functionals<SCHEME, PROJ, Fast>(a, c, g);
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
