// Store-path ceiling for the trajectory layout: P rows of nt doubles, each warp owns 32 rows and
// walks them in time.  Compares (a) per-lane 32-byte sector stores (what k_traj did in round 1) with
// (b) tiles staged in shared memory (STS.128, hardware swizzle) and written by 2-D TMA tile stores.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_store tma_store.cu && ./tma_store
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Maps { CUtensorMap m[8]; int head[8]; };

constexpr int THREADS = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// rows of class c (c = row % Q): q-th row of the class is row c + q*Q.  W doubles per tile row.
template <int W, int NBUF>
__global__ void __launch_bounds__(THREADS, 1) k_tma(const __grid_constant__ Maps maps, int Q, int nq, int nt, int waves) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int TILE = 32 * W * 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* buf = smem + warp * NBUF * TILE;
    const int wpc = (nq + 31) / 32;                               // warps per class
    const int nwarps = Q * wpc;
    for (int gw = blockIdx.x * (THREADS / 32) + warp; gw < nwarps; gw += gridDim.x * (THREADS / 32)) {
        const int cls = gw / wpc, q0 = (gw - cls * wpc) * 32;
        const int head = maps.head[cls];
        double x = 1.0 + lane;
        int b = 0;
        const int ntile = (nt - head + W - 1) / W;
        for (int g = 0; g < ntile; ++g) {
            unsigned char* tile = buf + b * TILE;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
            __syncwarp();
#pragma unroll
            for (int k = 0; k < W / 2; ++k) {
                x = fma(x, 1.0000001, 0.5);
                const double y = x + 1.0;
                uint32_t off = lane * (W * 8) + k * 16;
                if (W == 4) off ^= ((off >> 7) & 1) << 4;
                if (W == 8) off ^= ((off >> 7) & 3) << 4;
                if (W == 16) off ^= ((off >> 7) & 7) << 4;
                *reinterpret_cast<double2*>(tile + off) = make_double2(x, y);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                             ::"l"(&maps.m[cls]), "r"(g * W), "r"(q0), "r"(smem_u32(tile)) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            b = (b + 1) % NBUF;
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// coalesced pure-write ceiling
__global__ void __launch_bounds__(THREADS, 1) k_fill(double* X, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)THREADS + threadIdx.x; i < n4; i += (size_t)gridDim.x * THREADS)
        asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(X + 4 * i), "d"((double)i) : "memory");
}

// 3-D box: [16 doubles][32 rows][NCH chunks] -> NCH*128 bytes per row in one TMA store
template <int NCH, int NW = 8, int HINT = 0>       // HINT: L2 cache policy of the store, 0 none, 1 evict_first, 2 evict_last, 3 evict_unchanged
__global__ void __launch_bounds__(NW * 32, 1) k_tma3(const __grid_constant__ Maps maps, int Q, int nq, int nt) {
    uint64_t policy = 0;
    if (HINT == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    if (HINT == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    if (HINT == 3) asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(policy));
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int TILE = NCH * 32 * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* tile = smem + warp * TILE;
    const int wpc = (nq + 31) / 32, nwarps = Q * wpc;
    for (int gw = blockIdx.x * NW + warp; gw < nwarps; gw += gridDim.x * NW) {
        const int cls = gw / wpc, q0 = (gw - cls * wpc) * 32;
        const int head = maps.head[cls];
        double x = 1.0 + lane;
        const int ntile = (nt - head + 16 * NCH - 1) / (16 * NCH);
        for (int g = 0; g < ntile; ++g) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    x = fma(x, 1.0000001, 0.5);
                    uint32_t off = c * 4096 + lane * 128 + ((k ^ (lane & 7)) << 4);
                    *reinterpret_cast<double2*>(tile + off) = make_double2(x, x + 1.0);
                }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (HINT == 0)
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                                 ::"l"(&maps.m[cls]), "r"(0), "r"(q0), "r"(g * NCH), "r"(smem_u32(tile)) : "memory");
                else
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2, %3}], [%4], %5;"
                                 ::"l"(&maps.m[cls]), "r"(0), "r"(q0), "r"(g * NCH), "r"(smem_u32(tile)), "l"(policy) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// reference: per-lane 256-bit sector stores, 4 doubles per lane per store
template <int HINT, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_stg(double* X, int Q, int nq, int nt, const int* heads) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = (nq + 31) / 32, nwarps = Q * wpc;
    for (int gw = blockIdx.x * (THREADS / 32) + warp; gw < nwarps; gw += gridDim.x * (THREADS / 32)) {
        const int cls = gw / wpc, q = (gw - cls * wpc) * 32 + lane;
        if (q >= nq) continue;
        const int head = heads[cls];
        double* row = X + (size_t)(cls + (size_t)q * Q) * nt + head;
        double x = 1.0 + lane;
        const int ntile = (nt - head) / 4;
        for (int g = 0; g < ntile; ++g) {
            double v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { x = fma(x, 1.0000001, 0.5); v[k] = x; }
            if (HINT == 0) asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(row + 4 * g), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
            if (HINT == 1) asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(row + 4 * g), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
            if (HINT == 2) asm volatile("st.global.L2::evict_last.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(row + 4 * g), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
        }
    }
}

template <int W, int NBUF>
static void run_tma(const Maps& maps, int Q, int nq, int nt, int sm, double bytes, const char* name) {
    const size_t smem = (size_t)(THREADS / 32) * NBUF * 32 * W * 8;
    CK(cudaFuncSetAttribute(k_tma<W, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        k_tma<W, NBUF><<<sm, THREADS, smem>>>(maps, Q, nq, nt, 0);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep && ms < best) best = ms;
    }
    printf("%-28s W=%2d NBUF=%d  %8.3f ms  %8.1f GB/s\n", name, W, NBUF, best, bytes / best / 1e6);
}

int main(int argc, char** argv) {
    int sm = 148;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); sm = prop.multiProcessorCount;
    const int nt = argc > 1 ? atoi(argv[1]) : 10001;
    const int rows = argc > 2 ? atoi(argv[2]) : 2 * sm * 512 * 2;         // 2 waves x 2 components
    int Q = 1; while (((long long)Q * nt) % 4) Q *= 2;
    const int nq = rows / Q;
    double* X; CK(cudaMalloc(&X, (size_t)rows * nt * 8));
    CK(cudaMemset(X, 0, (size_t)rows * nt * 8));
    EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres));
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    const double bytes = (double)rows * nt * 8;
    std::vector<int> heads(Q);
    for (int W : {4, 8, 16}) {
        Maps maps;
        for (int c = 0; c < Q; ++c) {
            const int phi = (int)(((long long)c * nt) % 4), head = (4 - phi) % 4;
            heads[c] = maps.head[c] = head;
            cuuint64_t dims[2] = {(cuuint64_t)(nt - head), (cuuint64_t)nq}, strides[1] = {(cuuint64_t)Q * nt * 8};
            cuuint32_t box[2] = {(cuuint32_t)W, 32}, es[2] = {1, 1};
            CUtensorMapSwizzle sw = W == 4 ? CU_TENSOR_MAP_SWIZZLE_32B : W == 8 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
            CUresult r = enc(&maps.m[c], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, X + (size_t)c * nt + head, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d (W=%d c=%d)\n", (int)r, W, c); return 1; }
        }
        if (W == 4) { run_tma<4, 1>(maps, Q, nq, nt, sm, bytes, "tma tile store"); run_tma<4, 2>(maps, Q, nq, nt, sm, bytes, "tma tile store"); run_tma<4, 4>(maps, Q, nq, nt, sm, bytes, "tma tile store"); }
        if (W == 8) { run_tma<8, 1>(maps, Q, nq, nt, sm, bytes, "tma tile store"); run_tma<8, 2>(maps, Q, nq, nt, sm, bytes, "tma tile store"); }
        if (W == 16) { run_tma<16, 1>(maps, Q, nq, nt, sm, bytes, "tma tile store"); run_tma<16, 2>(maps, Q, nq, nt, sm, bytes, "tma tile store"); }
    }
    // spot check of the last TMA variant: row r, element i (i >= head) must be finite and non-zero
    {
        std::vector<double> h(nt);
        CK(cudaMemcpy(h.data(), X + (size_t)5 * nt, nt * 8, cudaMemcpyDeviceToHost));
        int zeros = 0; for (int i = heads[5 % Q]; i < nt; ++i) zeros += h[i] == 0.0;
        printf("row 5: %d zeros after head %d (expect 0); first values %.6f %.6f %.6f\n", zeros, heads[5 % Q], h[heads[5 % Q]], h[heads[5 % Q] + 1], h[nt - 1]);
    }
    int* dheads; CK(cudaMalloc(&dheads, Q * 4)); CK(cudaMemcpy(dheads, heads.data(), Q * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](const char* name, auto launch) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("%-44s %8.3f ms  %8.1f GB/s\n", name, best, bytes / best / 1e6);
    };
    timeit("per-lane STG.256 sectors, 16 warps", [&] { k_stg<0, 512><<<sm, 512>>>(X, Q, nq, nt, dheads); });
    timeit("per-lane STG.256 sectors, 12 warps", [&] { k_stg<0, 384><<<sm, 384>>>(X, Q, nq, nt, dheads); });
    timeit("per-lane STG.256 sectors, 8 warps", [&] { k_stg<0, 256><<<sm, 256>>>(X, Q, nq, nt, dheads); });
    timeit("per-lane STG.256 sectors, 6 warps", [&] { k_stg<0, 192><<<sm, 192>>>(X, Q, nq, nt, dheads); });
    timeit("per-lane STG.256 sectors, 4 warps", [&] { k_stg<0, 128><<<sm, 128>>>(X, Q, nq, nt, dheads); });
    timeit("per-lane STG.256 sectors, 2 warps", [&] { k_stg<0, 64><<<sm, 64>>>(X, Q, nq, nt, dheads); });
    timeit("per-lane STG.256 sectors .cs, 8 warps", [&] { k_stg<1, 256><<<sm, 256>>>(X, Q, nq, nt, dheads); });
    timeit("per-lane STG.256 sectors evict_last, 8 warps", [&] { k_stg<2, 256><<<sm, 256>>>(X, Q, nq, nt, dheads); });
    timeit("coalesced fill (write ceiling)", [&] { k_fill<<<sm * 4, THREADS>>>(X, (size_t)rows * nt / 4); });
    timeit("cudaMemsetAsync", [&] { CK(cudaMemsetAsync(X, 0, (size_t)rows * nt * 8)); });
    // 3-D boxes: NCH x 128 bytes per row per TMA store
    auto run3 = [&](auto kern, int nch, const char* name, int nw = 8) {
        Maps maps;
        for (int c = 0; c < Q; ++c) {
            const int phi = (int)(((long long)c * nt) % 4), head = (4 - phi) % 4;
            maps.head[c] = head;
            cuuint64_t dims[3] = {16, (cuuint64_t)nq, (cuuint64_t)((nt - head) / 16)}, strides[2] = {(cuuint64_t)Q * nt * 8, 128};
            cuuint32_t box[3] = {16, 32, (cuuint32_t)nch}, es[3] = {1, 1, 1};
            CUresult r = enc(&maps.m[c], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, X + (size_t)c * nt + head, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode3 failed %d\n", (int)r); exit(1); }
        }
        const int smem = nw * nch * 4096;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        timeit(name, [&] { kern<<<sm, nw * 32, smem>>>(maps, Q, nq, nt); });
    };
    run3(k_tma3<1>, 1, "tma 3-D box 128 B/row, 8 warps");
    run3(k_tma3<2>, 2, "tma 3-D box 256 B/row, 8 warps");
    run3(k_tma3<4>, 4, "tma 3-D box 512 B/row, 8 warps");
    run3(k_tma3<6>, 6, "tma 3-D box 768 B/row, 8 warps");
    run3(k_tma3<1, 16>, 1, "tma 3-D box 128 B/row, 16 warps", 16);
    run3(k_tma3<2, 16>, 2, "tma 3-D box 256 B/row, 16 warps", 16);
    run3(k_tma3<1, 4>, 1, "tma 3-D box 128 B/row, 4 warps", 4);
    run3(k_tma3<2, 4>, 2, "tma 3-D box 256 B/row, 4 warps", 4);
    run3(k_tma3<4, 4>, 4, "tma 3-D box 512 B/row, 4 warps", 4);
    run3(k_tma3<12, 4>, 12, "tma 3-D box 1536 B/row, 4 warps", 4);
    run3(k_tma3<2, 8, 1>, 2, "tma 3-D box 256 B/row, 8 warps, evict_first");
    run3(k_tma3<2, 8, 2>, 2, "tma 3-D box 256 B/row, 8 warps, evict_last");
    run3(k_tma3<2, 8, 3>, 2, "tma 3-D box 256 B/row, 8 warps, evict_unchanged");
    run3(k_tma3<2, 16, 1>, 2, "tma 3-D box 256 B/row, 16 warps, evict_first", 16);
    run3(k_tma3<2, 4, 1>, 2, "tma 3-D box 256 B/row, 4 warps, evict_first", 4);
    run3(k_tma3<2, 2>, 2, "tma 3-D box 256 B/row, 2 warps", 2);
    run3(k_tma3<8, 2>, 8, "tma 3-D box 1024 B/row, 2 warps", 2);
    CK(cudaDeviceSynchronize());
    return 0;
}
