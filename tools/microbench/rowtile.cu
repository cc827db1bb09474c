// Memory-side ceiling of the integral kernels' layout (ncols rows of n doubles, one row per lane) for two access patterns:
//   (a) sector: every lane walks its own row with 256-bit loads/stores (what k_integral does: 32 rows x 32 B per warp instruction)
//   (b) tile:   a warp moves a 32-row x 16-element tile cooperatively, 16 lanes per row (2 rows x 128 B per warp instruction,
//               every row's tile grid aligned to that row's own 32-byte sectors) — the pattern of a kernel that stages tiles in
//               shared memory (cp.async in, cooperative stores out) and sums them from there.
// NIN arrays read, NOUT arrays written, W warps per CTA, one CTA per SM.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rowtile rowtile.cu && ./rowtile
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Arrs { const double* in[2]; double* out[3]; };

template <int NIN, int NOUT>
__global__ void __launch_bounds__(1024, 1) k_tile(Arrs q, int n, long ncols, int depth2)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int sub = lane >> 4, col = lane & 15;
    for (long t = (long)blockIdx.x * nw + warp; t * 32 < ncols; t += (long)gridDim.x * nw) {
        const long row0 = t * 32;
        // per instruction i: row = row0 + 2i + sub; its tile grid starts at the row's first sector boundary
        long base[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const long row = row0 + 2 * i + sub, e0 = row * n;
            base[i] = e0 + ((4 - (e0 & 3)) & 3) + col;
        }
        const int ntile = (n - 3) / 16;
        double v[16], w[16], sink = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            v[i] = q.in[0][base[i]];
            if (NIN > 1) v[i] += q.in[1][base[i]];
        }
        for (int m = 0; m < ntile; ++m) {
            if (m + 1 < ntile) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    w[i] = q.in[0][base[i] + 16 * (m + 1)];
                    if (NIN > 1) w[i] += q.in[1][base[i] + 16 * (m + 1)];
                }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
#pragma unroll
                for (int o = 0; o < NOUT; ++o) q.out[o][base[i] + 16 * m] = v[i] + o;
                if (NOUT == 0) sink += v[i];
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = w[i];
        }
        if (NOUT == 0 && sink == 12345.678) q.out[0][0] = sink;      // keeps the loads alive
    }
}

__device__ __forceinline__ void ld_sector(const double* p, double (&s)[4])
{
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(s[0]), "=d"(s[1]), "=d"(s[2]), "=d"(s[3]) : "l"(p));
}
__device__ __forceinline__ void st_sector(double* p, const double (&s)[4])
{
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(s[0]), "d"(s[1]), "d"(s[2]), "d"(s[3]) : "memory");
}

template <int NIN, int NOUT>
__global__ void __launch_bounds__(1024, 1) k_sector(Arrs q, int n, long ncols)
{
    for (long row = (long)blockIdx.x * blockDim.x + threadIdx.x; row < ncols; row += (long)gridDim.x * blockDim.x) {
        const long e0 = row * n, b = e0 + ((4 - (e0 & 3)) & 3);
        const int ng = (n - 3) / 4;
        double a[4], c[4], na[4], nc[4], sink = 0.0;
        ld_sector(q.in[0] + b, a);
        if (NIN > 1) ld_sector(q.in[1] + b, c);
        for (int g = 0; g < ng; ++g) {
            if (g + 1 < ng) {
                ld_sector(q.in[0] + b + 4 * (g + 1), na);
                if (NIN > 1) ld_sector(q.in[1] + b + 4 * (g + 1), nc);
            }
            double s[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) s[k] = NIN > 1 ? a[k] + c[k] : a[k];
#pragma unroll
            for (int o = 0; o < NOUT; ++o) {
                double so[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) so[k] = s[k] + o;
                st_sector(q.out[o] + b + 4 * g, so);
            }
            if (NOUT == 0) sink += s[0] + s[1] + s[2] + s[3];
#pragma unroll
            for (int k = 0; k < 4; ++k) { a[k] = na[k]; c[k] = nc[k]; }
        }
        if (NOUT == 0 && sink == 12345.678) q.out[0][0] = sink;
    }
}

template <int NIN, int NOUT>
void run(const char* name, Arrs q, int n, long ncols, int sm)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double bytes = 8.0 * (double)ncols * ((n - 3) / 16 * 16) * (NIN + NOUT);
    for (int tile = 0; tile < 2; ++tile)
        for (int w : {2, 4, 8, 16, 32}) {
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaEventRecord(e0));
                if (tile) k_tile<NIN, NOUT><<<sm, 32 * w>>>(q, n, ncols, 0);
                else k_sector<NIN, NOUT><<<sm, 32 * w>>>(q, n, ncols);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            CK(cudaGetLastError());
            printf("%-22s %-7s W=%2d  %8.3f ms  %7.1f GB/s\n", name, tile ? "tile" : "sector", w, best, bytes / best / 1e6);
        }
}

int main()
{
    int dev = 0, sm = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    const int n = 1001;
    const long ncols = 524288;
    Arrs q;
    const size_t elems = (size_t)10001 * 75776;                   // the larger of the two shapes
    for (int i = 0; i < 2; ++i) { double* p; CK(cudaMalloc(&p, sizeof(double) * elems)); CK(cudaMemset(p, 0, sizeof(double) * elems)); q.in[i] = p; }
    for (int i = 0; i < 3; ++i) { CK(cudaMalloc(&q.out[i], sizeof(double) * elems)); }
    run<2, 3>("stochint l/r/c (2 in 3 out)", q, n, ncols, sm);
    run<2, 1>("itointegral (2 in 1 out)", q, n, ncols, sm);
    run<1, 1>("quadratic var (1 in 1 out)", q, n, ncols, sm);
    run<1, 0>("read only (1 in)", q, n, ncols, sm);
    // config C3 with a supplied B: 75,776 paths x 10,001 points, one row of B in, two rows of X out per path
    run<1, 2>("C3 supplied B (1 in 2 out)", q, 10001, 75776, sm);
    run<0 + 1, 2>("C3 supplied B, 2 waves", q, 10001, 75776, 2 * sm);
    return 0;
}
