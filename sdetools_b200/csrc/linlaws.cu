// lyap() and dLinSDE() at device dimension (R/sdetools.R:593-600, :508-572).
//
// The reference solves the Lyapunov equation through the d^2 x d^2 Kronecker system (O(d^6)) and takes its matrix
// exponentials from Matrix::expm.  Here every O(d^3) step is a product of square matrices:
//   expm   scaling and squaring with the [13/13] Pade approximant (Higham 2005): 6 products, one LU solve, s squarings
//   lyap   d <= 40: the reference's Kronecker system itself (LU with partial pivoting), so small systems follow its method;
//          larger: the bilinear transform A_d = (A - gI)^-1 (A + gI), Q_d = 2g (A - gI)^-1 Q (A - gI)^-T turns
//          A X + X A' + Q = 0 into the Stein equation X = A_d X A_d' + Q_d, summed by doubling
//          X <- X + A_k X A_k', A_{k+1} = A_k^2 (quadratic convergence for a Hurwitz A; -A is tried for an anti-stable one)
//   dLinSDE  the reference's own formulas on top of those: St = Sinf - eAt (Sinf - S0) eAt', Van Loan's block exponential
//          when the Lyapunov equation has no solution (:520-534), zero- and first-order hold for EX (:545-571)
// Products of matrices with d >= 96 run on the GPU (k_dgemm: 64 x 64 tiles in shared memory, 4 x 4 register micro-tiles,
// FP64 FMA; B200's FP64 tensor rate equals its CUDA-core FP64 rate, so plain FMAs are the right tool); the LU
// factorisations (one or two per call) and everything smaller stay on the host, threaded over right-hand sides.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/sdegpu.h"
#include "linlaws.cuh"

namespace sdegpu {

// C = A op(B), all n x n column-major; TB: B is used transposed
template <bool TB>
__global__ void __launch_bounds__(256) k_dgemm(const double *__restrict__ A, const double *__restrict__ B, double *__restrict__ C, int n)
{
    constexpr int T = 64, K = 16;
    __shared__ double As[K][T + 1], Bs[K][T + 1];                // As[k][i] = A[i0+i, k0+k], Bs[k][j] = op(B)[k0+k, j0+j]
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // micro-tile rows 4*tx.., columns 4*ty..
    const int i0 = blockIdx.x * T, j0 = blockIdx.y * T;
    double acc[4][4] = {};
    for (int k0 = 0; k0 < n; k0 += K) {
        for (int e = threadIdx.x; e < T * K; e += 256) {
            const int i = e % T, k = e / T;                      // consecutive threads walk a column of A
            As[k][i] = (i0 + i < n && k0 + k < n) ? A[(size_t)(k0 + k) * n + i0 + i] : 0.0;
        }
        for (int e = threadIdx.x; e < T * K; e += 256) {
            if (TB) {                                            // op(B)[k, j] = B[j, k]: walk a column of B over j
                const int j = e % T, k = e / T;
                Bs[k][j] = (j0 + j < n && k0 + k < n) ? B[(size_t)(k0 + k) * n + j0 + j] : 0.0;
            } else {
                const int k = e % K, j = e / K;
                Bs[k][j] = (j0 + j < n && k0 + k < n) ? B[(size_t)(j0 + j) * n + k0 + k] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double a[4], b[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) { a[r] = As[k][4 * tx + r]; b[r] = Bs[k][4 * ty + r]; }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = i0 + 4 * tx + r, j = j0 + 4 * ty + c;
            if (i < n && j < n) C[(size_t)j * n + i] = acc[r][c];
        }
}

typedef std::vector<double> Mat;                                 // n x n column-major (or a vector)

struct Work {                                                    // device scratch of one call
    cudaStream_t stream;
    double *dA = nullptr, *dB = nullptr, *dC = nullptr;
    size_t cap = 0;
    cudaError_t err = cudaSuccess;
    long gemms_on_device = 0;
    ~Work() { if (dA) cudaFree(dA); if (dB) cudaFree(dB); if (dC) cudaFree(dC); }
    bool reserve(size_t n2)
    {
        if (n2 <= cap) return true;
        if (dA) cudaFree(dA); if (dB) cudaFree(dB); if (dC) cudaFree(dC);
        dA = dB = dC = nullptr; cap = 0;
        if ((err = cudaMalloc(&dA, n2 * 8)) != cudaSuccess) return false;
        if ((err = cudaMalloc(&dB, n2 * 8)) != cudaSuccess) return false;
        if ((err = cudaMalloc(&dC, n2 * 8)) != cudaSuccess) return false;
        cap = n2;
        return true;
    }
};

#ifndef SDEGPU_LINLAWS_DEVICE_MIN
#define SDEGPU_LINLAWS_DEVICE_MIN 96                              // the CPU-only test build raises this so nothing touches CUDA
#endif
static const int DEVICE_GEMM_MIN = SDEGPU_LINLAWS_DEVICE_MIN;

// C = A op(B)
static bool gemm(Work &w, const Mat &A, const Mat &B, Mat &C, int n, bool transB = false)
{
    C.assign((size_t)n * n, 0.0);
    if (n < DEVICE_GEMM_MIN) {
        for (int j = 0; j < n; ++j)
            for (int k = 0; k < n; ++k) {
                const double b = transB ? B[(size_t)k * n + j] : B[(size_t)j * n + k];
                if (b == 0.0) continue;
                const double *a = &A[(size_t)k * n];
                double *c = &C[(size_t)j * n];
                for (int i = 0; i < n; ++i) c[i] = fma(a[i], b, c[i]);
            }
        return true;
    }
    const size_t n2 = (size_t)n * n;
    if (!w.reserve(n2)) return false;
    cudaMemcpyAsync(w.dA, A.data(), n2 * 8, cudaMemcpyHostToDevice, w.stream);
    cudaMemcpyAsync(w.dB, B.data(), n2 * 8, cudaMemcpyHostToDevice, w.stream);
    const dim3 grid((n + 63) / 64, (n + 63) / 64);
    if (transB) k_dgemm<true><<<grid, 256, 0, w.stream>>>(w.dA, w.dB, w.dC, n);
    else k_dgemm<false><<<grid, 256, 0, w.stream>>>(w.dA, w.dB, w.dC, n);
    cudaMemcpyAsync(C.data(), w.dC, n2 * 8, cudaMemcpyDeviceToHost, w.stream);
    w.err = cudaStreamSynchronize(w.stream);
    if (w.err == cudaSuccess) w.err = cudaGetLastError();
    ++w.gemms_on_device;
    return w.err == cudaSuccess;
}

// LU with partial pivoting in place; false when a pivot vanishes against the matrix' scale
static bool lu_factor(Mat &M, int n, std::vector<int> &piv)
{
    piv.resize((size_t)n);
    double scale = 0.0;
    for (double v : M) scale = fmax(scale, fabs(v));
    if (!(scale > 0.0) || !isfinite(scale)) return false;
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(M[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i)
            if (fabs(M[(size_t)k * n + i]) > best) { best = fabs(M[(size_t)k * n + i]); p = i; }
        if (!(best > 2.220446049250313e-16 * scale)) return false;
        piv[k] = p;
        if (p != k)
            for (int j = 0; j < n; ++j) std::swap(M[(size_t)j * n + k], M[(size_t)j * n + p]);
        const double inv = 1.0 / M[(size_t)k * n + k];
        for (int i = k + 1; i < n; ++i) M[(size_t)k * n + i] *= inv;
        auto update = [&](int j0, int j1) {
            for (int j = j0; j < j1; ++j) {
                const double f = M[(size_t)j * n + k];
                if (f == 0.0) continue;
                const double *l = &M[(size_t)k * n];
                double *c = &M[(size_t)j * n];
                for (int i = k + 1; i < n; ++i) c[i] = fma(-l[i], f, c[i]);
            }
        };
        const int cols = n - k - 1;
        if ((size_t)cols * cols > ((size_t)1 << 18)) {           // trailing update across host threads
            const unsigned nth = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nth; ++t) {
                const int a = k + 1 + (int)((size_t)cols * t / nth), b = k + 1 + (int)((size_t)cols * (t + 1) / nth);
                th.emplace_back(update, a, b);
            }
            for (auto &t : th) t.join();
        } else update(k + 1, n);
    }
    return true;
}

// solves (LU) X = Bm in place for nrhs right-hand sides (columns of Bm), threaded over columns
static void lu_solve(const Mat &LU, const std::vector<int> &piv, int n, double *Bm, int nrhs)
{
    auto cols = [&](int j0, int j1) {
        for (int j = j0; j < j1; ++j) {
            double *b = Bm + (size_t)j * n;
            for (int k = 0; k < n; ++k) std::swap(b[k], b[piv[k]]);
            for (int k = 0; k < n; ++k) {
                const double f = b[k];
                if (f == 0.0) continue;
                const double *l = &LU[(size_t)k * n];
                for (int i = k + 1; i < n; ++i) b[i] = fma(-l[i], f, b[i]);
            }
            for (int k = n - 1; k >= 0; --k) {
                b[k] /= LU[(size_t)k * n + k];
                const double f = b[k];
                const double *u = &LU[(size_t)k * n];
                for (int i = 0; i < k; ++i) b[i] = fma(-u[i], f, b[i]);
            }
        }
    };
    if ((size_t)n * n * nrhs > ((size_t)1 << 22) && nrhs > 1) {
        const unsigned nth = std::min({16u, std::max(1u, std::thread::hardware_concurrency()), (unsigned)nrhs});
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nth; ++t) th.emplace_back(cols, (int)((size_t)nrhs * t / nth), (int)((size_t)nrhs * (t + 1) / nth));
        for (auto &t : th) t.join();
    } else cols(0, nrhs);
}

static Mat transpose(const Mat &A, int n)
{
    Mat T((size_t)n * n);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) T[(size_t)i * n + j] = A[(size_t)j * n + i];
    return T;
}

// expm(M) by scaling and squaring, [13/13] Pade (Higham, SIAM J. Matrix Anal. Appl. 26 (2005), Algorithm 2.3)
static int expm(Work &w, const Mat &Min, int n, Mat &E)
{
    static const double b[14] = {64764752532480000.0, 32382376266240000.0, 7771770303897600.0, 1187353796428800.0, 129060195264000.0,
                                 10559470521600.0, 670442572800.0, 33522128640.0, 1323241920.0, 40840800.0, 960960.0, 16380.0, 182.0, 1.0};
    const size_t n2 = (size_t)n * n;
    double norm1 = 0.0;
    for (int j = 0; j < n; ++j) {
        double c = 0.0;
        for (int i = 0; i < n; ++i) c += fabs(Min[(size_t)j * n + i]);
        norm1 = fmax(norm1, c);
    }
    if (!isfinite(norm1)) return SDEGPU_ERR_ARG;
    int s = 0;
    if (norm1 > 5.371920351148152) s = (int)ceil(log2(norm1 / 5.371920351148152));
    Mat A(Min);
    const double sc = ldexp(1.0, -s);
    for (double &v : A) v *= sc;
    Mat A2, A4, A6, T1(n2), T2, U, V(n2);
    if (!gemm(w, A, A, A2, n) || !gemm(w, A2, A2, A4, n) || !gemm(w, A2, A4, A6, n)) return SDEGPU_ERR_CUDA;
    for (size_t e = 0; e < n2; ++e) T1[e] = b[13] * A6[e] + b[11] * A4[e] + b[9] * A2[e];
    if (!gemm(w, A6, T1, T2, n)) return SDEGPU_ERR_CUDA;
    for (size_t e = 0; e < n2; ++e) T2[e] += b[7] * A6[e] + b[5] * A4[e] + b[3] * A2[e];
    for (int i = 0; i < n; ++i) T2[(size_t)i * n + i] += b[1];
    if (!gemm(w, A, T2, U, n)) return SDEGPU_ERR_CUDA;
    for (size_t e = 0; e < n2; ++e) T1[e] = b[12] * A6[e] + b[10] * A4[e] + b[8] * A2[e];
    if (!gemm(w, A6, T1, T2, n)) return SDEGPU_ERR_CUDA;
    for (size_t e = 0; e < n2; ++e) V[e] = T2[e] + b[6] * A6[e] + b[4] * A4[e] + b[2] * A2[e];
    for (int i = 0; i < n; ++i) V[(size_t)i * n + i] += b[0];
    Mat P(n2);
    E.resize(n2);
    for (size_t e = 0; e < n2; ++e) { P[e] = V[e] - U[e]; E[e] = V[e] + U[e]; }
    std::vector<int> piv;
    if (!lu_factor(P, n, piv)) return SDEGPU_ERR_ARG;
    lu_solve(P, piv, n, E.data(), n);
    for (int k = 0; k < s; ++k) {
        if (!gemm(w, E, E, T1, n)) return SDEGPU_ERR_CUDA;
        E.swap(T1);
    }
    return SDEGPU_OK;
}

// the reference's method: P = kronecker(I,A) + kronecker(A,I), X = -solve(P, vec(Q))   (:596-599)
static int lyap_kronecker(const Mat &A, const Mat &Q, int d, Mat &X)
{
    const int n = d * d;
    Mat P((size_t)n * n, 0.0);
    for (int i = 0; i < d; ++i)
        for (int a = 0; a < d; ++a)
            for (int bb = 0; bb < d; ++bb) {
                P[(size_t)(i * d + bb) * n + (i * d + a)] += A[(size_t)bb * d + a];          // I (x) A: block (i,i) = A
                P[(size_t)(bb * d + i) * n + (a * d + i)] += A[(size_t)bb * d + a];          // A (x) I: block (a,b) = A[a,b] I
            }
    std::vector<int> piv;
    if (!lu_factor(P, n, piv)) return SDEGPU_ERR_ARG;
    X = Q;
    lu_solve(P, piv, n, X.data(), 1);
    for (double &v : X) v = -v;
    return SDEGPU_OK;
}

// Hurwitz A: bilinear transform + doubling.  Returns SDEGPU_ERR_ARG when the iteration does not contract (A not stable).
static int lyap_doubling(Work &w, const Mat &A, const Mat &Q, int d, Mat &X)
{
    const size_t n2 = (size_t)d * d;
    double fro = 0.0;
    for (double v : A) fro += v * v;
    const double g = sqrt(fro / d);
    if (!(g > 0.0) || !isfinite(g)) return SDEGPU_ERR_ARG;
    Mat M(A), Ad(A), W(Q);
    for (int i = 0; i < d; ++i) { M[(size_t)i * d + i] -= g; Ad[(size_t)i * d + i] += g; }
    std::vector<int> piv;
    if (!lu_factor(M, d, piv)) return SDEGPU_ERR_ARG;
    lu_solve(M, piv, d, Ad.data(), d);                           // A_d = (A - gI)^-1 (A + gI)
    lu_solve(M, piv, d, W.data(), d);                            // (A - gI)^-1 Q
    Mat Wt = transpose(W, d);
    lu_solve(M, piv, d, Wt.data(), d);                           // (A - gI)^-1 Q' (A - gI)^-T, transposed
    X = transpose(Wt, d);
    for (double &v : X) v *= 2.0 * g;
    Mat T, T2;
    double prev = INFINITY;
    for (int it = 0; it < 200; ++it) {
        if (!gemm(w, Ad, X, T, d) || !gemm(w, T, Ad, T2, d, true)) return SDEGPU_ERR_CUDA;
        double inc = 0.0, tot = 0.0;
        for (size_t e = 0; e < n2; ++e) { inc += T2[e] * T2[e]; X[e] += T2[e]; tot += X[e] * X[e]; }
        if (!isfinite(inc) || !isfinite(tot)) return SDEGPU_ERR_ARG;
        if (inc <= 1e-34 * tot || inc == 0.0) {
            for (int j = 0; j < d; ++j)
                for (int i = 0; i < j; ++i) {
                    const double s = 0.5 * (X[(size_t)j * d + i] + X[(size_t)i * d + j]);
                    X[(size_t)j * d + i] = X[(size_t)i * d + j] = s;
                }
            return SDEGPU_OK;
        }
        if (it > 60 && inc >= prev) return SDEGPU_ERR_ARG;       // not contracting: spectrum is not in the left half plane
        prev = inc;
        if (!gemm(w, Ad, Ad, T, d)) return SDEGPU_ERR_CUDA;
        Ad.swap(T);
    }
    return SDEGPU_ERR_ARG;
}

static const int KRONECKER_MAX = 40;

static int lyap_any(Work &w, const Mat &A, const Mat &Q, int d, Mat &X)
{
    if (d <= KRONECKER_MAX) return lyap_kronecker(A, Q, d, X);
    int rc = lyap_doubling(w, A, Q, d, X);
    if (rc == SDEGPU_ERR_ARG) {                                  // anti-stable A: (-A) X + X (-A)' + (-Q) = 0
        Mat nA(A), nQ(Q);
        for (double &v : nA) v = -v;
        for (double &v : nQ) v = -v;
        rc = lyap_doubling(w, nA, nQ, d, X);
    }
    return rc;
}

int linlaws_lyap(cudaStream_t stream, const double *A, const double *Q, int d, double *X, const char **why)
{
    Work w;
    w.stream = stream;
    Mat a(A, A + (size_t)d * d), q(Q, Q + (size_t)d * d), x;
    const int rc = lyap_any(w, a, q, d, x);
    if (rc == SDEGPU_ERR_ARG)
        *why = d <= KRONECKER_MAX ? "the Kronecker system is singular (A has eigenvalues with lambda_i + lambda_j = 0)"
                                  : "for d > 40 the doubling iteration needs all eigenvalues of A strictly on one side of the imaginary axis";
    else if (rc) *why = cudaGetErrorString(w.err);
    if (rc == SDEGPU_OK) memcpy(X, x.data(), sizeof(double) * (size_t)d * d);
    return rc;
}

int linlaws_dlinsde(cudaStream_t stream, const double *A, const double *G, int d, int nB, double t, const double *x0, const double *u,
                    int u_cols, const double *S0, double *eAt_out, double *EX, double *St, const char **why)
{
    Work w;
    w.stream = stream;
    const size_t n2 = (size_t)d * d;
    Mat a(A, A + n2), GG(n2, 0.0), s0(n2, 0.0);
    if (S0) s0.assign(S0, S0 + n2);
    for (int k = 0; k < nB; ++k)                                 // GG <- G %*% t(G)  (:513)
        for (int j = 0; j < d; ++j)
            for (int i = 0; i < d; ++i) GG[(size_t)j * d + i] += G[(size_t)k * d + i] * G[(size_t)k * d + j];
    Mat At(a), eAt;
    for (double &v : At) v *= t;
    int rc = expm(w, At, d, eAt);                                // eAt <- expm(A*t)  (:518)
    if (rc) { *why = rc == SDEGPU_ERR_ARG ? "expm: singular Pade denominator or non-finite input" : cudaGetErrorString(w.err); return rc; }
    Mat Sinf, st(n2);
    rc = lyap_any(w, a, GG, d, Sinf);
    if (rc == SDEGPU_ERR_CUDA) { *why = cudaGetErrorString(w.err); return rc; }
    if (rc == SDEGPU_OK) {                                       // St <- Sinf - eAt %*% (Sinf - S0) %*% t(eAt)  (:522-523)
        Mat D(n2), T, T2;
        for (size_t e = 0; e < n2; ++e) D[e] = Sinf[e] - s0[e];
        if (!gemm(w, eAt, D, T, d) || !gemm(w, T, eAt, T2, d, true)) { *why = cudaGetErrorString(w.err); return SDEGPU_ERR_CUDA; }
        for (size_t e = 0; e < n2; ++e) st[e] = Sinf[e] - T2[e];
    } else {                                                     // no stationary variance: Van Loan's block exponential (:525-533)
        const int m = 2 * d;
        Mat H((size_t)m * m, 0.0), eH;
        for (int j = 0; j < d; ++j)
            for (int i = 0; i < d; ++i) {
                H[(size_t)j * m + i] = -a[(size_t)i * d + j] * t;                 // -t(A)
                H[(size_t)j * m + d + i] = GG[(size_t)j * d + i] * t;             // GG
                H[(size_t)(d + j) * m + d + i] = a[(size_t)j * d + i] * t;        // A
            }
        rc = expm(w, H, m, eH);
        if (rc) { *why = "expm of the Van Loan block matrix failed"; return rc; }
        Mat Xm(n2), Ym(n2);                                      // XY <- eH %*% rbind(I, S0)
        for (int j = 0; j < d; ++j)
            for (int i = 0; i < d; ++i) {
                double xv = eH[(size_t)j * m + i], yv = eH[(size_t)j * m + d + i];
                for (int k = 0; k < d; ++k) {
                    xv = fma(eH[(size_t)(d + k) * m + i], s0[(size_t)j * d + k], xv);
                    yv = fma(eH[(size_t)(d + k) * m + d + i], s0[(size_t)j * d + k], yv);
                }
                Xm[(size_t)j * d + i] = xv;
                Ym[(size_t)j * d + i] = yv;
            }
        // St <- Y %*% solve(X):  St' = solve(X', Y')
        Mat Xt = transpose(Xm, d), Yt = transpose(Ym, d);
        std::vector<int> piv;
        if (!lu_factor(Xt, d, piv)) { *why = "dLinSDE: the Van Loan factor X is singular"; return SDEGPU_ERR_ARG; }
        lu_solve(Xt, piv, d, Yt.data(), d);
        st = transpose(Yt, d);
    }
    memcpy(St, st.data(), sizeof(double) * n2);
    if (!x0) {                                                   // list(eAt=, St=)  (:536-538)
        if (eAt_out) memcpy(eAt_out, eAt.data(), sizeof(double) * n2);
        return SDEGPU_OK;
    }
    if (!EX) { *why = "dLinSDE: EX is NULL although x0 is given"; return SDEGPU_ERR_ARG; }
    if (!u) {                                                    // EX = eAt %*% x0  (:540-542)
        for (int i = 0; i < d; ++i) {
            double s = 0.0;
            for (int k = 0; k < d; ++k) s = fma(eAt[(size_t)k * d + i], x0[k], s);
            EX[i] = s;
        }
        return SDEGPU_OK;
    }
    // zero-order hold: extended state (X,U), AA = rbind(cbind(A,I), 0)  (:545-553)
    // first-order hold: extended state (dU/dt, U, X), AAA = rbind(0, cbind(I,0,0), cbind(0,I,A))  (:559-571)
    const int blocks = u_cols == 1 ? 2 : 3, m = blocks * d;
    Mat AA((size_t)m * m, 0.0), eAA;
    std::vector<double> v((size_t)m, 0.0);
    if (u_cols == 1) {
        for (int j = 0; j < d; ++j) {
            for (int i = 0; i < d; ++i) AA[(size_t)j * m + i] = a[(size_t)j * d + i] * t;
            AA[(size_t)(d + j) * m + j] = t;
        }
        for (int i = 0; i < d; ++i) { v[i] = x0[i]; v[d + i] = u[i]; }
    } else {
        for (int j = 0; j < d; ++j) {
            AA[(size_t)j * m + d + j] = t;                       // row block 2, column block 1: I
            AA[(size_t)(d + j) * m + 2 * d + j] = t;             // row block 3, column block 2: I
            for (int i = 0; i < d; ++i) AA[(size_t)(2 * d + j) * m + 2 * d + i] = a[(size_t)j * d + i] * t;
        }
        for (int i = 0; i < d; ++i) { v[i] = (u[d + i] - u[i]) / t; v[d + i] = u[i]; v[2 * d + i] = x0[i]; }
    }
    rc = expm(w, AA, m, eAA);
    if (rc) { *why = "expm of the extended (state, input) system failed"; return rc; }
    const int row0 = u_cols == 1 ? 0 : 2 * d;
    for (int i = 0; i < d; ++i) {
        double s = 0.0;
        for (int k = 0; k < m; ++k) s = fma(eAA[(size_t)k * m + row0 + i], v[k], s);
        EX[i] = s;
    }
    return SDEGPU_OK;
}

}  // namespace sdegpu
