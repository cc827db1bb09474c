// Test-only C entry points over csrc/linlaws.cu so that its host logic (Pade expm, Kronecker and doubling Lyapunov
// solvers, Van Loan, input holds) can be exercised on a box without a GPU: below d = 96 no CUDA call is made.
#include <cuda_runtime.h>

#include "../../sdetools_b200/csrc/linlaws.cuh"

extern "C" {
const char *g_why = "";
int t_lyap(const double *A, const double *Q, int d, double *X) { return sdegpu::linlaws_lyap(nullptr, A, Q, d, X, &g_why); }
int t_dlinsde(const double *A, const double *G, int d, int nB, double t, const double *x0, const double *u, int u_cols, const double *S0,
              double *eAt, double *EX, double *St)
{
    return sdegpu::linlaws_dlinsde(nullptr, A, G, d, nB, t, x0, u, u_cols, S0, eAt, EX, St, &g_why);
}
const char *t_why(void) { return g_why; }
}
