// lyap() / dLinSDE() at device dimension (linlaws.cu).  Matrices are column-major host arrays; `why` receives a static
// explanation when the return code is not SDEGPU_OK.
#pragma once
#include <cuda_runtime.h>

namespace sdegpu {

int linlaws_lyap(cudaStream_t stream, const double *A, const double *Q, int d, double *X, const char **why);
int linlaws_dlinsde(cudaStream_t stream, const double *A, const double *G, int d, int nB, double t, const double *x0, const double *u,
                    int u_cols, const double *S0, double *eAt, double *EX, double *St, const char **why);

}  // namespace sdegpu
