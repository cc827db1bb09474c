// Model catalogue and the euler / heun step in reference operation order.
//   euler step  R/sdetools.R:450-453   X' = p( (X + fX*dt) + gX*dB )
//   heun step   R/sdetools.R:316-323   Y  = p( (X + fX*dt) + gX*dB )
//                                      X' = p( (X + (0.5*(fX+fY))*dt) + 0.5*((gX+gY)*dB) )
// All catalogue models below have a vector-valued g (one Brownian channel, :422-424).
#pragma once
#include "../../include/sdegpu.h"

namespace sdegpu {

struct KParams { double p[4]; };

// STRICT: every R operator is one correctly rounded op, never contracted.
struct Strict {
    static constexpr bool strict = true;
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double sqrt_(double a) { return __dsqrt_rn(a); }
};
// FAST: plain expressions, nvcc contracts a*b+c into DFMA.
struct Fast {
    static constexpr bool strict = false;
    static __device__ __forceinline__ double add(double a, double b) { return a + b; }
    static __device__ __forceinline__ double sub(double a, double b) { return a - b; }
    static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
    // sqrt of a finite a >= 0 without the special-case branch of the IEEE routine: MUFU.RSQ64H seed y
    // (relative error e ~ 2^-22) and one third-order step  sqrt(a) = a*y*(1 + e/2 + 3e^2/8 + O(e^3)),
    // e = 1 - a*y^2, i.e. 5 FP64 instructions and <= 1-2 ulp.  The high word is clamped to the smallest
    // normal exponent with an integer max so the seed stays finite at a = 0 (sqrt(0) -> 2^-511).
    static __device__ __forceinline__ double sqrt_(double a)
    {
        a = __hiloint2double(max(__double2hiint(a), 0x00100000), __double2loint(a));
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
        const double t = a * y;
        const double e = fma(-t, y, 1.0);
        const double p = fma(0.375, e, 0.5);
        return fma(t, p * e, t);
    }
};

template <int MODEL> struct Model;

template <> struct Model<SDEGPU_MODEL_OU> {            // R/sdetools.R:718; Killing.Rmd:157-161
    static constexpr int NX = 1, NPARAMS = 3;
    template <class A> static __device__ __forceinline__ void f(const KParams &P, const double *x, double *o)
    { o[0] = A::mul(P.p[0], A::sub(P.p[1], x[0])); }
    template <class A> static __device__ __forceinline__ void g(const KParams &P, const double *, double *o) { o[0] = P.p[2]; }
};
template <> struct Model<SDEGPU_MODEL_CIR> {           // HMM-filtering-of-scalar-SDEs.Rmd:33-51
    static constexpr int NX = 1, NPARAMS = 3;
    template <class A> static __device__ __forceinline__ void f(const KParams &P, const double *x, double *o)
    { o[0] = A::mul(P.p[0], A::sub(P.p[1], x[0])); }
    template <class A> static __device__ __forceinline__ void g(const KParams &P, const double *x, double *o)
    { o[0] = A::mul(P.p[2], A::sqrt_(fabs(x[0]))); }
};
template <> struct Model<SDEGPU_MODEL_VDP> {           // vanderPol.Rmd:27-53
    static constexpr int NX = 2, NPARAMS = 2;
    template <class A> static __device__ __forceinline__ void f(const KParams &P, const double *x, double *o)
    { o[0] = x[1]; o[1] = A::sub(A::mul(x[1], A::sub(P.p[0], A::mul(x[0], x[0]))), x[0]); }
    template <class A> static __device__ __forceinline__ void g(const KParams &P, const double *, double *o)
    { o[0] = 0.0; o[1] = P.p[1]; }
};
template <> struct Model<SDEGPU_MODEL_LOGISTIC> {      // FisheriesManagement.Rmd:30-51; ItosFormula.Rmd:38-45
    static constexpr int NX = 1, NPARAMS = 2;
    template <class A> static __device__ __forceinline__ void f(const KParams &P, const double *x, double *o)
    { o[0] = A::sub(A::mul(x[0], A::sub(1.0, x[0])), A::mul(P.p[1], A::mul(x[0], x[0]))); }
    template <class A> static __device__ __forceinline__ void g(const KParams &P, const double *x, double *o)
    { o[0] = A::mul(P.p[0], x[0]); }
};
template <> struct Model<SDEGPU_MODEL_GBM> {           // R/sdetools.R:222-228
    static constexpr int NX = 1, NPARAMS = 2;
    template <class A> static __device__ __forceinline__ void f(const KParams &P, const double *x, double *o)
    { o[0] = A::mul(P.p[0], x[0]); }
    template <class A> static __device__ __forceinline__ void g(const KParams &P, const double *x, double *o)
    { o[0] = A::mul(P.p[1], x[0]); }
};
template <> struct Model<SDEGPU_MODEL_DOUBLEWELL> {    // SDEtools.Rmd:50-55 (x^3 written x*x*x)
    static constexpr int NX = 1, NPARAMS = 1;
    template <class A> static __device__ __forceinline__ void f(const KParams &, const double *x, double *o)
    { o[0] = A::sub(x[0], A::mul(A::mul(x[0], x[0]), x[0])); }
    template <class A> static __device__ __forceinline__ void g(const KParams &P, const double *, double *o) { o[0] = P.p[0]; }
};

template <> struct Model<SDEGPU_MODEL_DWSQRT> {        // ItoIntegrals.Rmd:66-69: f = x - x^3, g = sqrt(1 + x^2)
    static constexpr int NX = 1, NPARAMS = 0;
    template <class A> static __device__ __forceinline__ void f(const KParams &, const double *x, double *o)
    { o[0] = A::sub(x[0], A::mul(A::mul(x[0], x[0]), x[0])); }
    template <class A> static __device__ __forceinline__ void g(const KParams &, const double *x, double *o)
    { o[0] = A::strict ? __dsqrt_rn(__dadd_rn(1.0, __dmul_rn(x[0], x[0]))) : sqrt(fma(x[0], x[0], 1.0)); }
};
template <> struct Model<SDEGPU_MODEL_LOGLOGISTIC> {   // ItosFormula.Rmd:113-116: Y = log X of logistic growth
    static constexpr int NX = 1, NPARAMS = 1;
    template <class A> static __device__ __forceinline__ void f(const KParams &P, const double *x, double *o)
    { o[0] = A::sub(A::sub(1.0, exp(x[0])), A::mul(0.5, A::mul(P.p[0], P.p[0]))); }
    template <class A> static __device__ __forceinline__ void g(const KParams &P, const double *, double *o) { o[0] = P.p[0]; }
};

template <int PROJ> __device__ __forceinline__ double project(double v)
{
    if (PROJ == SDEGPU_PROJ_ABS) return fabs(v);
    if (PROJ == SDEGPU_PROJ_RELU) return v < 0.0 ? 0.0 : v;
    return v;
}

// Time dependence of the closures, tabulated on the grid (sdegpu.h ABI 5): f(t,x) = f0(x) + c(t), g(t,x) = s(t) g0(x).
// cX/sX belong to t_i (predictor / Euler), cY/sY to t_{i+1} (Heun corrector evaluates ff(times[i+1], Y), R/sdetools.R:321-322).
template <int NX> struct TDep {
    double cX[NX], cY[NX], sX, sY;
    bool has_c, has_s;                                           // warp-uniform: which of the two tables exist
};

// One time step of one path, state in registers.
template <class M, int SCHEME, int PROJ, class A, bool TD = false>
__device__ __forceinline__ void sde_step(const KParams &P, double (&x)[M::NX], double dt, double dB, const TDep<M::NX> *td = nullptr)
{
    constexpr int NX = M::NX;
    double fX[NX], gX[NX];
    M::template f<A>(P, x, fX);
    M::template g<A>(P, x, gX);
    if (TD) {
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            if (td->has_c) fX[j] = A::add(fX[j], td->cX[j]);
            if (td->has_s) gX[j] = A::mul(td->sX, gX[j]);
        }
    }
    if (SCHEME == SDEGPU_EULER) {
#pragma unroll
        for (int j = 0; j < NX; ++j)
            x[j] = project<PROJ>(A::add(A::add(x[j], A::mul(fX[j], dt)), A::mul(gX[j], dB)));
    } else {
        double Y[NX], fY[NX], gY[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j)
            Y[j] = project<PROJ>(A::add(A::add(x[j], A::mul(fX[j], dt)), A::mul(gX[j], dB)));
        M::template f<A>(P, Y, fY);
        M::template g<A>(P, Y, gY);
        if (TD) {
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                if (td->has_c) fY[j] = A::add(fY[j], td->cY[j]);
                if (td->has_s) gY[j] = A::mul(td->sY, gY[j]);
            }
        }
        if (A::strict) {
#pragma unroll
            for (int j = 0; j < NX; ++j)
                x[j] = project<PROJ>(A::add(A::add(x[j], A::mul(A::mul(0.5, A::add(fX[j], fY[j])), dt)),
                                            A::mul(0.5, A::mul(A::add(gX[j], gY[j]), dB))));
        } else {
            const double hdt = 0.5 * dt, hdB = 0.5 * dB;     // scaling by 0.5 is exact: same value, fewer ops
#pragma unroll
            for (int j = 0; j < NX; ++j)
                x[j] = project<PROJ>((x[j] + (fX[j] + fY[j]) * hdt) + (gX[j] + gY[j]) * hdB);
        }
    }
}

// ---- FAST-arithmetic steps of the fused-generator kernels, per-step constants folded on the host ------------
// Everything that depends only on (params, dt_i) is precomputed once per run into a row of FAST_NCOEF doubles
// per step; row[0] is the factor that turns the standard normal into the noise term the step consumes.  On a
// uniform grid there is ONE row: it travels as a kernel parameter (constant-bank operands, no loads) and row[0]
// is folded into the staged quantile table, so the generator hands over row[0]*z directly.
//   generic   row = {sqrt(dt), dt}                          dB = row[0] z, sde_step<Fast>
//   OU        row = {sigma sqrt(dt), A, Bc, A2, H}          A = 1 - lambda dt, Bc = lambda xi dt, A2 = 1 - H, H = lambda dt/2
//             Y  = p( w + (A x + Bc) ),                     w = row[0] z
//             X' = p( w + (A2 x + Bc - H Y) )               (g constant: 0.5 (g + g) = g)
//   CIR       row = {gamma sqrt(dt)/2, A, Bc, A2, H}
//             Y  = p( s2 w + (A x + Bc) ),                  s2 = sqrt(4|x|) = 2 sqrt|x|, w = row[0] z
//             X' = p( w (s2/2 + sY) + (A2 x + Bc - H Y) ),  sY = sqrt|Y|          — 16 FP64 per Heun step
// (allowed in FAST mode, never in STRICT).
constexpr int FAST_NCOEF = 6;

// sqrt(k |x|) for finite x, k = 1 or 4, without the special cases of the IEEE routine: a = k|x| + 2^-1000 keeps the
// MUFU.RSQ64H seed y finite at x = 0 (one FP64 instruction, no integer clamp), then one second-order step.
//   ORDER 21 (default)  Heron's step on the residual: t = a y, r = a - t^2, sqrt(a) = t + r (y/2) with y/2 an exponent decrement
//                       of the seed's high word: 4 FP64 + 1 integer instruction per root, relative error ~ e^2/2 ~ 1e-14 (seed
//                       error e ~ 2^-22).  Measured on config C2: 4.53e11 -> 4.60e11 path-steps/s against ORDER 2.
//   ORDER 22            also forms a with integer operations on the high word: two FP64 instructions fewer, six integer ones more
//                       — slower (4.35e11); kept as the record of the experiment (profiles/r2_tune_terminal_sqrt.txt).
//   ORDER 2             sqrt(a) = t + t e / 2, e = 1 - t y: 5 FP64.      ORDER 3: the third-order step of Fast::sqrt_.
#ifndef SDEGPU_FASTSQRT_ORDER
#define SDEGPU_FASTSQRT_ORDER 21
#endif
template <int K> __device__ __forceinline__ double fast_sqrt_abs(double x)
{
#if SDEGPU_FASTSQRT_ORDER == 21 || SDEGPU_FASTSQRT_ORDER == 22
    // Heron's step on the residual: t = a y, r = a - t^2 (one FMA, exact product), sqrt(a) = t + r (y/2).  y/2 is an exponent
    // decrement of the seed's high word (its low word is zero), so the step is DMUL + 2 DFMA and one integer add: one FP64
    // instruction fewer than t + (t e)/2 with e = 1 - t y, same second-order error (~ e^2/2, e ~ 2^-22).
    // ORDER 22 also forms a = K|x| (kept away from zero) with integer operations on the high word instead of an FP64 add.
    double a;
    if (SDEGPU_FASTSQRT_ORDER == 22) {
        const int hi = __double2hiint(x) & 0x7FFFFFFF;
        // K = 4: two exponent steps up (x = 0 becomes 2^-1021); K = 1: at least the smallest normal exponent
        a = __hiloint2double(K == 4 ? hi + 0x00200000 : max(hi, 0x00100000), __double2loint(x));
    } else {
        const double tiny = 9.3326361850321888e-302;             // 2^-1000
        a = K == 1 ? fabs(x) + tiny : fma((double)K, fabs(x), tiny);
    }
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double t = a * y;
    const double r = fma(-t, t, a);
    const double h = __hiloint2double(__double2hiint(y) - 0x00100000, 0);
    return fma(r, h, t);
#else
    const double tiny = 9.3326361850321888e-302;                 // 2^-1000
    const double a = K == 1 ? fabs(x) + tiny : fma((double)K, fabs(x), tiny);
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double t = a * y;
    const double e = fma(-t, y, 1.0);
    if (SDEGPU_FASTSQRT_ORDER == 2) return fma(t * e, 0.5, t);
    const double p = fma(0.375, e, 0.5);
    return fma(t, p * e, t);
#endif
}

template <class M, int SCHEME, int PROJ> struct FastStep {
    template <class C> static __device__ __forceinline__ void apply(const KParams &P, double (&x)[M::NX], const C &c, double w)
    { sde_step<M, SCHEME, PROJ, Fast>(P, x, c[1], w); }
};

template <int SCHEME, int PROJ> struct FastStep<Model<SDEGPU_MODEL_CIR>, SCHEME, PROJ> {
    template <class C> static __device__ __forceinline__ void apply(const KParams &, double (&x)[1], const C &c, double w)
    {
        const double s2 = fast_sqrt_abs<4>(x[0]);
        const double Y = project<PROJ>(fma(s2, w, fma(c[1], x[0], c[2])));
        if (SCHEME == SDEGPU_EULER) { x[0] = Y; return; }
        const double sY = fast_sqrt_abs<1>(Y);
        const double q = fma(0.5, s2, sY);
        x[0] = project<PROJ>(fma(w, q, fma(-c[4], Y, fma(c[3], x[0], c[2]))));
    }
};

template <int SCHEME, int PROJ> struct FastStep<Model<SDEGPU_MODEL_OU>, SCHEME, PROJ> {
    template <class C> static __device__ __forceinline__ void apply(const KParams &, double (&x)[1], const C &c, double w)
    {
        const double Y = project<PROJ>(w + fma(c[1], x[0], c[2]));
        if (SCHEME == SDEGPU_EULER) { x[0] = Y; return; }
        x[0] = project<PROJ>(w + fma(-c[4], Y, fma(c[3], x[0], c[2])));
    }
};

// van der Pol: the recurrence's critical cycle is what a kernel with few warps per SM waits for, so the step is grouped for
// depth, not for operation count.  row = {sigma sqrt(dt), dt, 1 + mu dt, mu, dt/2}, w = row[0] z.
//   Euler   X0' = x0 + dt x1                                     (depth 1)
//           X1' = x1 (1 + mu dt - dt x0^2) + (w - dt x0)          (depth 3 from x0, 1 from x1)
//   Heun    the predictor as written in the reference, then the trapezoid with g constant (0.5 (g + g) dB = w)
template <int SCHEME, int PROJ> struct FastStep<Model<SDEGPU_MODEL_VDP>, SCHEME, PROJ> {
    template <class C> static __device__ __forceinline__ void apply(const KParams &, double (&x)[2], const C &c, double w)
    {
        if (SCHEME == SDEGPU_EULER) {
            const double t = c[1] * x[0];
            const double a = fma(-t, x[0], c[2]);
            const double b = fma(-c[1], x[0], w);
            const double n0 = fma(x[1], c[1], x[0]);
            x[1] = project<PROJ>(fma(x[1], a, b));
            x[0] = project<PROJ>(n0);
            return;
        }
        const double xw = x[1] + w;
        const double u = fma(x[1], fma(-x[0], x[0], c[3]), -x[0]);
        const double Y0 = project<PROJ>(fma(x[1], c[1], x[0])), Y1 = project<PROJ>(fma(u, c[1], xw));
        const double uY = fma(Y1, fma(-Y0, Y0, c[3]), -Y0);
        x[0] = project<PROJ>(fma(c[4], x[1] + Y1, x[0]));
        x[1] = project<PROJ>(fma(c[4], u + uY, xw));
    }
};

// host side: the coefficient row of one step
inline void fast_step_coefficients(int model, const double *p, double dt, double sq, double *c)
{
    for (int k = 0; k < FAST_NCOEF; ++k) c[k] = 0.0;
    if (model == SDEGPU_MODEL_CIR || model == SDEGPU_MODEL_OU) {
        const double lam = p[0], xi = p[1], vol = p[2];
        c[0] = model == SDEGPU_MODEL_CIR ? 0.5 * vol * sq : vol * sq;
        c[1] = 1.0 - lam * dt; c[2] = lam * xi * dt; c[3] = 1.0 - 0.5 * lam * dt; c[4] = 0.5 * lam * dt;
    } else if (model == SDEGPU_MODEL_VDP) {
        c[0] = p[1] * sq; c[1] = dt; c[2] = 1.0 + p[0] * dt; c[3] = p[0]; c[4] = 0.5 * dt;
    } else {
        c[0] = sq; c[1] = dt;
    }
}

}  // namespace sdegpu
