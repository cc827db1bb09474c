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

// One time step of one path, state in registers.
template <class M, int SCHEME, int PROJ, class A>
__device__ __forceinline__ void sde_step(const KParams &P, double (&x)[M::NX], double dt, double dB)
{
    constexpr int NX = M::NX;
    double fX[NX], gX[NX];
    M::template f<A>(P, x, fX);
    M::template g<A>(P, x, gX);
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

// ---- FAST-arithmetic step kernels with per-step coefficients folded on the host ------------------
// For the register-resident fused-generator kernel the step is regrouped algebraically (allowed in FAST
// mode, never in STRICT): everything that depends only on (params, dt_i) is precomputed once per run
// into a 6-double row per step, so e.g. heun/CIR drops from 31 to 18 FP64 instructions:
//   Y  = p( s w + (A x + Bc) ),  s = sqrt|x|, w = C z
//   X' = p( hw (s + sY) + (A2 x + Bc - H Y) ),  sY = sqrt|Y|, hw = C2 z
// with A = 1-lambda dt, Bc = lambda xi dt, C = gamma sqrt(dt), C2 = C/2, A2 = 1-lambda dt/2, H = lambda dt/2.
constexpr int FAST_NCOEF = 6;

template <class M, int SCHEME, int PROJ> struct FastStep {
    static constexpr bool uses_coef = false;
    static __device__ __forceinline__ void apply(const KParams &P, double (&x)[M::NX], double2 d, const double *, double z)
    { sde_step<M, SCHEME, PROJ, Fast>(P, x, d.x, d.y * z); }
};

template <int SCHEME, int PROJ> struct FastStep<Model<SDEGPU_MODEL_CIR>, SCHEME, PROJ> {
    static constexpr bool uses_coef = true;
    static __device__ __forceinline__ void apply(const KParams &, double (&x)[1], double2, const double *c, double z)
    {
        const double2 ab = __ldg(reinterpret_cast<const double2 *>(c)), cc = __ldg(reinterpret_cast<const double2 *>(c + 2));
        const double s = Fast::sqrt_(fabs(x[0])), w = cc.x * z;
        const double Y = project<PROJ>(fma(s, w, fma(ab.x, x[0], ab.y)));
        if (SCHEME == SDEGPU_EULER) { x[0] = Y; return; }
        const double2 ah = __ldg(reinterpret_cast<const double2 *>(c + 4));
        const double sY = Fast::sqrt_(fabs(Y)), hw = cc.y * z;
        x[0] = project<PROJ>(fma(hw, s + sY, fma(-ah.y, Y, fma(ah.x, x[0], ab.y))));
    }
};

template <int SCHEME, int PROJ> struct FastStep<Model<SDEGPU_MODEL_OU>, SCHEME, PROJ> {
    static constexpr bool uses_coef = true;
    static __device__ __forceinline__ void apply(const KParams &, double (&x)[1], double2, const double *c, double z)
    {
        const double2 ab = __ldg(reinterpret_cast<const double2 *>(c)), cc = __ldg(reinterpret_cast<const double2 *>(c + 2));
        const double w = cc.x * z;                                // sigma sqrt(dt) z
        const double Y = project<PROJ>(w + fma(ab.x, x[0], ab.y));
        if (SCHEME == SDEGPU_EULER) { x[0] = Y; return; }
        const double2 ah = __ldg(reinterpret_cast<const double2 *>(c + 4));
        x[0] = project<PROJ>(w + fma(-ah.y, Y, fma(ah.x, x[0], ab.y)));      // g constant: 0.5*(g+g) = g
    }
};

// host side: the coefficient row of one step (returns false when the model has no folded form)
inline bool fast_step_coefficients(int model, const double *p, double dt, double sq, double *c)
{
    if (model == SDEGPU_MODEL_CIR || model == SDEGPU_MODEL_OU) {
        const double lam = p[0], xi = p[1], vol = p[2];
        c[0] = 1.0 - lam * dt; c[1] = lam * xi * dt; c[2] = vol * sq; c[3] = 0.5 * vol * sq;
        c[4] = 1.0 - 0.5 * lam * dt; c[5] = 0.5 * lam * dt;
        return true;
    }
    return false;
}

}  // namespace sdegpu
