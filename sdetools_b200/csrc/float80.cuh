// Integer emulation of the x87 80-bit accumulator behind R's cumsum().
//
// The reference integrals (itointegral R/sdetools.R:155-158, stochint :105-113,
// QuadraticVariation :134-137, CrossVariation :184-187) are all c(0,cumsum(...)),
// and R's cumsum keeps an x87 `long double` running sum that is rounded to double
// once per element (R src/main/cum.c: `LDOUBLE sum; sum += x[i]; s[i]=(double)sum`).
// The GPU has no 80-bit type, so the accumulator is held as (sign, exponent,
// 64-bit significand) and updated with exact integer arithmetic + round-to-
// nearest-even at 64 bits: results are bit-identical to x86-64, not "within an ulp".
//
// Compiles for host too (tests build it with g++ and compare against the FPU).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SDEGPU_HD __host__ __device__ __forceinline__
#else
#define SDEGPU_HD inline
#endif

namespace sdegpu {

SDEGPU_HD int f80_clz64(uint64_t v)
{
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}

SDEGPU_HD uint64_t f80_dbits(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}

SDEGPU_HD double f80_bitsd(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}

// value = (-1)^sign * mant * 2^(exp-63), mant has bit 63 set unless the value is 0.
// `special` != 0 means a NaN/Inf entered the sum; the value then lives in `sp`.
struct F80 {
    uint64_t mant;
    int32_t exp;
    uint32_t sign;
    uint32_t special;
    double sp;
};

SDEGPU_HD F80 f80_zero()
{
    F80 r; r.mant = 0; r.exp = 0; r.sign = 0; r.special = 0; r.sp = 0.0; return r;
}

// exact conversion double -> extended (finite inputs)
SDEGPU_HD void f80_from_double(double d, uint64_t &m, int32_t &e, uint32_t &s)
{
    const uint64_t b = f80_dbits(d);
    s = (uint32_t)(b >> 63);
    const int32_t be = (int32_t)((b >> 52) & 0x7FF);
    const uint64_t frac = b & 0x000FFFFFFFFFFFFFull;
    if (be == 0) {
        if (frac == 0) { m = 0; e = 0; return; }
        const int lz = f80_clz64(frac);          // subnormal double: normal in extended
        m = frac << lz;
        e = -1022 - 52 + (63 - lz);
        return;
    }
    m = (frac | 0x0010000000000000ull) << 11;
    e = be - 1023;
}

// acc += d   (x87 FADD, precision control = extended, rounding = nearest-even)
SDEGPU_HD void f80_add(F80 &acc, double d)
{
    const uint64_t db = f80_dbits(d);
    if (acc.special || ((db >> 52) & 0x7FF) == 0x7FF) {       // Inf/NaN: value class follows IEEE
        double cur;
        if (acc.special) cur = acc.sp;
        else {                                                 // finite accumulator: only its sign/zero-ness matters
            cur = acc.mant == 0 ? 0.0 : (acc.sign ? -1.0 : 1.0);
        }
        acc.sp = cur + d;
        acc.special = 1;
        return;
    }
    uint64_t mb; int32_t eb; uint32_t sb;
    f80_from_double(d, mb, eb, sb);
    if (mb == 0) {                                             // x + (+-0)
        if (acc.mant == 0) acc.sign &= sb;
        return;
    }
    if (acc.mant == 0) { acc.mant = mb; acc.exp = eb; acc.sign = sb; return; }

    uint64_t ma = acc.mant; int32_t ea = acc.exp; uint32_t sa = acc.sign;
    if (ea < eb || (ea == eb && ma < mb)) {                    // make |a| >= |b|
        uint64_t tm = ma; ma = mb; mb = tm;
        int32_t te = ea; ea = eb; eb = te;
        uint32_t ts = sa; sa = sb; sb = ts;
    }
    const uint32_t diff = (uint32_t)(ea - eb);
    // b aligned into a 128-bit window below a's binary point: (bh:bl), sticky = bits lost below
    uint64_t bh, bl; uint32_t sticky = 0;
    if (diff == 0) { bh = mb; bl = 0; }
    else if (diff < 64) { bh = mb >> diff; bl = mb << (64 - diff); }
    else if (diff == 64) { bh = 0; bl = mb; }
    else if (diff < 128) { bh = 0; bl = mb >> (diff - 64); sticky = (mb << (128 - diff)) != 0; }
    else { bh = 0; bl = 0; sticky = 1; }

    uint64_t rh, rl; int32_t er = ea;
    if (sa == sb) {
        rl = bl;                                               // a's low word is 0
        rh = ma + bh;
        if (rh < ma) {                                         // carry out: shift right one
            sticky |= (uint32_t)(rl & 1);
            rl = (rl >> 1) | (rh << 63);
            rh = (rh >> 1) | 0x8000000000000000ull;
            er += 1;
        }
    } else {
        // a - b - (sticky ? epsilon : 0): borrow one unit of the 128-bit window when bits were lost
        uint64_t sl = bl, sh = bh;
        if (sticky) { sl += 1; if (sl == 0) sh += 1; }
        rl = 0 - sl;
        rh = ma - sh - (sl != 0 ? 1 : 0);
        if ((rh | rl) == 0) {                                  // exact cancellation -> +0 (nearest-even mode)
            acc.mant = 0; acc.exp = 0; acc.sign = 0; return;
        }
        if (rh == 0) { rh = rl; rl = 0; er -= 64; }
        const int lz = f80_clz64(rh);
        if (lz) { rh = (rh << lz) | (rl >> (64 - lz)); rl <<= lz; er -= lz; }
    }
    // round to nearest even at 64 bits
    const uint32_t round_bit = (uint32_t)(rl >> 63);
    const uint32_t below = ((rl << 1) != 0) | sticky;
    if (round_bit && (below || (rh & 1))) {
        rh += 1;
        if (rh == 0) { rh = 0x8000000000000000ull; er += 1; }
    }
    acc.mant = rh; acc.exp = er; acc.sign = sa;
}

// (double) acc   — FST to a 64-bit slot: round-to-nearest-even to 53 bits
SDEGPU_HD double f80_to_double(const F80 &acc)
{
    if (acc.special) return acc.sp;
    const uint64_t sbit = (uint64_t)acc.sign << 63;
    if (acc.mant == 0) return f80_bitsd(sbit);
    int32_t e = acc.exp;
    if (e > 1023) return f80_bitsd(sbit | 0x7FF0000000000000ull);
    int shift = 11;
    uint64_t expfield;
    if (e < -1022) { shift += (-1022 - e); expfield = 0; }
    else expfield = (uint64_t)(e + 1022);                      // hidden bit adds the last +1
    if (shift > 64) return f80_bitsd(sbit);
    uint64_t q, rem_top, rem_rest;
    if (shift == 64) { q = 0; rem_top = acc.mant >> 63; rem_rest = (acc.mant << 1) != 0; }
    else {
        q = acc.mant >> shift;
        rem_top = (acc.mant >> (shift - 1)) & 1;
        rem_rest = (acc.mant & ((1ull << (shift - 1)) - 1)) != 0;
    }
    uint64_t bits = (expfield << 52) + q;
    if (rem_top && (rem_rest || (q & 1))) bits += 1;           // carries ripple into the exponent correctly
    if (bits >= 0x7FF0000000000000ull) bits = 0x7FF0000000000000ull;
    return f80_bitsd(sbit | bits);
}

}  // namespace sdegpu
