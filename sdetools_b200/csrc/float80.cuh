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


// ---- floating-point fast path ------------------------------------------------------------------
// The same accumulator as a pair of doubles: S = h + l with h = RN53(S), so that (double)S — what
// cumsum stores — is h itself.  S has a 64-bit significand, hence l (at most 11 bits) is exact.
// S' = RN64(S + x) is formed without integer arithmetic on the significand:
//   two TwoSums and two FastTwoSums give V = S + x = a + b + c exactly (|b| <= ulp(a)/2, |c| <= ulp(b)/2);
//   the 64-bit grid q = 2^(e-63) follows from a's exponent (one binade lower when a is a power of
//   two and the remainder points back towards zero); b is rounded to that grid, ties to even, with
//   the add-and-subtract constant 1.5 * 2^52 * q, c only ever deciding an exact tie;
//   S' = a + r is split back into (h, l) by FastTwoSum.
// Every operation is an individually rounded IEEE add/subtract (no products, nothing to contract).
// Returns false — state untouched — when the case is left to the integer path: non-finite values,
// magnitudes whose 64-bit grid or sum leaves the double range (|.| < 2^-895 or >= 2^992).
// tests/native/float80_check.cpp compares it with f80_add and with the x87 FPU.
#ifndef SDEGPU_F80P_COUNT
#define SDEGPU_F80P_COUNT(k)          // host test hook: counts tie-broken / lower-binade / exact-tie cases
#endif

// high / low 32-bit words of a double and back (register-pair moves on the device, no 64-bit shifts)
SDEGPU_HD uint32_t f80_hi(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2hiint(d);
#else
    return (uint32_t)(f80_dbits(d) >> 32);
#endif
}
SDEGPU_HD uint32_t f80_lo(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(d);
#else
    return (uint32_t)f80_dbits(d);
#endif
}
SDEGPU_HD double f80_from_hi(uint32_t hi)                        // the double with that high word and a zero low word
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double((int)hi, 0);
#else
    return f80_bitsd((uint64_t)hi << 32);
#endif
}

SDEGPU_HD void f80p_two_sum(double x, double y, double &s, double &e)
{
    s = x + y;
    const double t = s - x;
    e = (x - (s - t)) + (y - t);
}

SDEGPU_HD bool f80p_add(double &h, double &l, double x)
{
    const uint32_t xh = f80_hi(x), hh = f80_hi(h);
    const uint32_t xe = (xh >> 20) & 0x7FFu, he = (hh >> 20) & 0x7FFu;
    if (xe >= 0x7DFu || he >= 0x7DFu) return false;            // Inf/NaN, or close enough to overflow a TwoSum
    if (((xh << 1) | f80_lo(x)) == 0) {                        // x = +-0: only the sign of a zero sum can change
        if (((hh << 1) | f80_lo(h)) == 0) h = h + x;
        return true;
    }
    double s1, e1, s2, e2, a, b0, b, c;
    f80p_two_sum(h, x, s1, e1);
    f80p_two_sum(e1, l, s2, e2);
    // the last two sums are ordered, so the three-operation FastTwoSum is exact for them:
    //  |s2| <= ulp(s1)/2 + ulp(h)/2 <= |s1| unless h + x cancelled exactly — and then s1 is a non-zero multiple of
    //  ulp(h)/2 >= |l| = |s2|, or zero (FastTwoSum(0, y) is exact);  b0 is zero or at least one ulp of s2 > 2|e2|.
    a = s1 + s2;
    b0 = s2 - (a - s1);
    b = b0 + e2;
    c = e2 - (b - b0);
    const uint32_t ah = f80_hi(a), al = f80_lo(a);
    const uint32_t ae = (ah >> 20) & 0x7FFu;
    if (((ah << 1) | al) == 0) {                               // exact cancellation: +0 in round-to-nearest
        if (b != 0.0 || c != 0.0) return false;
        h = 0.0; l = 0.0;
        return true;
    }
    if (ae < 0x080u) return false;                             // 64-bit grid would leave the double range
    const bool pow2 = ((ah & 0x000FFFFFu) | al) == 0;
    const bool back = pow2 && b != 0.0 && ((f80_hi(b) ^ ah) >> 31) != 0;      // V lies in the binade below a
    const uint32_t qe = ae - 63u - (back ? 1u : 0u);           // biased exponent of the grid q = ulp64(V)
    const double half_q = f80_from_hi((qe - 1u) << 20);
    const double C = f80_from_hi(((qe + 52u) << 20) | 0x00080000u);           // 1.5 * 2^52 * q
    double r = (b + C) - C;                                    // b to a multiple of q, ties to even
    const double d = b - r;                                    // exact
    if (c != 0.0 && (d == half_q || d == -half_q)) {           // b sat on a tie and c breaks it
        if ((d > 0.0) == (c > 0.0)) r = r + (d + d);
        SDEGPU_F80P_COUNT(0);
    }
    if (back) SDEGPU_F80P_COUNT(1);
    if (c == 0.0 && (d == half_q || d == -half_q)) SDEGPU_F80P_COUNT(2);
    const double hn = a + r;                                   // FastTwoSum, |a| >= |r|
    l = r - (hn - a);
    h = hn;
    return true;
}

// S = h + l as an F80 (both conversions and the add are exact: S has 64 significant bits)
SDEGPU_HD F80 f80_from_pair(double h, double l)
{
    F80 acc = f80_zero();
    const uint64_t hb = f80_dbits(h);
    if (((hb >> 52) & 0x7FF) == 0x7FF) { acc.special = 1; acc.sp = h; return acc; }
    f80_from_double(h, acc.mant, acc.exp, acc.sign);
    if (l != 0.0) f80_add(acc, l);
    return acc;
}

#if defined(__CUDACC__)
// R's cumsum accumulator on the device: the floating-point pair while the column stays in its range (always,
// for data that is not within 2^-895 of underflow or 2^992 of overflow), the integer emulation — out of line,
// its state in local memory — from the first element that is not, for the rest of the column.
// add() returns (double)sum, the value cumsum stores.
static __device__ __noinline__ double f80_slow_add(F80 *big, bool first, double h, double l, double x)
{
    if (first) *big = f80_from_pair(h, l);
    f80_add(*big, x);
    return f80_to_double(*big);
}

// The integer state lives in a separate object (`big`, one per accumulator, its address is taken): kept inside this
// struct it dragged h, l and the flag into local memory with it — a load and a store per element (ncu: half of all
// stalls were long-scoreboard waits on those).
struct CumsumAcc {
    double h = 0.0, l = 0.0;
    bool slow = false;
    __device__ __forceinline__ double add(double x, F80 *big)
    {
        if (!slow && f80p_add(h, l, x)) return h;
        const double out = f80_slow_add(big, !slow, h, l, x);
        slow = true;
        return out;
    }
};
#endif

}  // namespace sdegpu
