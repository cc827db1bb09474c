// Counter-based increment generator fused into the step loop.
// Replaces rnorm() in rBM (R/sdetools.R:19): no increment array ever exists in HBM.
// Like R's default rnorm it draws by inversion (uniform -> qnorm); the uniform comes from
// Philox4x32-10 instead of Mersenne-Twister so that a draw is a pure function of
// (seed, global path id, step, channel).
//
//   key     = (seed lo, seed hi)
//   r0..r3  = Philox4x32-10(counter, key)   (Salmon et al. SC'11; KATs in SURVEY.md App. F)
//
//   One Brownian channel (every catalogue model with a vector-valued g):
//     counter = (path lo, path hi, block, channel), one block = FOUR consecutive steps, step 4*block+s uses r_s:
//       sign = bit 31, m = low 31 bits, u = m * 2^-31                               (P(octave k) = 2^-(k+1) exactly)
//       if m < 2^19 (probability 2^-12: the draw lies beyond 2^-12 in the tail) the 21 low mantissa bits of u
//       are filled from word s of the extension block, counter (path, block, channel | 0x40000000): tail draws
//       keep full 52-bit resolution down to u = 2^-52 (|z| up to 8.2) while the bulk spends 32 bits per normal.
//   Matrix-valued g (linear model, nB channels):
//     counter = (path lo, path hi, step, 0x80000000 + pair), channel 2*pair uses (r0,r1), 2*pair+1 uses (r2,r3):
//       52-bit u from the top 52 bits of hi:lo, sign = bit 11 of lo.
//   z = sign * a(u), a(u) = -qnorm(u/2) (half-normal quantile): a table of cubics on 52 octaves x 32 mantissa
//   sub-intervals indexed straight from the exponent/mantissa bits of u (tools/gen_icdf_table.py, |error| < 8e-10:
//   a Kolmogorov distance ~3e-10 from the normal law, three orders below the 1/sqrt(N) resolution of a 10^13-draw
//   ensemble), staged in shared memory: 2 DADD + 3 DFMA + 2 LDS.128, no transcendental, no conversion, no branch.
#pragma once
#include <stdint.h>

#include "icdf_table.inc"

namespace sdegpu {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;
constexpr int ICDF_TABLE_DOUBLES = 4 * ICDF_NSEG;      // two planes of 16-byte entries: {c0,c1} and {c2,c3}
constexpr int ICDF_TABLE_BYTES = ICDF_TABLE_DOUBLES * 8;

// The ten round keys of a seed, expanded once on the host: as kernel parameters they sit in the
// constant bank and feed the round's LOP3 directly (no per-step key-schedule instructions).
struct PhiloxKey {
    uint32_t k[20];      // k[2r], k[2r+1] = key of round r
};

inline PhiloxKey make_philox_key(uint64_t seed)
{
    PhiloxKey key;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        key.k[2 * r] = k0; key.k[2 * r + 1] = k1;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    return key;
}

__device__ __forceinline__ void philox4x32_10(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, const PhiloxKey &key)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
        const uint32_t hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ key.k[2 * r], n2 = hi0 ^ c3 ^ key.k[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
}

// Shared-memory layouts of the quantile table (entries are double2; plane B follows plane A).
//   plain : 2 x ICDF_NSEG entries, 52 KB — kernels that draw a few normals per step.
//   hot   : each plane followed by a 16-way replica of its top ICDF_HOT_SEG segments (the six highest
//           octaves, 98.4 % of all draws): lane L reads replica L%16, so the eight lanes a 128-bit load
//           serves per wavefront sit in eight different bank quads and never conflict unless a lane holds
//           a tail draw.  Measured on the first layout (three 8-byte planes, random segments): 5.4
//           wavefronts per LDS.64, and the LSU — shared by the four sub-partitions — was the busiest unit
//           of k_terminal (65 of 118 cycles per warp-step); after replication 42.
constexpr int ICDF_HOT_SEG = 6 * 32;
constexpr int ICDF_HOT_T0 = ICDF_NSEG - ICDF_HOT_SEG;
constexpr int ICDF_HOT_STRIDE = ICDF_NSEG + 16 * ICDF_HOT_SEG;           // entries per plane in the hot layout
constexpr int ICDF_HOT_TABLE_DOUBLES = 4 * ICDF_HOT_STRIDE;
constexpr int ICDF_HOT_TABLE_BYTES = ICDF_HOT_TABLE_DOUBLES * 8;

// cooperative copy of the quantile table into shared memory (call before __syncthreads)
__device__ __forceinline__ void icdf_stage(double *sh_tab, const double *__restrict__ g_tab, int tid, int nthreads)
{
    const double2 *src = reinterpret_cast<const double2 *>(g_tab);
    double2 *dst = reinterpret_cast<double2 *>(sh_tab);
    for (int i = tid; i < 2 * ICDF_NSEG; i += nthreads) dst[i] = __ldg(src + i);
}

__device__ __forceinline__ void icdf_stage_hot(double *sh_tab, const double *__restrict__ g_tab, int tid, int nthreads)
{
    const double2 *src = reinterpret_cast<const double2 *>(g_tab);
    double2 *dst = reinterpret_cast<double2 *>(sh_tab);
    for (int i = tid; i < 2 * ICDF_HOT_STRIDE; i += nthreads) {
        const int plane = i / ICDF_HOT_STRIDE, j = i - plane * ICDF_HOT_STRIDE;
        const int t = j < ICDF_NSEG ? j : ICDF_HOT_T0 + ((j - ICDF_NSEG) >> 4);
        dst[i] = __ldg(src + plane * ICDF_NSEG + t);
    }
}

// a(u) with sign, from the two words of y0 = 1 + u (a [1,2) double) — HOT selects the layout staged above
template <bool HOT>
__device__ __forceinline__ double icdf_eval(uint32_t y0hi, uint32_t y0lo, uint32_t signbit, const double *sh_tab)
{
    constexpr int STRIDE = HOT ? ICDF_HOT_STRIDE : ICDF_NSEG;
    const double u = __hiloint2double((int)y0hi, (int)y0lo) - 1.0;   // [0,1): exponent = octave, mantissa = position
    const uint32_t uh = (uint32_t)__double2hiint(u);
    int t = (int)(uh >> ICDF_SHIFT) - ICDF_BASE;
    t = t < 0 ? 0 : t;                                           // u < 2^-52 (incl. 0) -> last segment
    if (HOT) {
        const int lane16 = (int)(threadIdx.x & 15u);
        t = t >= ICDF_HOT_T0 ? ICDF_NSEG + ((t - ICDF_HOT_T0) << 4) + lane16 : t;
    }
    // r = y - c with y = u*2^(k+1) in [1,2) and c the segment centre; both share the exponent and the top
    // mantissa bits, so r is the low mantissa of u re-based at 1 minus half a segment: one LOP3 + one DADD
    // against an immediate (1 + 2^-6), exact.
    constexpr uint32_t LOWMASK = (1u << ICDF_SHIFT) - 1u;
    const double d = __hiloint2double((int)((uh & LOWMASK) | 0x3FF00000u), __double2loint(u));
    const double r = d - (1.0 + 1.0 / (double)(1u << (21 - ICDF_SHIFT)));
    const double2 *p = reinterpret_cast<const double2 *>(sh_tab) + t;
    const double2 c01 = p[0], c23 = p[STRIDE];
    double a = fma(c23.y, r, c23.x);
    a = fma(a, r, c01.y);
    a = fma(a, r, c01.x);
    return __hiloint2double(__double2hiint(a) ^ (int)signbit, __double2loint(a));
}

// one standard normal from 64 random bits (matrix-valued g: two per Philox block)
template <bool HOT = false>
__device__ __forceinline__ double icdf_normal(uint32_t lo, uint32_t hi, const double *sh_tab)
{
    return icdf_eval<HOT>(0x3FF00000u | (hi >> 12), (hi << 20) | (lo >> 12), (lo << 20) & 0x80000000u, sh_tab);
}

// one standard normal from a 32-bit word plus the (usually zero) 21 extension bits
template <bool HOT = false>
__device__ __forceinline__ double icdf_normal32(uint32_t w, uint32_t ext21, const double *sh_tab)
{
    return icdf_eval<HOT>(0x3FF00000u | ((w >> 11) & 0x000FFFFFu), (w << 21) | ext21, w & 0x80000000u, sh_tab);
}

// The 128 random bits of (path, block, channel).
__device__ __forceinline__ void philox_block(const PhiloxKey &key, uint64_t path, uint32_t block, uint32_t channel,
                                             uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3)
{
    r0 = (uint32_t)path; r1 = (uint32_t)(path >> 32); r2 = block; r3 = channel;
    philox4x32_10(r0, r1, r2, r3, key);
}

// Extension block of the tail draws, out of line on purpose: it runs for ~1 block in 1000 and, inlined, its
// twenty round keys would be copied into vector registers at the top of every iteration of the hot loop.
static __device__ __noinline__ uint4 philox_extension(uint32_t seed_lo, uint32_t seed_hi, uint32_t path_lo, uint32_t path_hi,
                                               uint32_t block, uint32_t channel)
{
    uint32_t c0 = path_lo, c1 = path_hi, c2 = block, c3 = channel | 0x40000000u, k0 = seed_lo, k1 = seed_hi;
#pragma unroll 1
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
        const uint32_t hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

// Four standard normals (steps 4*block .. 4*block+3 of one channel) from the words of a Philox block.
// The tail extension is evaluated once per block and taken by ~1 block in 1000.
template <bool HOT = false>
__device__ __forceinline__ void normals_from_block(const PhiloxKey &key, uint64_t path, uint32_t block, uint32_t channel,
                                                   const uint32_t (&r)[4], const double *sh_tab, double (&z)[4])
{
    uint32_t ext[4] = {0u, 0u, 0u, 0u};
    const uint32_t m0 = r[0] & 0x7FFFFFFFu, m1 = r[1] & 0x7FFFFFFFu, m2 = r[2] & 0x7FFFFFFFu, m3 = r[3] & 0x7FFFFFFFu;
    if (min(min(m0, m1), min(m2, m3)) < (1u << 19)) {
        const uint4 e = philox_extension(key.k[0], key.k[1], (uint32_t)path, (uint32_t)(path >> 32), block, channel);
        ext[0] = m0 < (1u << 19) ? e.x >> 11 : 0u;
        ext[1] = m1 < (1u << 19) ? e.y >> 11 : 0u;
        ext[2] = m2 < (1u << 19) ? e.z >> 11 : 0u;
        ext[3] = m3 < (1u << 19) ? e.w >> 11 : 0u;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) z[k] = icdf_normal32<HOT>(r[k], ext[k], sh_tab);
}

template <bool HOT = false>
__device__ __forceinline__ void normal_quad(const PhiloxKey &key, uint64_t path, uint32_t block, uint32_t channel,
                                            const double *sh_tab, double (&z)[4])
{
    uint32_t r[4];
    philox_block(key, path, block, channel, r[0], r[1], r[2], r[3]);
    normals_from_block<HOT>(key, path, block, channel, r, sh_tab, z);
}

// Two standard normals for a channel pair of the matrix-valued-g generator (52-bit uniforms).
template <bool HOT = false>
__device__ __forceinline__ void normal_pair(const PhiloxKey &key, uint64_t path, uint32_t block, uint32_t channel,
                                            const double *sh_tab, double &z0, double &z1)
{
    uint32_t c0 = (uint32_t)path, c1 = (uint32_t)(path >> 32), c2 = block, c3 = channel;
    philox4x32_10(c0, c1, c2, c3, key);
    z0 = icdf_normal<HOT>(c0, c1, sh_tab);
    z1 = icdf_normal<HOT>(c2, c3, sh_tab);
}

}  // namespace sdegpu
