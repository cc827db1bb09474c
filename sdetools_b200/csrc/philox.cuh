// Counter-based increment generator fused into the step loop.
// Replaces rnorm() in rBM (R/sdetools.R:19): no increment array ever exists in HBM.
// Like R's default rnorm it draws by inversion (uniform -> qnorm); the uniform comes from
// Philox4x32-10 instead of Mersenne-Twister so that a draw is a pure function of
// (seed, global path id, step, channel).
//
//   key     = (seed lo, seed hi)
//   r0..r3  = Philox4x32-10(counter, key)   (Salmon et al. SC'11; KATs in SURVEY.md App. F)
//
//   One Brownian channel (every catalogue model with a vector-valued g) — "stream v2":
//     counter = (path lo, path hi, block, channel), one block = FOUR consecutive steps, step 4*block+s uses word w = r_s:
//       s = (w | 1) - 2^31   an odd integer in (-2^31, 2^31): symmetric, never zero;  sign(z) = sign(s), v = |s|
//       u = v * 2^-31        (31-bit resolution, P(v < 2^k) = 2^(k-31) exactly)
//       if v < 2^19 (probability 2^-12: the draw lies beyond 2^-12 in the tail) the cell is refined with the top 22
//       bits e of word s of the extension block, counter (path, block, channel | 0x40000000):
//       u = ((v-1) 2^21 + e + 1) 2^-52, so tail draws keep 52-bit resolution down to u = 2^-52 (|z| up to 8.2)
//       while the bulk spends 32 bits per normal.
//   Matrix-valued g (linear model, nB channels):
//     counter = (path lo, path hi, step, 0x80000000 + pair), channel 2*pair uses (r0,r1), 2*pair+1 uses (r2,r3):
//       52-bit u from the top 52 bits of hi:lo, sign = bit 11 of lo.
//   z = sign * a(u), a(u) = -qnorm(u/2) (half-normal quantile), evaluated from piecewise cubics (|error| < 8e-10:
//   a Kolmogorov distance ~3e-10 from the normal law, three orders below the 1/sqrt(N) resolution of a 10^13-draw
//   ensemble) in one of two ways that approximate the same function:
//     generic  52 octaves x 32 mantissa sub-intervals of u, indexed from the exponent/mantissa bits of u, local
//              coordinate r (icdf_table.inc, 52 KB in shared memory): any u in [2^-52, 1).
//     fast     the register-resident kernels: s is formed from w in ONE exact FP64 subtraction ({w|1, 0x43300000} as a
//              double is 2^52 + (w|1)), its high word indexes 12 octaves x 32 sub-intervals of v directly and the cubic
//              runs in the raw variable |s| (icdf32_table.inc, 12 KB, replicated REPL-fold in shared memory so that the
//              lanes of a quarter-warp hit different bank quads): LOP3 + DADD + LOP3 + LEA + 2 LDS.128 + 3 DFMA + LOP3
//              per normal, no branch; the 2^-12 tail leaves through an out-of-line call that uses the generic tables.
#pragma once
#include <stdint.h>

#include "icdf_table.inc"
#include "icdf32_table.inc"

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

// one standard normal of stream v2 from a 32-bit word and the word of the extension block (used only when the
// draw is in the tail), generic evaluation: u's 52 mantissa bits are v << 21, or (v-1) << 21 | e22, + 1 in the tail
template <bool HOT = false>
__device__ __forceinline__ double icdf_normal32(uint32_t w, uint32_t extw, const double *sh_tab)
{
    const int32_t si = (int32_t)((w | 1u) ^ 0x80000000u);         // (w|1) - 2^31
    const uint32_t v = (uint32_t)(si < 0 ? -si : si);
    uint32_t hi = v >> 11, lo = v << 21;
    if (v < (1u << 19)) {
        const uint64_t m = (((uint64_t)(v - 1u)) << 21) + (uint64_t)(extw >> 10) + 1ull;
        hi = (uint32_t)(m >> 32); lo = (uint32_t)m;
    }
    return icdf_eval<HOT>(0x3FF00000u | hi, lo, si < 0 ? 0x80000000u : 0u, sh_tab);
}

// The 128 random bits of (path, block, channel).
__device__ __forceinline__ void philox_block(const PhiloxKey &key, uint64_t path, uint32_t block, uint32_t channel,
                                             uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3)
{
    r0 = (uint32_t)path; r1 = (uint32_t)(path >> 32); r2 = block; r3 = channel;
    philox4x32_10(r0, r1, r2, r3, key);
}

// The same 128 bits for many blocks of ONE path and channel, with what depends on the path alone computed once.  Round 1
// multiplies M0 by the path's low word (fixed) and M1 by the block index (the same in every lane: a loop counter the compiler
// keeps in the uniform datapath); behind it c2 and c3 depend on the path only, so round 2's M1 * c2 is fixed as well: 17
// instead of 20 wide multiplies per block (they are the generator's expensive instruction, 5 cycles each), the same bits.
struct PhiloxPath {
    uint32_t a;              // path hi ^ k[0]:         c0 after round 1 is hi(M1 * block) ^ a
    uint32_t c3;             // lo(M0 * path lo):       c3 after round 1
    uint32_t h1, l1;         // hi, lo of M1 * c2 with c2 = hi(M0 * path lo) ^ channel ^ k[1], the c2 after round 1

    __device__ __forceinline__ void init(const PhiloxKey &key, uint64_t path, uint32_t channel)
    {
        const uint32_t plo = (uint32_t)path, phi = (uint32_t)(path >> 32);
        a = phi ^ key.k[0];
        c3 = PHILOX_M0 * plo;
        const uint32_t c2 = __umulhi(PHILOX_M0, plo) ^ channel ^ key.k[1];
        h1 = __umulhi(PHILOX_M1, c2);
        l1 = PHILOX_M1 * c2;
    }

    __device__ __forceinline__ void block(const PhiloxKey &key, uint32_t blk, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) const
    {
        // round 1: only M1 * block is left of it
        uint32_t c0 = __umulhi(PHILOX_M1, blk) ^ a, c1 = PHILOX_M1 * blk;
        // round 2: M1 * c2 is (h1, l1)
        {
            const uint32_t hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
            const uint32_t n0 = h1 ^ c1 ^ key.k[2], n2 = hi0 ^ c3 ^ key.k[3];
            c0 = n0; c1 = l1; r2 = n2; r3 = lo0;
        }
        uint32_t c2 = r2, c3r = r3;
#pragma unroll
        for (int r = 2; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
            const uint32_t hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
            const uint32_t n0 = hi1 ^ c1 ^ key.k[2 * r], n2 = hi0 ^ c3r ^ key.k[2 * r + 1];
            c0 = n0; c1 = lo1; c2 = n2; c3r = lo0;
        }
        r0 = c0; r1 = c1; r2 = c2; r3 = c3r;
    }
};

// Extension block of the tail draws, out of line on purpose: it runs for ~1 block in 1000 and, inlined, its
// twenty round keys would be copied into vector registers at the top of every iteration of the hot loop.
__device__ __forceinline__ uint4 philox_extension_inl(uint32_t seed_lo, uint32_t seed_hi, uint32_t path_lo, uint32_t path_hi,
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
static __device__ __noinline__ uint4 philox_extension(uint32_t seed_lo, uint32_t seed_hi, uint32_t path_lo, uint32_t path_hi,
                                               uint32_t block, uint32_t channel)
{
    return philox_extension_inl(seed_lo, seed_hi, path_lo, path_hi, block, channel);
}

// is any of the four words a tail draw?  v < 2^19  <=>  (w|1) - (2^31 - 2^19) < 2^20 as unsigned
__device__ __forceinline__ bool philox_block_has_tail(const uint32_t (&r)[4])
{
    const uint32_t off = 0x80000000u - (1u << 19);
    const uint32_t d0 = (r[0] | 1u) - off, d1 = (r[1] | 1u) - off, d2 = (r[2] | 1u) - off, d3 = (r[3] | 1u) - off;
    return min(min(d0, d1), min(d2, d3)) < (1u << 20);
}

// Four standard normals (steps 4*block .. 4*block+3 of one channel) from the words of a Philox block, generic
// evaluation.  The tail extension is generated once per block and taken by ~1 block in 1000.
template <bool HOT = false>
__device__ __forceinline__ void normals_from_block(const PhiloxKey &key, uint64_t path, uint32_t block, uint32_t channel,
                                                   const uint32_t (&r)[4], const double *sh_tab, double (&z)[4])
{
    uint4 e = make_uint4(0u, 0u, 0u, 0u);
    if (philox_block_has_tail(r)) e = philox_extension(key.k[0], key.k[1], (uint32_t)path, (uint32_t)(path >> 32), block, channel);
    z[0] = icdf_normal32<HOT>(r[0], e.x, sh_tab);
    z[1] = icdf_normal32<HOT>(r[1], e.y, sh_tab);
    z[2] = icdf_normal32<HOT>(r[2], e.z, sh_tab);
    z[3] = icdf_normal32<HOT>(r[3], e.w, sh_tab);
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

// ---- fast evaluation of stream v2 (register-resident kernels) ------------------------------------------------
// Shared-memory layout: plane A {c0,c1} then plane B {c2,c3}, each ICDF32_NSEG segments x REPL replicas of 16 bytes;
// lane L reads replica L % REPL, so with REPL = 8 the eight lanes a 128-bit load serves per wavefront sit in eight
// different bank quads whatever segments they want.  `scale` multiplies every coefficient while staging: a kernel
// whose grid is uniform folds sqrt(dt) (times the model's noise constant) into the table and gets dB straight
// out of the Horner chain.
constexpr int ICDF32_DOUBLES = 4 * ICDF32_NSEG;
template <int REPL> struct Icdf32Layout {
    static_assert(REPL == 1 || REPL == 2 || REPL == 4 || REPL == 8 || REPL == 16, "REPL must be a power of two");
    static constexpr int LOG = REPL == 1 ? 0 : REPL == 2 ? 1 : REPL == 4 ? 2 : REPL == 8 ? 3 : 4;
    static constexpr int PLANE_BYTES = ICDF32_NSEG * REPL * 16;
    static constexpr int BYTES = 2 * PLANE_BYTES;
    static constexpr int DOUBLES = BYTES / 8;
};

template <int REPL>
__device__ __forceinline__ void icdf32_stage(double *sh_tab, const double *__restrict__ g32, double scale, int tid, int nthreads)
{
    const double2 *src = reinterpret_cast<const double2 *>(g32);
    double2 *dst = reinterpret_cast<double2 *>(sh_tab);
    constexpr int PLANE = ICDF32_NSEG * REPL;
    for (int i = tid; i < 2 * PLANE; i += nthreads) {
        const int plane = i / PLANE, seg = (i - plane * PLANE) / REPL;
        const double2 c = __ldg(src + plane * ICDF32_NSEG + seg);
        dst[i] = make_double2(c.x * scale, c.y * scale);
    }
}

__device__ __forceinline__ uint32_t smem_addr_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ double2 lds_f64x2(uint32_t addr)
{
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// The rare path of the fast evaluation, out of line on purpose (defined behind NormalGen).  A lane holds a tail draw in one
// block of ~1000, but the branch is taken by the warp: with 32 lanes, one block in 32 pays for it.  Per word: a tail word
// (|s| < 2^19) goes through the extension block and the generic tables in global memory, every other word through the same
// bulk table as on the common path — so a normal is a function of its own word (and its extension word) alone, whichever
// kernel draws it.  zout is a scratch array of the caller.
template <int REPL>
static __device__ __noinline__ void icdf32_rare_quad(uint32_t seed_lo, uint32_t seed_hi, uint32_t path_lo, uint32_t path_hi,
                                                     uint32_t block, uint32_t channel, uint32_t w0, uint32_t w1, uint32_t w2,
                                                     uint32_t w3, const double *g52, double scale, uint32_t base, double *zout);

// One tail word alone (extension block, generic tables), result in registers: for the kernels that run one warp per
// scheduler and feel the length of the rare path (k_traj; measured against icdf32_rare_quad there: 0.760 vs 0.749 on C3).
static __device__ __noinline__ double icdf32_tail_one(uint32_t seed_lo, uint32_t seed_hi, uint32_t path_lo, uint32_t path_hi,
                                                      uint32_t block, uint32_t channel, uint32_t k, uint32_t w, const double *g52,
                                                      double scale)
{
    const uint4 e = philox_extension_inl(seed_lo, seed_hi, path_lo, path_hi, block, channel);
    const uint32_t ew = k == 0u ? e.x : (k == 1u ? e.y : (k == 2u ? e.z : e.w));
    return scale * icdf_normal32<false>(w, ew, g52);
}

template <int REPL> struct NormalGen {
    using L = Icdf32Layout<REPL>;
    uint32_t base;            // shared-space byte address of plane A with the lane's replica and the exponent bias folded in
    const double *g52;        // generic 52-octave table in global memory (tail draws only)
    double scale;

    __device__ __forceinline__ void init(const double *sh_tab, const double *g52_, double scale_)
    {
        const uint32_t b = smem_addr_u32(sh_tab) + ((threadIdx.x & (REPL - 1)) << 4) - (((1023u + ICDF32_OCT0) << 5) << (4 + L::LOG));
        asm volatile("mov.u32 %0, %1;" : "=r"(base) : "r"(b));   // opaque: kept in a register, not recomputed in the hot loop
        g52 = g52_;
        scale = scale_;
    }

    // (exponent, 5 mantissa bits) of |s| for the exact integer-valued double s = (w|1) - 2^31
    static __device__ __forceinline__ double word_to_s(uint32_t w) { return __hiloint2double(0x43300000, (int)(w | 1u)) - 4503601774854144.0; }
    static __device__ __forceinline__ uint32_t seg_bits(double s) { return (uint32_t)__double2hiint(s) & 0x7FFF8000u; }

    // does the block hold a tail draw (some |s| < 2^19, 1 block in ~1000)?
    static __device__ __forceinline__ bool has_tail(const uint32_t (&r)[4])
    {
        const uint32_t x0 = seg_bits(word_to_s(r[0])), x1 = seg_bits(word_to_s(r[1])), x2 = seg_bits(word_to_s(r[2])), x3 = seg_bits(word_to_s(r[3]));
        return min(min(x0, x1), min(x2, x3)) < 0x41200000u;
    }

    // scale * z of one bulk word (|s| >= 2^19): two 128-bit shared loads, three FMAs, the sign by LOP3
    __device__ __forceinline__ double bulk(uint32_t w) const
    {
        const double s = word_to_s(w);
        const uint32_t addr = base + (seg_bits(s) >> (11 - L::LOG));
        const double2 c01 = lds_f64x2(addr), c23 = lds_f64x2(addr + L::PLANE_BYTES);
        const double v = fabs(s);
        double a = fma(c23.y, v, c23.x);
        a = fma(a, v, c01.y);
        a = fma(a, v, c01.x);
        // sign(s) onto scale*a(|s|) with one LOP3 (an XOR, so a negative scale keeps its meaning)
        return __hiloint2double(__double2hiint(a) ^ (__double2hiint(s) & (int)0x80000000u), __double2loint(a));
    }

    // scale * z for steps 4*block .. 4*block+3 of (path, channel) from the words of its Philox block
    // bulk() with s and its segment bits already at hand
    __device__ __forceinline__ double bulk_at(double s, uint32_t seg) const
    {
        // (the same address by one IMAD.HI on the FMA pipe instead of this LEA.HI on the ALU pipe was measured: 5 % slower,
        // profiles/r2_tune_terminal_imad.txt — its 64-bit addend costs two moves per normal)
        const uint32_t addr = base + (seg >> (11 - L::LOG));
        const double2 c01 = lds_f64x2(addr), c23 = lds_f64x2(addr + L::PLANE_BYTES);
        const double v = fabs(s);
        double a = fma(c23.y, v, c23.x);
        a = fma(a, v, c01.y);
#ifndef SDEGPU_GEN_PLAIN_SIGN
        // sign(s) (c0 + b |s|) = sign(s) c0 + b s: the sign goes onto the loaded constant term (in place, off the critical path)
        // and the last Horner step multiplies by s itself, so the result needs no sign fix-up and no copy of its low word
        const double c0s = __hiloint2double(__double2hiint(c01.x) ^ (__double2hiint(s) & (int)0x80000000u), __double2loint(c01.x));
        return fma(a, s, c0s);
#else
        a = fma(a, v, c01.x);
        return __hiloint2double(__double2hiint(a) ^ (__double2hiint(s) & (int)0x80000000u), __double2loint(a));
#endif
    }

    // a block that holds a tail draw, out of line (see icdf32_rare_quad)
    __device__ __forceinline__ void rare(const PhiloxKey &key, uint64_t path, uint32_t block, uint32_t channel,
                                         const uint32_t (&r)[4], double (&z)[4]) const
    {
        double zt[4];                                            // its address is taken: kept apart from z, which stays in registers
        icdf32_rare_quad<REPL>(key.k[0], key.k[1], (uint32_t)path, (uint32_t)(path >> 32), block, channel, r[0], r[1], r[2], r[3],
                               g52, scale, base, zt);
#pragma unroll
        for (int k = 0; k < 4; ++k) z[k] = zt[k];
    }

    // the tail words of a block whose other words are evaluated already (bulk_quad_clamped): same values as rare()
    __device__ __forceinline__ void fix_tail(const PhiloxKey &key, uint64_t path, uint32_t block, uint32_t channel,
                                             const uint32_t (&r)[4], double (&z)[4]) const
    {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (seg_bits(word_to_s(r[k])) < 0x41200000u)
                z[k] = icdf32_tail_one(key.k[0], key.k[1], (uint32_t)path, (uint32_t)(path >> 32), block, channel, (uint32_t)k, r[k],
                                       g52, scale);
    }

    // The four normals of a block on the bulk path WITHOUT a branch: a tail word (|s| < 2^19) is evaluated on the lowest
    // bulk segment (a finite placeholder) and the return value tells the caller to replace them through fix_tail().  For
    // kernels that schedule the evaluation into the shadow of another dependent chain.
    __device__ __forceinline__ bool bulk_quad_clamped(const uint32_t (&r)[4], double (&z)[4]) const
    {
        uint32_t lowest = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double s = word_to_s(r[k]);
            uint32_t sg = seg_bits(s);
            asm volatile("" : "+r"(sg));
            lowest = min(lowest, sg);
            z[k] = bulk_at(s, max(sg, 0x41200000u));
        }
        return lowest < 0x41200000u;
    }

    __device__ __forceinline__ void quad(const PhiloxKey &key, uint64_t path, uint32_t block, uint32_t channel,
                                         const uint32_t (&r)[4], double (&z)[4]) const
    {
#ifndef SDEGPU_GEN_PLAIN_QUAD
        // the masked high words feed both the tail test and the table addresses: pinned in registers so that the compiler does
        // not derive the address from the high word a second time (one LOP3 per normal less)
        double s[4];
        uint32_t sg[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            s[k] = word_to_s(r[k]);
            sg[k] = seg_bits(s[k]);
            asm volatile("" : "+r"(sg[k]));
        }
        if (min(min(sg[0], sg[1]), min(sg[2], sg[3])) < 0x41200000u) rare(key, path, block, channel, r, z);
        else {
#pragma unroll
            for (int k = 0; k < 4; ++k) z[k] = bulk_at(s[k], sg[k]);
        }
        return;
#endif
        if (has_tail(r)) rare(key, path, block, channel, r, z);
        else {
#pragma unroll
            for (int k = 0; k < 4; ++k) z[k] = bulk(r[k]);
        }
    }

    // the same for NB consecutive blocks with ONE branch: the bulk case is a single basic block of 4*NB independent
    // evaluations that the scheduler can interleave (kernels that run few warps per SM need the ILP)
    template <int NB>
    __device__ __forceinline__ void multi(const PhiloxKey &key, uint64_t path, uint32_t block0, uint32_t channel,
                                          const uint32_t (&r)[NB][4], double (&z)[4 * NB]) const
    {
#ifndef SDEGPU_GEN_PLAIN_QUAD
        // as in quad(): s and its masked high word once per normal, used by the tail test and by the evaluation
        double s[NB][4];
        uint32_t sg[NB][4], lowest = 0xFFFFFFFFu;
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                s[b][k] = word_to_s(r[b][k]);
                sg[b][k] = seg_bits(s[b][k]);
                asm volatile("" : "+r"(sg[b][k]));
                lowest = min(lowest, sg[b][k]);
            }
        if (lowest >= 0x41200000u) {
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int k = 0; k < 4; ++k) z[4 * b + k] = bulk_at(s[b][k], sg[b][k]);
            return;
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            double zq[4];
            quad(key, path, block0 + (uint32_t)b, channel, r[b], zq);
#pragma unroll
            for (int k = 0; k < 4; ++k) z[4 * b + k] = zq[k];
        }
        return;
#endif
        bool tail = false;
#pragma unroll
        for (int b = 0; b < NB; ++b) tail |= has_tail(r[b]);
        if (tail) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                double zq[4];
                quad(key, path, block0 + (uint32_t)b, channel, r[b], zq);
#pragma unroll
                for (int k = 0; k < 4; ++k) z[4 * b + k] = zq[k];
            }
        } else {
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int k = 0; k < 4; ++k) z[4 * b + k] = bulk(r[b][k]);
        }
    }
};

template <int REPL>
static __device__ __noinline__ void icdf32_rare_quad(uint32_t seed_lo, uint32_t seed_hi, uint32_t path_lo, uint32_t path_hi,
                                                     uint32_t block, uint32_t channel, uint32_t w0, uint32_t w1, uint32_t w2,
                                                     uint32_t w3, const double *g52, double scale, uint32_t base, double *zout)
{
    NormalGen<REPL> gen;
    gen.base = base; gen.g52 = g52; gen.scale = scale;
    const uint4 e = philox_extension_inl(seed_lo, seed_hi, path_lo, path_hi, block, channel);
    const uint32_t w[4] = {w0, w1, w2, w3}, ew[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double s = NormalGen<REPL>::word_to_s(w[k]);
        const uint32_t sg = NormalGen<REPL>::seg_bits(s);
        zout[k] = sg < 0x41200000u ? scale * icdf_normal32<false>(w[k], ew[k], g52) : gen.bulk_at(s, sg);
    }
}

}  // namespace sdegpu
