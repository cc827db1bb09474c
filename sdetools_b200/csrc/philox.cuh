// Counter-based increment generator fused into the step loop.
// Replaces rnorm() in rBM (R/sdetools.R:19): no increment array ever exists in HBM.
//
//   key     = (seed lo, seed hi)
//   counter = (path lo, path hi, block, channel)        one block = two consecutive steps
//   r0..r3  = Philox4x32-10 (Salmon et al. SC'11; KATs in SURVEY.md App. F)
//   u1 = 2 - [1,2)-double built from the top 52 bits of r1:r0   in (0,1]
//   v  =     [1,2)-double built from the top 52 bits of r3:r2   (angle fraction + 1)
//   z0 = sqrt(-2 log u1) * cospi(2 v),  z1 = sqrt(-2 log u1) * sinpi(2 v)
#pragma once
#include <stdint.h>

namespace sdegpu {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;

__device__ __forceinline__ void philox4x32_10(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3,
                                              uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
        const uint32_t hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
}

// [1,2) double from the top 52 bits of hi:lo
__device__ __forceinline__ double bits_to_12(uint32_t lo, uint32_t hi)
{
    return __hiloint2double((int)(0x3FF00000u | (hi >> 12)), (int)((hi << 20) | (lo >> 12)));
}

// Two independent standard normals for (path, block, channel).
__device__ __forceinline__ void normal_pair(uint64_t seed, uint64_t path, uint32_t block, uint32_t channel,
                                            double &z0, double &z1)
{
    uint32_t c0 = (uint32_t)path, c1 = (uint32_t)(path >> 32), c2 = block, c3 = channel;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u1 = 2.0 - bits_to_12(c0, c1);
    const double v = bits_to_12(c2, c3);
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * v, &s, &c);
    z0 = rad * c;
    z1 = rad * s;
}

}  // namespace sdegpu
