// Host check of sdetools_b200/csrc/float80.cuh (the device's x87 emulation)
// against the real x87 FPU: same header, compiled with g++, random + adversarial
// double sequences.  Prints the number of mismatching partial sums.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

static unsigned long long g_cases[3] = {0, 0, 0};      // fast-path corner cases actually exercised
#define SDEGPU_F80P_COUNT(k) (++g_cases[k])
#include "../../sdetools_b200/csrc/float80.cuh"

using namespace sdegpu;

static uint64_t g_fast = 0, g_slow = 0;

static bool same_bits(const F80 &acc, long double ref)
{
    unsigned char raw[16] = {0};
    std::memcpy(raw, &ref, 10);
    uint64_t m; uint16_t se;
    std::memcpy(&m, raw, 8); std::memcpy(&se, raw + 8, 2);
    return (m == acc.mant) && ((se & 0x7FFF) == (uint16_t)(acc.exp + 16383)) && ((uint32_t)(se >> 15) == acc.sign);
}

static uint64_t run(const std::vector<double> &xs, bool verbose)
{
    long double ref = 0.0L;
    F80 acc = f80_zero();
    double ph = 0.0, pl = 0.0;                 // the floating-point pair, with the kernel's fallback rule:
    bool slow = false;                         // once a case is refused the column stays on the integer path
    F80 sacc = f80_zero();
    uint64_t bad = 0;
    for (size_t i = 0; i < xs.size(); ++i) {
        ref += (long double)xs[i];
        f80_add(acc, xs[i]);
        double pe;
        if (!slow && f80p_add(ph, pl, xs[i])) { pe = ph; ++g_fast; }
        else {
            if (!slow) { sacc = f80_from_pair(ph, pl); slow = true; }
            f80_add(sacc, xs[i]);
            pe = f80_to_double(sacc);
            ++g_slow;
        }
        const double r = (double)ref, e = f80_to_double(acc);
        bool ok = (std::memcmp(&r, &e, 8) == 0) || (std::isnan(r) && std::isnan(e));
        ok = ok && ((std::memcmp(&r, &pe, 8) == 0) || (std::isnan(r) && std::isnan(pe)));
        if (ok && !slow && std::isfinite(r) && ref != 0.0L) ok = same_bits(f80_from_pair(ph, pl), ref);
        if (ok && std::isfinite(r) && ref != 0.0L) {       // accumulator itself must match bit for bit
            ok = same_bits(acc, ref);
        }
        if (!ok) {
            if (verbose && bad < 5) std::fprintf(stderr, "mismatch at %zu: x=%a ref=%La int=%a pair=%a (h=%a l=%a slow=%d)\n", i, xs[i], ref, e, pe, ph, pl, (int)slow);
            ++bad;
        }
    }
    return bad;
}

int main(int argc, char **argv)
{
    std::mt19937_64 rng(12345);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::uniform_int_distribution<int> ex(-60, 60), wide(-1070, 1000), coin(0, 9);
    uint64_t bad = 0, total = 0;
    const int reps = argc > 1 ? std::atoi(argv[1]) : 400;
    for (int rep = 0; rep < reps; ++rep) {
        std::vector<double> xs;
        const int n = 5000;
        long double cur = 0.0L;                                          // x87 sum of xs so far (for targeted cases)
        for (int i = 0; i < n; ++i) {
            double v = nd(rng);
            switch (rep % 15) {
            case 0: break;                                              // N(0,1) increments
            case 1: v = std::ldexp(v, ex(rng)); break;                  // mixed magnitudes
            case 2: v = std::ldexp(v, wide(rng)); break;                // incl. subnormals / huge
            case 3: v = (i & 1) ? -xs.back() * (1 + std::ldexp(1.0, -52) * coin(rng)) : v; break;   // near-cancellation
            case 4: v = std::ldexp((double)(rng() >> 11), ex(rng) - 53) * ((rng() & 1) ? 1 : -1); break;  // full mantissas
            case 5: v = (coin(rng) == 0) ? 0.0 : ((coin(rng) == 1) ? -0.0 : std::ldexp(1.0, ex(rng))); break;   // zeros, powers of two
            case 6: v = std::ldexp(1.0, -i % 120) * ((i % 3) ? 1 : -1); break;   // ties / sticky patterns
            case 7: v = v * v * 1e-4; break;                            // QV-like positive sums
            case 8: {                                                    // lands the sum on / next to powers of two and ties
                const long double t = std::ldexp(1.0L, ex(rng) / 4) + std::ldexp((long double)(coin(rng) - 4), ex(rng) / 4 - 64 + coin(rng) % 3);
                v = (double)(t - cur);
                break;
            }
            case 9: v = std::ldexp((double)((rng() >> 11) | 1), -52 - (int)(rng() % 24)) * ((rng() & 1) ? 1 : -1); break;  // long carries
            case 10: v = (i % 2) ? std::ldexp(1.0, -(i % 70)) : -std::ldexp(1.0, -(i % 70) - 1 - (i % 13)); break;        // borrow chains
            case 11: v = std::ldexp(nd(rng), (i % 50 == 0) ? 40 : -20); break;                                            // rare large jumps
            case 12: {                                                   // exact 64-bit ties and near-ties of the running sum
                int e; std::frexp((double)cur, &e);                      // cur = m * 2^e, 0.5 <= m < 1: ulp64 = 2^(e-64)
                if (cur == 0.0L || (i % 3) == 0) { v = std::ldexp((double)((rng() >> 11) | 1), ex(rng) / 3 - 52); break; }
                const double halfulp = std::ldexp(1.0, e - 65);
                v = halfulp * (double)(2 * (int)(rng() % 9) - 7);        // odd multiple of ulp64/2: a tie
                const int k = coin(rng);
                if (k < 3) v += std::ldexp(1.0, e - 65 - 20 - (int)(rng() % 30));   // just above
                else if (k < 6) v -= std::ldexp(1.0, e - 65 - 20 - (int)(rng() % 30));  // just below
                if (coin(rng) < 2) v = -v;
                break;
            }
            case 14: {                                                   // cancel the rounded sum itself: h + x = 0 or a few ulps, l survives
                if (cur == 0.0L || (i % 2) == 0) { v = std::ldexp((double)((rng() >> 11) | 1), ex(rng) / 2 - 52) * ((rng() & 1) ? 1 : -1); break; }
                const double hcur = (double)cur;
                int e; std::frexp(hcur, &e);
                v = -hcur + std::ldexp((double)((int)(rng() % 5) - 2), e - 53 - (int)(rng() % 2));
                break;
            }
            case 13: {                                                   // sums that cross a power of two downwards / upwards
                int e; std::frexp((double)cur, &e);
                if (cur == 0.0L || (i % 4) == 0) { v = std::ldexp(1.0 + std::ldexp((double)(rng() % 4096), -52 - 11 + 11), ex(rng) / 3); break; }
                const long double p2 = std::ldexp(1.0L, e - 1 + (int)(rng() % 2));
                const long double target = (cur < 0 ? -p2 : p2) + std::ldexp((long double)((int)(rng() % 7) - 3), e - 66 + (int)(rng() % 3));
                v = (double)(target - cur);
                break;
            }
            }
            xs.push_back(v);
            cur += (long double)v;
        }
        bad += run(xs, true);
        total += xs.size();
    }
    // specials
    {
        std::vector<double> xs = {1.0, INFINITY, 2.0, -INFINITY, 3.0};
        bad += run(xs, true);
        std::vector<double> ys = {5e-324, 5e-324, -5e-324, 1e308, 1e308, 1e308, -1e308, 4.9e-324};
        bad += run(ys, true);
        total += xs.size() + ys.size();
    }
    std::printf("checked=%llu mismatches=%llu fast=%llu slow=%llu\n", (unsigned long long)total, (unsigned long long)bad,
                (unsigned long long)g_fast, (unsigned long long)g_slow);
    std::printf("fast-path corners: tie broken by the third term=%llu, sum below a power of two=%llu, exact tie=%llu\n", g_cases[0],
                g_cases[1], g_cases[2]);
    if (!g_cases[0] || !g_cases[1] || !g_cases[2]) { std::printf("corner cases not exercised\n"); return 2; }
    return bad ? 1 : 0;
}
