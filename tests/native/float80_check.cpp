// Host check of sdetools_b200/csrc/float80.cuh (the device's x87 emulation)
// against the real x87 FPU: same header, compiled with g++, random + adversarial
// double sequences.  Prints the number of mismatching partial sums.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../../sdetools_b200/csrc/float80.cuh"

using namespace sdegpu;

static uint64_t run(const std::vector<double> &xs, bool verbose)
{
    long double ref = 0.0L;
    F80 acc = f80_zero();
    uint64_t bad = 0;
    for (size_t i = 0; i < xs.size(); ++i) {
        ref += (long double)xs[i];
        f80_add(acc, xs[i]);
        const double r = (double)ref, e = f80_to_double(acc);
        bool ok = (std::memcmp(&r, &e, 8) == 0) || (std::isnan(r) && std::isnan(e));
        if (ok && std::isfinite(r) && ref != 0.0L) {       // accumulator itself must match bit for bit
            unsigned char raw[16] = {0};
            std::memcpy(raw, &ref, 10);
            uint64_t m; uint16_t se;
            std::memcpy(&m, raw, 8); std::memcpy(&se, raw + 8, 2);
            ok = (m == acc.mant) && ((se & 0x7FFF) == (uint16_t)(acc.exp + 16383)) && ((uint32_t)(se >> 15) == acc.sign);
        }
        if (!ok) {
            if (verbose && bad < 5) std::fprintf(stderr, "mismatch at %zu: x=%a ref=%La got=%a\n", i, xs[i], ref, e);
            ++bad;
        }
    }
    return bad;
}

int main()
{
    std::mt19937_64 rng(12345);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::uniform_int_distribution<int> ex(-60, 60), wide(-1070, 1000), coin(0, 9);
    uint64_t bad = 0, total = 0;
    for (int rep = 0; rep < 200; ++rep) {
        std::vector<double> xs;
        const int n = 5000;
        for (int i = 0; i < n; ++i) {
            double v = nd(rng);
            switch (rep % 8) {
            case 0: break;                                              // N(0,1) increments
            case 1: v = std::ldexp(v, ex(rng)); break;                  // mixed magnitudes
            case 2: v = std::ldexp(v, wide(rng)); break;                // incl. subnormals / huge
            case 3: v = (i & 1) ? -xs.back() * (1 + std::ldexp(1.0, -52) * coin(rng)) : v; break;   // near-cancellation
            case 4: v = std::ldexp((double)(rng() >> 11), ex(rng) - 53) * ((rng() & 1) ? 1 : -1); break;  // full mantissas
            case 5: v = (coin(rng) == 0) ? 0.0 : ((coin(rng) == 1) ? -0.0 : std::ldexp(1.0, ex(rng))); break;   // zeros, powers of two
            case 6: v = std::ldexp(1.0, -i % 120) * ((i % 3) ? 1 : -1); break;   // ties / sticky patterns
            case 7: v = v * v * 1e-4; break;                            // QV-like positive sums
            }
            xs.push_back(v);
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
    std::printf("checked=%llu mismatches=%llu\n", (unsigned long long)total, (unsigned long long)bad);
    return bad ? 1 : 0;
}
