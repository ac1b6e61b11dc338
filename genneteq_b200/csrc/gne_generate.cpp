// genneteq_b200 — deterministic synthetic-instance generator for the large benchmark configurations
// (SURVEY.md section 8d).  Self-contained (no other translation unit of the library is needed): it is
// compiled into the product library (gne_generate_instance, include/genneteq_b200.h) and, by the test
// infrastructure's own Makefile, into a stand-alone libgne_gen.so that bench.py's reference arm loads, so
// that the reference process never maps the product library.
#include <algorithm>
#include <cmath>
#include <stdint.h>
#include <vector>

// The Appendix-A family (ring + random distinct arcs, same value ranges, 6-decimal values) from a
// counter-based generator, so that 10^8 arcs take seconds and every rank builds the same arrays.
static inline uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
static inline double unit(uint64_t h) { return (double) (h >> 11) * (1.0 / 9007199254740992.0); }
static inline double r6(double x) { return floor(x * 1.0e6 + 0.5) / 1.0e6; }

extern "C" int gne_generate_instance_impl(int n, int64_t m, uint64_t seed, int mode,
                                     int32_t *tail, int32_t *head, double *ub, double *cost, double *mult, double *supply)
{
    if (n < 3 || m < n || m > (int64_t) n * (n - 1)) return 2;      // GNE_ERR_ARG: need n >= 3 and n <= m <= n(n-1)
    const int64_t base = m / n, rem = m % n;
    int64_t a = 0;
    std::vector<int32_t> hs;
    for (int i = 0; i < n; ++i) {
        const int64_t d = base + (i < rem ? 1 : 0);
        hs.clear();
        hs.push_back((i + 1) % n);                      // ring arc keeps the graph connected
        uint64_t ctr = 0;
        while ((int64_t) hs.size() < d) {
            const int64_t want = d - (int64_t) hs.size();
            for (int64_t k = 0; k < want; ++k) {
                const uint64_t h = mix64(mix64(seed ^ ((uint64_t) i << 20)) + (ctr++));
                const int32_t j = (int32_t) (h % (uint64_t) n);
                if (j != i) hs.push_back(j);
            }
            std::sort(hs.begin(), hs.end());
            hs.erase(std::unique(hs.begin(), hs.end()), hs.end());
        }
        std::sort(hs.begin(), hs.end());
        for (size_t k = 0; k < hs.size(); ++k, ++a) {
            const uint64_t key = mix64(seed * 0x100000001B3ULL + (uint64_t) a);
            tail[a] = i; head[a] = hs[k];
            ub[a] = r6(10.0 + 90.0 * unit(mix64(key + 1)));
            cost[a] = r6(1.0 + 99.0 * unit(mix64(key + 2)));
            const double u = unit(mix64(key + 3));
            double g;
            if (mode == 1) g = 0.70 + 0.29 * u;
            else if (mode == 2) g = 0.9 + 0.2 * u;
            else if (mode == 3) { static const double pw[5] = { 0.5, 0.75, 1.25, 1.5, 2.0 }; g = pw[(int) (u * 5.0) % 5]; }
            else g = 0.5 + u;
            mult[a] = r6(g);
        }
    }
    for (int i = 0; i < n; ++i) {
        const uint64_t h = mix64(seed + 0xABCDEFULL * (uint64_t) (i + 1));
        const int kind = (int) (h % 20);
        const double amount = (double) (1 + (int) ((h >> 8) % 20));
        if (kind == 0) supply[i] = amount * (mode == 1 ? 3.0 : 1.0);
        else if (kind == 1) supply[i] = -amount;
        else supply[i] = 0.0;
    }
    return 0;
}
