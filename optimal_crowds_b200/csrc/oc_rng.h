// oc_rng.h -- numpy's LEGACY random stream (np.random.seed / RandomState: MT19937), restated for the host side of the
// library so that the per-step randomness of the GCFM sweep and the crowd placement can be drawn in C with exactly the
// reference's stream (simulations.py:130-131,140,271,303):
//   next_double      genrand_res53
//   interval(max)    random_interval: masked rejection on 32-bit draws (max < 2^32)
//   permutation(n)   RandomState.permutation(n) == choice(arange(n), n, replace=False): Fisher-Yates from the top
//   gauss()          legacy_gauss: polar Box-Muller with one cached value (has_gauss / cached_gaussian of get_state())
// Pinned against numpy itself by tests/test_cpu_host.py (bit-identical draws and generator state).
#pragma once
#include <cmath>
#include <cstdint>

namespace ocrng {

struct Mt {
    uint32_t *key;   // 624 words (np.random.get_state()[1]), advanced in place
    int pos;         // get_state()[2]
    int has_gauss;   // get_state()[3]
    double gauss_;   // get_state()[4]

    void gen() {  // numpy's mt19937_gen == the reference genrand
        const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAT = 0x9908b0dfu;
        int i;
        uint32_t y;
        for (i = 0; i < 624 - 397; i++) {
            y = (key[i] & UPPER) | (key[i + 1] & LOWER);
            key[i] = key[i + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MAT);
        }
        for (; i < 623; i++) {
            y = (key[i] & UPPER) | (key[i + 1] & LOWER);
            key[i] = key[i + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MAT);
        }
        y = (key[623] & UPPER) | (key[0] & LOWER);
        key[623] = key[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MAT);
        pos = 0;
    }
    uint32_t next32() {
        if (pos == 624) gen();
        uint32_t y = key[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    double next_double() {
        const int32_t a = next32() >> 5, b = next32() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
    uint32_t interval(uint32_t max) {
        if (max == 0) return 0;
        uint32_t mask = max;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        while ((v = (next32() & mask)) > max) {}
        return v;
    }
    void permutation(int n, int *out) {
        for (int i = 0; i < n; i++) out[i] = i;
        for (int i = n - 1; i > 0; i--) {
            const int j = (int)interval((uint32_t)i);
            const int t = out[i]; out[i] = out[j]; out[j] = t;
        }
    }
    double gauss() {
        if (has_gauss) {
            const double t = gauss_;
            has_gauss = 0;
            gauss_ = 0.0;
            return t;
        }
        double f, x1, x2, r2;
        do {
            x1 = 2.0 * next_double() - 1.0;
            x2 = 2.0 * next_double() - 1.0;
            r2 = x1 * x1 + x2 * x2;
        } while (r2 >= 1.0 || r2 == 0.0);
        f = std::sqrt(-2.0 * std::log(r2) / r2);  // libm log, as numpy's legacy_gauss
        gauss_ = f * x1;
        has_gauss = 1;
        return f * x2;
    }
};

}  // namespace ocrng
