// oc_rng.h -- numpy's LEGACY random stream (np.random.seed / RandomState: MT19937), restated for the host side of the
// library so that the per-step randomness of the GCFM sweep and the crowd placement can be drawn in C with exactly the
// reference's stream (simulations.py:130-131,140,271,303):
//   next_double      genrand_res53
//   interval(max)    random_interval: masked rejection on 32-bit draws (max < 2^32)
//   permutation(n)   RandomState.permutation(n) == choice(arange(n), n, replace=False): Fisher-Yates from the top
//   gauss()          legacy_gauss: polar Box-Muller with one cached value (has_gauss / cached_gaussian of get_state())
// Pinned against numpy itself by tests/test_cpu_host.py (bit-identical draws and generator state).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

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
    // n consecutive gauss() values into out[], same values and same final state as n calls (the generator is advanced
    // first -- the rejection test needs only x1, x2 -- then f = sqrt(-2 log(r2) / r2) is evaluated).  ckpt (optional): snapshots of the generator state as it is after
    // value number n - 2q, q = 0 .. n_ckpt-1 (q-th snapshot: key at ckpt_key + 624 q, ...), for a caller that drew more
    // values than it turns out to need (look-ahead of the GCFM step: values come in pairs, n even).
    void gauss_fill(double *out, long long n, int n_ckpt = 0, uint32_t *ckpt_key = nullptr, int *ckpt_pos = nullptr,
                    int *ckpt_has = nullptr, double *ckpt_cached = nullptr) {
        long long idx = 0;
        const int had = has_gauss;
        const double had_val = gauss_;
        if (n > 0 && has_gauss) { out[idx++] = gauss_; has_gauss = 0; gauss_ = 0.0; }
        const long long rem = n - idx, n_ev = (rem + 1) / 2;
        std::vector<double> X1(n_ev), X2(n_ev), R2(n_ev);
        // snapshot q is wanted after value n - 2q = after event e_q = (n - 2q - idx + 1) / 2 - 1 (or the start state)
        auto snap = [&](int q) {
            if (!ckpt_key) return;
            std::memcpy(ckpt_key + (size_t)624 * q, key, 624 * sizeof(uint32_t));
            ckpt_pos[q] = pos;
        };
        std::vector<long long> ev_of(n_ckpt, -2);
        for (int q = 0; q < n_ckpt; q++) {
            const long long v = n - 2ll * q;            // values drawn at that point
            if (v < 0) continue;
            ev_of[q] = (v <= idx) ? -1 : (v - idx + 1) / 2 - 1;
            if (ev_of[q] == -1) snap(q);                // the state before any new event (only the cached value was used)
        }
        long long first_ck_ev = n_ev;
        for (int q = 0; q < n_ckpt; q++)
            if (ev_of[q] >= 0) first_ck_ev = std::min(first_ck_ev, ev_of[q]);
        for (long long e = 0; e < n_ev; e++) {
            double x1, x2, r2;
            do {
                x1 = 2.0 * next_double() - 1.0;
                x2 = 2.0 * next_double() - 1.0;
                r2 = x1 * x1 + x2 * x2;
            } while (r2 >= 1.0 || r2 == 0.0);
            X1[e] = x1; X2[e] = x2; R2[e] = r2;
            if (e >= first_ck_ev)   // the snapshots sit at the last n_ckpt events
                for (int q = 0; q < n_ckpt; q++)
                    if (ev_of[q] == e) snap(q);
        }
        // the transform, one event at a time (measured: helper threads / OpenMP teams cost more per call than the ~0.3 ms
        // they could save at 12.5 k pairs; the run loop overlaps this whole function with the GPU step instead)
        for (long long e = 0; e < n_ev; e++) {
            const double f = std::sqrt(-2.0 * std::log(R2[e]) / R2[e]);
            out[idx + 2 * e] = f * X2[e];
            if (idx + 2 * e + 1 < n) out[idx + 2 * e + 1] = f * X1[e];
        }
        // cache flag / value as they are after v values (v = n for the generator itself, n - 2q for the snapshots)
        auto cached_after = [&](long long v, int *has, double *cv) {
            if (v == 0) { *has = had; *cv = had_val; return; }
            if (v <= idx || (v - idx) % 2 == 0) { *has = 0; *cv = 0.0; return; }
            const long long e = (v - idx + 1) / 2 - 1;      // the event whose second value is still cached
            *has = 1;
            *cv = std::sqrt(-2.0 * std::log(R2[e]) / R2[e]) * X1[e];
        };
        cached_after(n, &has_gauss, &gauss_);
        for (int q = 0; q < n_ckpt; q++) {
            const long long v = n - 2ll * q;
            if (v < 0 || !ckpt_key) continue;
            cached_after(v, &ckpt_has[q], &ckpt_cached[q]);
        }
    }
};

}  // namespace ocrng
