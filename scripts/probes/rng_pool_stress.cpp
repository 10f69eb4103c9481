// stress test of the C legacy RNG transform pool (host only): 3000 back-to-back gauss_fill jobs against the serial stream.
// g++ -O2 -std=c++17 -o rng_pool_stress rng_pool_stress.cpp -lpthread && OC_RNG_THREADS=7 ./rng_pool_stress
#include "../../optimal_crowds_b200/csrc/oc_rng.h"
#include <cstdio>
int main() {
    std::vector<uint32_t> key(624), key2(624);
    for (int i = 0; i < 624; i++) key[i] = key2[i] = 1812433253u * (i + 1) + 99u;
    ocrng::Mt a{key.data(), 624, 0, 0.0};
    ocrng::Mt b{key2.data(), 624, 0, 0.0};
    std::vector<double> oa(40000), ob(40000);
    long long bad = 0;
    for (int it = 0; it < 3000; it++) {
        const long long n = 2 * (4096 + (it * 37) % 9000);   // jobs of 8..25 blocks, back to back
        a.gauss_fill(oa.data(), n);                          // pool (threads from OC_RNG_THREADS)
        for (long long q = 0; q < n; q++) ob[q] = b.gauss(); // serial reference
        for (long long q = 0; q < n; q++) bad += oa[q] != ob[q];
        if (a.pos != b.pos || a.has_gauss != b.has_gauss) bad++;
    }
    printf("threads %d mismatches %lld\n", ocrng::TransformPool::get().threads(), bad);
    return bad != 0;
}
