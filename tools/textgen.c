/* textgen.c — order-2 Markov text sampler: splitmix64, inverse-CDF on 32-bit draws. */
#include <stddef.h>
#include <stdint.h>

static inline uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* thr: [97*97][97] cumulative thresholds scaled to 2^32 (last entry of a row = 0xffffffff); alpha: the 97 byte values */
void orz_textgen(const uint32_t *thr, const uint8_t *alpha, uint64_t seed, uint8_t *out, size_t n) {
    uint64_t st = seed;
    unsigned a = 2, b = 2; /* initial state "  " (two spaces: symbol index 2) */
    for (size_t i = 0; i < n; i++) {
        const uint32_t *row = thr + (size_t)(a * 97 + b) * 97;
        uint32_t r = (uint32_t)(splitmix64(&st) >> 32);
        unsigned lo = 0, hi = 96;
        while (lo < hi) {
            unsigned mid = (lo + hi) >> 1;
            if (r <= row[mid]) hi = mid;
            else lo = mid + 1;
        }
        out[i] = alpha[lo];
        a = b;
        b = lo;
    }
}
