// Division of a 31-bit index by an invariant divisor as multiply-high + shift, with the constants computed once on the host.
// Used by the K1c vector kernels for (vector index -> row, offset) and (offset -> input-channel group); plain C++ so that the
// CPU test suite can compile it with g++ and check it against the built-in division (tests/test_host_cpu.py).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SSQ_HD __host__ __device__ __forceinline__
#else
#define SSQ_HD inline
#endif

namespace ssq {

struct FastDiv { uint32_t mul, sh; };            // n / d for n < 2^31, d >= 2: umulhi(n, mul) >> sh

// s = ceil(log2 d) >= 1, mul = floor(2^(31+s) / d) + 1 (< 2^32 because d > 2^(s-1)), sh = s - 1.
// Exactness: mul * d = 2^(31+s) + e with 0 < e <= d <= 2^s, so n * mul / 2^(31+s) = n / d + n * e / (d * 2^(31+s)), and the
// error term stays below 1 / d for every n < 2^31 — the floor cannot move to the next integer.
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    uint32_t s = 0;
    while ((1ull << s) < (uint64_t)d) ++s;
    if (s < 1) s = 1;                              // d == 1 is never divided by (callers special-case it); keep sh valid
    f.mul = (uint32_t)(((1ull << (31 + s)) / d) + 1ull);
    f.sh = s - 1;
    return f;
}

SSQ_HD uint32_t fastdiv(uint32_t n, const FastDiv& f) {
    return (uint32_t)(((uint64_t)n * f.mul) >> 32) >> f.sh;   // IMAD.HI + SHF on the device
}

}  // namespace ssq
